O=gpurun_out; T=${1:-r02e}; mkdir -p $O
for d in 0 1 2 3; do for dh in 64 88; do SPRC_VIT_ATTN_DBG=$d timeout 120 python tests/gpu_prof_attn.py 128 $dh 2>&1 | sed "s/^/dbg=$d /" >> $O/${T}_vit_attn_dbg.log; done; done
cat $O/${T}_vit_attn_dbg.log
