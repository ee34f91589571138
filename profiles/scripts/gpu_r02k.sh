# r02k: batch invariance / determinism of encode_gallery with the second-generation attention kernels vs the first
O=gpurun_out; T=${1:-r02k}; mkdir -p $O
(echo "== default (v2 kernels)"; timeout 300 python tests/gpu_diag_batchinv.py 2 2
 echo "== SPRC_VIT_ATTN_V1=1"; SPRC_VIT_ATTN_V1=1 timeout 300 python tests/gpu_diag_batchinv.py 2 2
 echo "== SPRC_CROSS_ATTN_V1=1"; SPRC_CROSS_ATTN_V1=1 timeout 300 python tests/gpu_diag_batchinv.py 2 2
 echo "== both v1"; SPRC_VIT_ATTN_V1=1 SPRC_CROSS_ATTN_V1=1 timeout 300 python tests/gpu_diag_batchinv.py 2 2
 echo "== SPRC_SWEEP=0"; SPRC_SWEEP=0 timeout 300 python tests/gpu_diag_batchinv.py 2 2
 echo "== both v1 + SPRC_SWEEP=0"; SPRC_SWEEP=0 SPRC_VIT_ATTN_V1=1 SPRC_CROSS_ATTN_V1=1 timeout 300 python tests/gpu_diag_batchinv.py 2 2
) > $O/${T}_batchinv.log 2>&1
cat $O/${T}_batchinv.log
