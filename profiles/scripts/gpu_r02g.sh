# r02g: ncu --set full of the ViT attention kernel (v2) + cross-attention v2 op tests are run separately
O=gpurun_out; T=${1:-r02g}; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vit_attention_v2 -s 3 -c 1 -o $O/${T}_vit_attn python tests/gpu_prof_attn.py 128 64 > $O/${T}_ncu_run.log 2>&1
ls -la $O/${T}_vit_attn.ncu-rep
