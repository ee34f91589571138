# r02j: full GPU suite at HEAD, index-batch sweep (ViT-L), ncu --set full of the ViT residual / fc1 GEMMs and a ViT LayerNorm
O=gpurun_out; T=${1:-r02j}; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/${T}_gpu_tests.log 2>&1; echo EXIT=$? >> $O/${T}_gpu_tests.log
tail -4 $O/${T}_gpu_tests.log
timeout 600 python tests/gpu_index_sweep.py clip_L 64 96 128 192 256 384 > $O/${T}_index_sweep.log 2>&1; cat $O/${T}_index_sweep.log | tail -8
for f in $O/index_sweep_*.csv; do mv $f $O/${T}_$(basename $f); done
# ncu: proj GEMM (fp32 residual reduce-add), fc1 GEMM (QuickGELU, 16-bit out)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -c 2 -f -o $O/${T}_gemm_proj \
  python tests/gpu_prof_gemm.py 32896 1024 1024 0 1 1 0 2 > $O/${T}_ncu_proj.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -c 2 -f -o $O/${T}_gemm_fc1 \
  python tests/gpu_prof_gemm.py 32896 4096 1024 2 0 0 0 2 > $O/${T}_ncu_fc1.log 2>&1
tail -2 $O/${T}_ncu_proj.log $O/${T}_ncu_fc1.log
ls -la $O | grep ${T}
