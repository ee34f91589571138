# r02l: full GPU suite (tie-aware submission test), sustained GEMM ours vs cuBLAS under the power cap
O=gpurun_out; T=${1:-r02l}; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/${T}_gpu_tests.log 2>&1; echo EXIT=$? >> $O/${T}_gpu_tests.log
grep -E "^\[submission|passed|failed|EXIT" $O/${T}_gpu_tests.log | tail -8
timeout 900 python tests/gpu_sustained_gemm.py 1.5 > $O/${T}_sustained_gemm.log 2>&1; cat $O/${T}_sustained_gemm.log | tail -14
