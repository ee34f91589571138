#!/bin/bash
# cta_group::2 GEMM: op tests, parity, kernel microbench, bench A/B
O=gpurun_out; mkdir -p $O; T=exp4
timeout 240 python -m pytest tests/test_ops_gpu.py -m gpu -x -q > $O/${T}_ops.log 2>&1; echo "EXIT=$?" >> $O/${T}_ops.log
if grep -q "EXIT=0" $O/${T}_ops.log; then
timeout 400 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > $O/${T}_parity.log 2>&1; echo "EXIT=$?" >> $O/${T}_parity.log
timeout 200 python tests/gpu_diag.py perf > $O/${T}_perf_2cta.log 2>&1
SPRC_GEMM_2CTA=0 timeout 200 python tests/gpu_diag.py perf > $O/${T}_perf_1cta.log 2>&1
COMMON="--steps 20 --warmup 3 --index-images 4096 --no-cpu-baseline"
timeout 300 python bench.py $COMMON --profile-dump $O/${T}_shapes > $O/${T}_bench_2cta.log 2>&1
SPRC_GEMM_2CTA=0 timeout 300 python bench.py $COMMON > $O/${T}_bench_1cta.log 2>&1
fi
ls -la $O | tail -8
