#!/bin/bash
# packed GELU + head-major K/V
O=gpurun_out; mkdir -p $O; T=exp6
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_parity_gpu.py -m gpu -x -q > $O/${T}_tests.log 2>&1; echo "EXIT=$?" >> $O/${T}_tests.log
timeout 100 python tests/gpu_diag.py attn > $O/${T}_attn_diag.log 2>&1
COMMON="--steps 20 --warmup 3 --index-images 4096 --no-cpu-baseline"
timeout 300 python bench.py $COMMON --profile-dump $O/${T}_shapes > $O/${T}_bench.log 2>&1
ls -la $O | tail -5
