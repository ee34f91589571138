#!/bin/bash
# fused GEMM+LN kernel + scan NQ=2: tests, then A/B bench
O=gpurun_out; mkdir -p $O; T=exp2
timeout 180 python -m pytest tests/test_ops_gpu.py -m gpu -x -q > $O/${T}_ops.log 2>&1; echo "EXIT=$?" >> $O/${T}_ops.log
timeout 300 python -m pytest tests/test_scan_gpu.py -m gpu -x -q > $O/${T}_scan.log 2>&1; echo "EXIT=$?" >> $O/${T}_scan.log
timeout 400 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > $O/${T}_parity.log 2>&1; echo "EXIT=$?" >> $O/${T}_parity.log
COMMON="--steps 20 --warmup 3 --index-images 2048 --no-cpu-baseline"
export SPRC_DEBUG=1
timeout 300 python bench.py $COMMON --profile-dump $O/${T}_shapes > $O/${T}_bench_fused1.log 2>&1
SPRC_FUSED_LN=0 timeout 300 python bench.py $COMMON > $O/${T}_bench_fused0.log 2>&1
timeout 300 python bench.py $COMMON --batch 588 > $O/${T}_bench_fused1_b588.log 2>&1
timeout 300 python bench.py $COMMON --batch 576 > $O/${T}_bench_fused1_b576.log 2>&1
ls -la $O | tail -8
