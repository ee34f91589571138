#!/bin/bash
# scan v2 + sweep-direction experiment
O=gpurun_out; mkdir -p $O; T=exp1
timeout 600 python -m pytest tests/test_scan_gpu.py tests/test_parity_gpu.py -m gpu -x -q > $O/${T}_tests.log 2>&1; echo "EXIT=$?" >> $O/${T}_tests.log
COMMON="--steps 20 --warmup 3 --index-images 2048 --no-cpu-baseline"
timeout 300 python bench.py $COMMON --profile-dump $O/${T}_shapes > $O/${T}_bench_sweep1.log 2>&1
SPRC_SWEEP=0 timeout 300 python bench.py $COMMON > $O/${T}_bench_sweep0.log 2>&1
timeout 300 python bench.py $COMMON --batch 296 > $O/${T}_bench_b296.log 2>&1
timeout 300 python bench.py $COMMON --batch 148 > $O/${T}_bench_b148.log 2>&1
timeout 300 python bench.py $COMMON --batch 1184 > $O/${T}_bench_b1184.log 2>&1
ls -la $O | tail -12
