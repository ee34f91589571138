#!/bin/bash
# scan with N=128 MMAs over slot pairs
O=gpurun_out; mkdir -p $O; T=exp3
timeout 300 python -m pytest tests/test_scan_gpu.py -m gpu -x -q > $O/${T}_scan.log 2>&1; echo "EXIT=$?" >> $O/${T}_scan.log
COMMON="--steps 20 --warmup 3 --index-images 2048 --no-cpu-baseline"
timeout 300 python bench.py $COMMON > $O/${T}_bench.log 2>&1
timeout 300 python bench.py $COMMON --batch 128 > $O/${T}_bench_b128.log 2>&1
ls -la $O | tail -4
