#!/bin/bash
# ragged query passes
O=gpurun_out; mkdir -p $O; T=exp9
timeout 400 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > $O/${T}_parity.log 2>&1; echo "EXIT=$?" >> $O/${T}_parity.log
if grep -q "EXIT=0" $O/${T}_parity.log; then
COMMON="--steps 20 --warmup 3 --index-images 4096 --no-cpu-baseline"
timeout 300 python bench.py $COMMON --profile-dump $O/${T}_shapes > $O/${T}_bench_ragged1.log 2>&1
SPRC_RAGGED=0 timeout 300 python bench.py $COMMON > $O/${T}_bench_ragged0.log 2>&1
timeout 600 python -m pytest tests/test_dropin_gpu.py tests/test_scan_gpu.py -m gpu -x -q > $O/${T}_dropin.log 2>&1; echo "EXIT=$?" >> $O/${T}_dropin.log
fi
ls -la $O | tail -6
