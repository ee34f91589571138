#!/bin/bash
O=gpurun_out; mkdir -p $O; T=exp5
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_gpu_tests.log 2>&1; echo "EXIT=$?" >> $O/${T}_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "EXIT=$?" >> $O/${T}_smoke.log
timeout 200 python tests/gpu_diag.py perf > $O/${T}_perf.log 2>&1
timeout 300 python bench.py --vit eva_clip_g --steps 10 --warmup 3 --index-images 2048 --index-batch 64 --no-cpu-baseline > $O/${T}_bench_vitg.log 2>&1
ls -la $O | tail -5
