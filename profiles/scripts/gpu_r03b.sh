# r03b (2 GPUs): grouped-output scan in the multi-GPU step: NCCL dist test + bench at N = 2
O=gpurun_out; T=${1:-r03b}; mkdir -p $O
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q -x -s > $O/${T}_dist_tests.log 2>&1; echo EXIT=$? >> $O/${T}_dist_tests.log
grep -E "passed|failed|EXIT|skipped" $O/${T}_dist_tests.log | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 --no-vitg --no-rerank > $O/${T}_bench_n2.log 2> $O/${T}_bench_n2.err
python - <<PY
import json
l=[x for x in open("$O/${T}_bench_n2.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print(round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["step_breakdown_ms"], d["sharded_equals_single"], d["clocks"]["sm_mhz"], d["roofline_scan"]["avg_launch_us"])
else:
    print(open("$O/${T}_bench_n2.err").read()[-3000:])
PY
