# r03n (8 GPUs): two-stream pipelined multi-GPU step vs the lockstep step, back to back on the same box
O=gpurun_out; T=${1:-r03n}; mkdir -p $O
for mode in 0 1 0; do
SPRC_BENCH_LOCKSTEP=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2953$mode bench.py --gpus 8 --steps 20 --warmup 3 --no-vitg --no-rerank > $O/${T}_bench_n8_lockstep$mode.log 2> $O/${T}_bench_n8_lockstep$mode.err
python - <<PY
import json
f = "$O/${T}_bench_n8_lockstep$mode"
l=[x for x in open(f + ".log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("lockstep", $mode, round(d["value"]), round(d["value_repeat_after_e2e"]), round(d["e2e"]["value"]), d["ms_per_step"], d["sharded_equals_single"], d["rank_skew"]["kernel_ms_per_step"], d["rank_skew"]["sm_mhz"])
else:
    print(open(f + ".err").read()[-2000:])
PY
done
