# r03m (2 GPUs): two-stream pipelined multi-GPU step vs the lockstep step
O=gpurun_out; T=${1:-r03m}; mkdir -p $O
for mode in 0 1; do
SPRC_BENCH_LOCKSTEP=$mode timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$mode bench.py --gpus 2 --steps 20 --warmup 3 --no-vitg --no-rerank > $O/${T}_bench_n2_lockstep$mode.log 2> $O/${T}_bench_n2_lockstep$mode.err
done
python - <<PY
import json
for mode in (0, 1):
    f = "$O/${T}_bench_n2_lockstep%d" % mode
    l=[x for x in open(f + ".log") if x.startswith("{")]
    if l:
        d=json.loads(l[-1]); print("lockstep", mode, round(d["value"]), round(d["value_repeat_after_e2e"]), round(d["e2e"]["value"]), d["ms_per_step"], d["sharded_equals_single"], d["rank_skew"]["kernel_ms_per_step"])
    else:
        print(open(f + ".err").read()[-3000:])
PY
