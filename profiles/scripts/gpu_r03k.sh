# r03k (2 GPUs): bench with the rank_skew diagnostics
O=gpurun_out; T=${1:-r03k}; mkdir -p $O
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 --no-vitg --no-rerank > $O/${T}_bench_n2.log 2> $O/${T}_bench_n2.err
python - <<PY
import json
l=[x for x in open("$O/${T}_bench_n2.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print(round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["sharded_equals_single"], d["rank_skew"])
else:
    print(open("$O/${T}_bench_n2.err").read()[-3000:])
PY
