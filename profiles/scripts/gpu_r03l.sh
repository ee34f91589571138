# r03l (8 GPUs): where do the 9 ms between rank 0's kernels and the lockstep step go?  rank_skew diagnostics
O=gpurun_out; T=${1:-r03l}; mkdir -p $O
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 20 --warmup 3 --no-vitg --no-rerank > $O/${T}_bench_n8.log 2> $O/${T}_bench_n8.err
python - <<PY
import json
l=[x for x in open("$O/${T}_bench_n8.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print(round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["step_breakdown_ms"], d["rank_skew"])
else:
    print(open("$O/${T}_bench_n8.err").read()[-3000:])
PY
