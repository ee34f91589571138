# r02u (8 GPUs): the driver's own launch line at N = 8
O=gpurun_out; T=${1:-r02u}; mkdir -p $O
s0=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 3 > $O/${T}_bench_n8.log 2> $O/${T}_bench_n8.err
echo "wall $(( $(date +%s) - s0 )) s rc=$?" | tee -a $O/${T}_bench_n8.err
python - <<PY
import json
l=[x for x in open("$O/${T}_bench_n8.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print(round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["step_breakdown_ms"], d["sharded_equals_single"], d["clocks"], json.dumps(d["vit_g"])[:500], json.dumps(d["rerank"])[:200])
else:
    print(open("$O/${T}_bench_n8.err").read()[-3000:])
PY
