# r03t (4 GPUs): the driver's launch line at N = 4 (the one N not run yet with the final bench.py) + reference arm under torchrun
O=gpurun_out; T=${1:-r03t}; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 20 --warmup 3 > $O/${T}_bench_n4.log 2> $O/${T}_bench_n4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 4 --steps 3 --warmup 1 > $O/${T}_bench_n4_reference.log 2> $O/${T}_bench_n4_reference.err
python - <<PY
import json
l=[x for x in open("$O/${T}_bench_n4.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print(round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["sharded_equals_single"], d["rank_skew"]["kernel_ms_per_step"], json.dumps(d["vit_g"])[:300])
else:
    print(open("$O/${T}_bench_n4.err").read()[-2000:])
l=[x for x in open("$O/${T}_bench_n4_reference.log") if x.startswith("{")]
print(l[-1][:200] if l else open("$O/${T}_bench_n4_reference.err").read()[-1500:])
PY
