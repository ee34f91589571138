# r02m: where the value -> e2e gap comes from (pipeline diagnostics), strings/pipelined vs ids/serial
O=gpurun_out; T=${1:-r02m}; mkdir -p $O
timeout 900 python bench.py --no-cpu-baseline --no-vitg --no-eager-gpu --no-rerank --index-images 4096 --steps 10 > $O/${T}_bench_strings.log 2>&1
SPRC_E2E_IDS=1 timeout 900 python bench.py --no-cpu-baseline --no-vitg --no-eager-gpu --no-rerank --index-images 4096 --steps 10 > $O/${T}_bench_ids.log 2>&1
python - <<PY
import json
for f in ("$O/${T}_bench_strings.log", "$O/${T}_bench_ids.log"):
    l=[x for x in open(f) if x.startswith("{")][-1]; d=json.loads(l)
    print(f, round(d["value"]), round(d["value_repeat_after_e2e"]), round(d["e2e"]["value"]), d["e2e"].get("pipeline"), d["clocks"])
PY
