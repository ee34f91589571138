# r03f: default bench after allocating the pinned rings up front (value / e2e / pipeline diagnostics), twice
O=gpurun_out; T=${1:-r03f}; mkdir -p $O
for i in 1 2; do
timeout 900 python bench.py --no-cpu-baseline --no-vitg --no-eager-gpu --no-index-feed --no-gemm-points --no-rerank --index-images 8192 > $O/${T}_bench$i.log 2>&1
done
python - <<PY
import json
for i in (1, 2):
    l=[x for x in open("$O/${T}_bench%d.log" % i) if x.startswith("{")][-1]; d=json.loads(l)
    print(round(d["value"]), round(d["value_repeat_after_e2e"]), round(d["e2e"]["value"]), d["ms_per_step"], d["e2e"]["pipeline"]["host_submit_ms_per_step"], d["e2e"]["pipeline"]["gpu_busy_ms_per_step"], d["clocks"]["sm_mhz"])
PY
