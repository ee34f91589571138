# r03c: full GPU suite after the fold removal / grouped scan, default bench with the same-shape GEMM points
O=gpurun_out; T=${1:-r03c}; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/${T}_gpu_tests.log 2>&1; echo EXIT=$? >> $O/${T}_gpu_tests.log
grep -E "passed|failed|EXIT" $O/${T}_gpu_tests.log | tail -4
s0=$(date +%s); timeout 1500 python bench.py > $O/${T}_bench.log 2> $O/${T}_bench.err; echo "bench wall $(( $(date +%s) - s0 )) s" | tee -a $O/${T}_bench.err
python - <<PY
import json
l=[x for x in open("$O/${T}_bench.log") if x.startswith("{")][-1]; d=json.loads(l)
print(round(d["value"]), round(d["e2e"]["value"]), d["roofline"]["frac"], d["step_breakdown_ms"], d["parity"]["pass"], d["clocks"])
print(json.dumps(d["roofline"]["same_shapes_at_the_power_cap"])[:900])
print(json.dumps(d["index_feed"])[:300])
PY
