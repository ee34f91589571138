# r03e: K / V halves of the cross-attention stage on separate barriers: op tests, parity, step profile
O=gpurun_out; T=${1:-r03e}; mkdir -p $O
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "cross_attention" > $O/${T}_cross_tests.log 2>&1; echo EXIT=$? >> $O/${T}_cross_tests.log
grep -E "passed|failed|EXIT|Error|assert" $O/${T}_cross_tests.log | tail -4
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x > $O/${T}_parity.log 2>&1; echo EXIT=$? >> $O/${T}_parity.log
grep -E "passed|failed|EXIT" $O/${T}_parity.log | tail -3
timeout 900 python bench.py --no-cpu-baseline --no-vitg --no-eager-gpu --no-index-feed --no-gemm-points --index-images 8192 --steps 10 --profile-dump $O/${T}_prof > $O/${T}_bench.log 2>&1
python tools/show_profile.py $O/${T}_prof.query.csv 10 2>/dev/null | grep -E "cross|total"
python - <<PY
import json
l=[x for x in open("$O/${T}_bench.log") if x.startswith("{")][-1]; d=json.loads(l)
print(round(d["value"]), round(d["e2e"]["value"]), d["step_breakdown_ms"], d["parity"]["pass"], d["clocks"]["sm_mhz"], d["rerank"]["pairs_per_s_per_gpu"])
PY
