# r02o: two-group cross-attention (op tests, parity, timing against the one-group kernel), native PNG feeder test + index_feed row
O=gpurun_out; T=${1:-r02o}; mkdir -p $O
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -s -k "cross_attention" > $O/${T}_cross_tests.log 2>&1; echo EXIT=$? >> $O/${T}_cross_tests.log
grep -E "passed|failed|EXIT|Error|error|assert" $O/${T}_cross_tests.log | tail -8
timeout 600 python -m pytest tests/test_dropin_gpu.py -m gpu -q -x -s -k "png_feeder" > $O/${T}_png_feeder_test.log 2>&1; echo EXIT=$? >> $O/${T}_png_feeder_test.log
grep -E "^\[png|passed|failed|EXIT|Error|error|assert" $O/${T}_png_feeder_test.log | tail -8
if grep -q "EXIT=0" $O/${T}_cross_tests.log; then
  timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x > $O/${T}_parity.log 2>&1; echo EXIT=$? >> $O/${T}_parity.log
  grep -E "passed|failed|EXIT" $O/${T}_parity.log | tail -4
  SPRC_CROSS_ATTN_1G=1 timeout 600 python bench.py --no-cpu-baseline --no-vitg --no-eager-gpu --no-rerank --no-index-feed --index-images 2048 --steps 10 --profile-dump $O/${T}_prof1g > $O/${T}_bench_1g.log 2>&1
  timeout 900 python bench.py --no-cpu-baseline --no-vitg --no-eager-gpu --index-images 2048 --steps 10 --profile-dump $O/${T}_prof > $O/${T}_bench.log 2>&1
  python - <<PY
import json
for f in ("$O/${T}_bench_1g.log", "$O/${T}_bench.log"):
    l=[x for x in open(f) if x.startswith("{")][-1]; d=json.loads(l)
    print(f, round(d["value"]), round(d["e2e"]["value"]), d["step_breakdown_ms"], d["parity"]["pass"], d["clocks"]["sm_mhz"], json.dumps(d.get("index_feed"))[:900])
PY
  python tools/show_profile.py $O/${T}_prof1g.query.csv 10 | grep -E "cross|total"
  python tools/show_profile.py $O/${T}_prof.query.csv 10 | grep -E "cross|total"
fi
