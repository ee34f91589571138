# r02s: PNG feed with the arena copied on a side stream (png feeder test + index_feed row)
O=gpurun_out; T=${1:-r02s}; mkdir -p $O
timeout 600 python -m pytest tests/test_dropin_gpu.py -m gpu -q -k "png or preprocess" > $O/${T}_png_tests.log 2>&1; echo EXIT=$? >> $O/${T}_png_tests.log
grep -E "passed|failed|EXIT" $O/${T}_png_tests.log | tail -3
timeout 900 python bench.py --no-cpu-baseline --no-vitg --no-eager-gpu --no-rerank --index-images 2048 --steps 5 > $O/${T}_bench.log 2>&1
python - <<PY
import json
l=[x for x in open("$O/${T}_bench.log") if x.startswith("{")][-1]; d=json.loads(l)
print(round(d["value"]), round(d["e2e"]["value"]), d["clocks"]["sm_mhz"], json.dumps(d["index_feed"])[:400])
PY
