# r02r: PNG feed with the own inflater (tests + index_feed row), short bench
O=gpurun_out; T=${1:-r02r}; mkdir -p $O
timeout 600 python -m pytest tests/test_png.py tests/test_dropin_gpu.py -q -k "png" > $O/${T}_png_tests.log 2>&1; echo EXIT=$? >> $O/${T}_png_tests.log
grep -E "passed|failed|EXIT" $O/${T}_png_tests.log | tail -3
timeout 900 python bench.py --no-cpu-baseline --no-vitg --no-eager-gpu --no-rerank --index-images 8192 --steps 10 > $O/${T}_bench.log 2>&1
python - <<PY
import json
l=[x for x in open("$O/${T}_bench.log") if x.startswith("{")][-1]; d=json.loads(l)
print(round(d["value"]), round(d["e2e"]["value"]), d["clocks"]["sm_mhz"], json.dumps(d["index_feed"])[:900]); print(json.dumps(d["index_build"]["gemm_kernels"])[:300])
PY
