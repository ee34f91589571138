# r02h: second-generation cross-attention (257 / 514 keys): op tests, model-level parity, rerank + bench timing
O=gpurun_out; T=${1:-r02h}; mkdir -p $O
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -s -k "cross_attention" > $O/${T}_cross_tests.log 2>&1; echo EXIT=$? >> $O/${T}_cross_tests.log
grep -E "^\[|passed|failed|EXIT|Error|error|assert" $O/${T}_cross_tests.log | tail -20
if grep -q "EXIT=0" $O/${T}_cross_tests.log; then
  timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -s > $O/${T}_parity.log 2>&1; echo EXIT=$? >> $O/${T}_parity.log
  grep -E "^\[|passed|failed|EXIT" $O/${T}_parity.log | tail -30
  SPRC_CROSS_ATTN_V1=1 timeout 300 python tests/gpu_bench_rerank.py 8 100 > $O/${T}_rerank_v1.log 2>&1
  timeout 300 python tests/gpu_bench_rerank.py 8 100 > $O/${T}_rerank_v2.log 2>&1; mv $O/rerank_shapes.csv $O/${T}_rerank_shapes.csv
  tail -1 $O/${T}_rerank_v1.log; tail -1 $O/${T}_rerank_v2.log
  timeout 600 python bench.py --no-cpu-baseline --no-vitg --no-eager-gpu --index-images 8192 --steps 10 --profile-dump $O/${T}_prof > $O/${T}_bench.log 2>&1
  python tools/show_profile.py $O/${T}_prof.query.csv 10 2>/dev/null | head -6
  python - <<PY
import json
l=[x for x in open("$O/${T}_bench.log") if x.startswith("{")][-1]; d=json.loads(l); print(d["value"], d["e2e"]["value"], d["step_breakdown_ms"], d["rerank"]["pairs_per_s_per_gpu"], d["parity"]["pass"], d["clocks"])
PY
fi
