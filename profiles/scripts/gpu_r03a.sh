# r03a: LayerNorm fold v2 wired into the model: op tests, reference parity with SPRC_LN_FOLD=1, bench fold vs default
O=gpurun_out; T=${1:-r03a}; mkdir -p $O
timeout 900 python -m pytest tests/test_ln_fold_gpu.py -m gpu -q -x -k "not other_schedule" > $O/${T}_fold_op_tests.log 2>&1; echo EXIT=$? >> $O/${T}_fold_op_tests.log
grep -E "passed|failed|EXIT|Error|assert" $O/${T}_fold_op_tests.log | tail -5
SPRC_LN_FOLD=1 timeout 1500 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -s > $O/${T}_parity_fold.log 2>&1; echo EXIT=$? >> $O/${T}_parity_fold.log
grep -E "^\[|passed|failed|EXIT|Error|assert" $O/${T}_parity_fold.log | tail -30
for v in 0 1; do
  SPRC_LN_FOLD=$v timeout 900 python bench.py --no-cpu-baseline --no-vitg --no-eager-gpu --no-rerank --no-index-feed --index-images 8192 --steps 10 --profile-dump $O/${T}_prof_fold$v > $O/${T}_bench_fold$v.log 2>&1
done
python - <<PY
import json
for v in (0, 1):
    try:
        l=[x for x in open("$O/${T}_bench_fold%d.log" % v) if x.startswith("{")][-1]; d=json.loads(l)
        print("fold", v, round(d["value"]), round(d["e2e"]["value"]), d["step_breakdown_ms"], d["parity"]["pass"], d["parity"]["relF_raws"], d["parity"]["relF_feats"], d["parity"]["max_abs_dsim"], d["clocks"]["sm_mhz"], round(d["index_build"]["images_per_s_per_gpu"]))
    except Exception as e:
        print("fold", v, "failed", e); print(open("$O/${T}_bench_fold%d.log" % v).read()[-1500:])
PY
