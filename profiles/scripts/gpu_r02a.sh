# r02a: LayerNorm fold on a GPU for the first time + fp16-vs-bf16 bench A/B in the same box.
O=gpurun_out; T=${1:-r02a}; mkdir -p $O
SPRC_TEST_LN_FOLD=1 timeout 500 python -m pytest tests/test_ln_fold_gpu.py -m gpu -q -s > $O/${T}_fold_tests.log 2>&1; echo EXIT=$? >> $O/${T}_fold_tests.log
tail -15 $O/${T}_fold_tests.log
timeout 300 python bench.py > $O/${T}_bench_bf16.log 2>$O/${T}_bench_bf16.err; echo EXIT=$? >> $O/${T}_bench_bf16.log
SPRC_ACT_DTYPE=fp16 timeout 300 python bench.py > $O/${T}_bench_fp16.log 2>$O/${T}_bench_fp16.err; echo EXIT=$? >> $O/${T}_bench_fp16.log
SPRC_LN_FOLD=1 timeout 300 python bench.py > $O/${T}_bench_fold.log 2>$O/${T}_bench_fold.err; echo EXIT=$? >> $O/${T}_bench_fold.log
for f in bf16 fp16 fold; do python - <<PY
import json
try:
    l=[x for x in open("$O/${T}_bench_$f.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("$f", d["value"], d["e2e"]["value"], d.get("step_breakdown_ms"), d["roofline"]["frac"], d["index_build"]["images_per_s_per_gpu"])
except Exception as e: print("$f", "ERR", e)
PY
done
