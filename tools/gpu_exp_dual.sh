O=gpurun_out; T=r01i
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "two_weight" > $O/${T}_ops.log 2>&1; echo EXIT=$? >> $O/${T}_ops.log
timeout 400 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -s -k "golden or ragged or invariant" > $O/${T}_parity.log 2>&1; echo EXIT=$? >> $O/${T}_parity.log
SPRC_DUAL_FFN=0 timeout 300 python bench.py --steps 20 --warmup 3 --index-images 2048 --no-cpu-baseline > $O/${T}_bench_dual0.log 2>&1
SPRC_DUAL_FFN=1 timeout 300 python bench.py --steps 20 --warmup 3 --index-images 2048 --no-cpu-baseline --profile-dump $O/${T}_shapes > $O/${T}_bench_dual1.log 2>&1
tail -3 $O/${T}_ops.log; tail -3 $O/${T}_parity.log
for f in $O/${T}_bench_dual0.log $O/${T}_bench_dual1.log; do python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], round(d['value']), d['ms_per_step'], d['roofline']['achieved'], d['step_breakdown_ms'], d['clocks'])
PY
done
