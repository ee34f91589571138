O=gpurun_out; T=r01l
for mode in 1 0; do
SPRC_E2E_PIPELINE=$mode timeout 300 python bench.py --steps 20 --warmup 3 --index-images 2048 --no-cpu-baseline > $O/${T}_bench_p$mode.log 2>&1
python - "$O/${T}_bench_p$mode.log" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'repeat', round(d['value_repeat_after_e2e']), d['clocks'])
PY
done
