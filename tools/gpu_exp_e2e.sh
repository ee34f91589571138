O=gpurun_out; T=r01k
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "host_end_to_end or ragged or invariant or rerank" > $O/${T}_hosttest.log 2>&1; echo EXIT=$? >> $O/${T}_hosttest.log
SPRC_E2E_PIPELINE=0 timeout 300 python bench.py --steps 20 --warmup 3 --index-images 2048 --no-cpu-baseline > $O/${T}_bench_serial.log 2>&1
SPRC_E2E_PIPELINE=1 timeout 300 python bench.py --steps 20 --warmup 3 --index-images 2048 --no-cpu-baseline > $O/${T}_bench_pipelined.log 2>&1
tail -4 $O/${T}_hosttest.log
for f in $O/${T}_bench_serial.log $O/${T}_bench_pipelined.log; do python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], d['clocks'])
PY
done
