O=gpurun_out; T=r01o
for b in 592 1184 2368; do
  timeout 250 python bench.py --steps 20 --warmup 3 --index-images 2048 --batch $b --no-cpu-baseline > $O/${T}_b$b.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r01o_*.log")):
    ok=False
    for l in open(f):
        if l.startswith("{"):
            ok=True
            d=json.loads(l); print(f, "q/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "repeat", round(d["value_repeat_after_e2e"]), "gemm TF", round(d["roofline"]["achieved"]), d["step_breakdown_ms"], d["clocks"])
    if not ok: print(f, open(f).read()[-600:])
PY
