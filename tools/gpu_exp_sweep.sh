O=gpurun_out; T=r01n
for ib in 64 96 192 256; do
  timeout 200 python bench.py --steps 3 --warmup 3 --index-images 6144 --index-batch $ib --no-cpu-baseline > $O/${T}_ib$ib.log 2>&1
done
for b in 1184 888; do
  timeout 200 python bench.py --steps 12 --warmup 3 --index-images 2048 --batch $b --no-cpu-baseline > $O/${T}_b$b.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r01n_*.log")):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print(f, "q/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "index img/s", round(d["index_build"]["images_per_s_per_gpu"]), d["step_breakdown_ms"])
PY
