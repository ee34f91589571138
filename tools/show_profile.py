"""Pretty-print a per-shape profile CSV written by bench.py --profile-dump (sprc_profile_dump)."""
import csv
import sys

CATS = {0: "gemm", 1: "attn", 2: "layernorm", 3: "scan", 4: "merge"}
rows = list(csv.DictReader(open(sys.argv[1])))
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
tot = sum(float(r["total_ms"]) for r in rows)
print(f"total {tot/div:.3f} ms per step (div {div})")
for r in sorted(rows, key=lambda r: -float(r["total_ms"])):
    ms, n, fl, by = float(r["total_ms"]), float(r["launches"]), float(r["flops"]), float(r["bytes"])
    print(f"{CATS.get(int(r['cat']), r['cat']):9s} {r['tag']:44s} n/step={n/div:5.1f} ms/step={ms/div:7.3f} "
          f"({ms/tot*100:4.1f}%) avg={ms/n*1e3:7.1f} us  {fl/ms/1e9 if ms else 0:7.1f} TF/s  {by/ms/1e6 if ms else 0:7.1f} GB/s")
