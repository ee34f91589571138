#!/usr/bin/env python
"""Summarise ONE ncu pass over a composed-query step into profiles/:
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
      --log-file gpurun_out/<TAG>_ncu_qstep.csv python tests/gpu_prof_qstep.py <Bq> 1
-> profiles/<TAG>_ncu_qstep_summary.txt (per-kernel launches, time share, DRAM bytes) and the GEMM fields of
   profiles/ncu_traffic.json (bench.py roofline.traffic).   usage: python tools/summarize_qstep_ncu.py TAG Bq"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG, BQ = sys.argv[1], int(sys.argv[2])
src = os.path.join(ROOT, "gpurun_out", f"{TAG}_ncu_qstep.csv")
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
idx = {h: i for i, h in enumerate(rows[hi])}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0,
         "nsecond": 1e-3, "msecond": 1e3}
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) != len(rows[hi]):
        continue
    d = per.setdefault(r[idx["ID"]], {"name": r[idx["Kernel Name"]]})
    d[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", "")) * scale.get(r[idx["Metric Unit"]], 1.0)


def short(n):
    return re.sub(r"^void ", "", n).replace("sprc::", "").replace("(anonymous namespace)::", "").split("(")[0]


# pack2d_kernel = weight packing inside load_state_dict, before the step
ours = [d for d in per.values() if "sprc" in d["name"] and "pack2d_kernel" not in d["name"]]
agg = collections.OrderedDict()
for d in ours:
    a = agg.setdefault(short(d["name"]), [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
out = [f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv  "
       f"python tests/gpu_prof_qstep.py {BQ} 1   ({TAG})",
       f"one composed-query step of {BQ} queries (fusion + text pass of the ViT-L Q-Former; bench.py adds one scan and "
       "one merge launch per step); per-launch times are cold-cache and serialised,",
       "so the SHARE of the step is the comparable figure (bench.py roofline.share_of_step is the live CUDA-event share).",
       ""]
for k, (n, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k:52s} n={n:4d} us={us:9.1f} share={100 * us / tot:5.1f}% avg={us / n:7.1f}  dram {by / n / 1e6:8.1f} MB/launch")
out.append(f"total us {tot:.1f}  launches {len(ours)}")
open(os.path.join(ROOT, "profiles", f"{TAG}_ncu_qstep_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
gemm = [d for d in ours if "gemm_bf16_tcgen05" in d["name"]]
tj = os.path.join(ROOT, "profiles", "ncu_traffic.json")
t = json.load(open(tj)) if os.path.exists(tj) else {}
t["gemm_bytes_per_launch"] = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in gemm) / len(gemm)
t["gemm_launches"] = len(gemm)
t["gemm_source"] = (f"mean of dram__bytes_read.sum + dram__bytes_write.sum over the {len(gemm)} tcgen05 GEMM launches of one "
                    f"{BQ}-query fusion step (ncu, tests/gpu_prof_qstep.py, {TAG})")
json.dump(t, open(tj, "w"), indent=1)
print(t)
