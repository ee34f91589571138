#!/usr/bin/env python
"""Summarise the ncu outputs of profiles/scripts/gpu_round.sh (gpurun_out/<TAG>_*) into profiles/:
  <TAG>_ncu_launch_summary.txt   per-kernel share of one query step (cold, serialised launch list)
  <TAG>_ncu_full_summary.txt     key metrics of the --set full captures
  ncu_traffic.json               dram bytes per launch (bench.py roofline.traffic)
usage: python tools/summarize_ncu.py TAG"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"


def read_metric_csv(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    return idx, data


def unit_scale(unit):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0,
            "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1.0)


def short(name):
    return re.sub(r"^void ", "", name).replace("sprc::", "").replace("(anonymous namespace)::", "").split("(")[0]


# ---- launch list ----
f = os.path.join(O, f"{TAG}_ncu_launches.csv")
if os.path.exists(f):
    idx, data = read_metric_csv(f)
    names = [r[idx["Kernel Name"]] for r in data]
    us = [float(r[idx["Metric Value"]].replace(",", "")) * unit_scale(r[idx["Metric Unit"]]) for r in data]
    scan = [i for i, n in enumerate(names) if "scan_topk_kernel" in n]
    # one device-timed query step = launches after a step's merge up to and including the next step's merge;
    # use the pair of scans in the middle of the list whose distance is the modal launch count
    if len(scan) >= 4:
        a, b = scan[len(scan) // 2 - 1] + 2, scan[len(scan) // 2] + 2
        agg = collections.OrderedDict()
        for n, v in zip(names[a:b], us[a:b]):
            d = agg.setdefault(short(n), [0, 0.0])
            d[0] += 1
            d[1] += v
        tot = sum(v for _, v in agg.values())
        out = [f"ncu --metrics gpu__time_duration.sum --clock-control none --csv  python bench.py --steps 2 --warmup 3 "
               f"--index-images 128 --no-cpu-baseline [...]   (tools/gpu_*.sh, tag {TAG})",
               "one query step of bench.py's default batch (2368 composed queries since r01p; ViT-L Q-Former, gallery 50k -> top-50); per-launch times are "
               "cold-cache and serialised,", "so the SHARE of the step is the comparable figure "
               "(bench.py roofline.share_of_step reports the live CUDA-event share).", ""]
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.append(f"{k:52s} n={n:4d} us={v:9.1f} share={100 * v / tot:5.1f}% avg={v / n:7.1f}")
        out.append(f"total us {tot:.1f}  launches {b - a}")
        open(os.path.join(P, f"{TAG}_ncu_launch_summary.txt"), "w").write("\n".join(out) + "\n")
        print("\n".join(out))

# ---- GEMM traffic ----
traffic = {}
f = os.path.join(O, f"{TAG}_ncu_gemm_traffic.csv")
if os.path.exists(f):
    idx, data = read_metric_csv(f)
    per = collections.defaultdict(dict)
    for r in data:
        per[r[idx["ID"]]][r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", "")) * unit_scale(
            r[idx["Metric Unit"]])
    tot = [d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in per.values()]
    if tot:
        traffic["gemm_bytes_per_launch"] = sum(tot) / len(tot)
        traffic["gemm_launches"] = len(tot)
        traffic["gemm_source"] = (f"mean of dram__bytes_read.sum + dram__bytes_write.sum over the {len(tot)} tcgen05 "
                                  f"GEMM launches of one 592-query fusion step (ncu, tests/gpu_prof_qstep.py, {TAG})")

# ---- full captures ----
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"),
        ("dram__bytes_write.sum", "dram write"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__cluster_size", "cluster"), ("launch__registers_per_thread", "regs/thread"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %")]
out = [f"ncu --set full --clock-control none --import-source on (profiles/scripts/gpu_round.sh {TAG}; one B200; cold caches)", ""]
for name in ["gemm", "scan", "ln", "attnqf"]:
    f = os.path.join(O, f"{TAG}_full_{name}_raw.csv")
    if not os.path.exists(f):
        continue
    rows = list(csv.reader(open(f)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for d in data:
        out.append("== " + short(d[idx["Kernel Name"]]))
        for m, label in want:
            if m in idx:
                out.append(f"   {label:24s} {d[idx[m]]:>16s} {units[idx[m]]}   [{m}]")
        out.append("")
        if name == "scan" and "dram__bytes_read.sum" in idx:
            traffic["scan_bytes_per_launch"] = (
                float(d[idx["dram__bytes_read.sum"]].replace(",", "")) * unit_scale(units[idx["dram__bytes_read.sum"]])
                + float(d[idx["dram__bytes_write.sum"]].replace(",", "")) * unit_scale(units[idx["dram__bytes_write.sum"]]))
if len(out) > 2:
    open(os.path.join(P, f"{TAG}_ncu_full_summary.txt"), "w").write("\n".join(out))
if traffic:
    json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
    print(traffic)
