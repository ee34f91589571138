#!/usr/bin/env python
"""Headline benchmark: composed queries/sec at gallery = 50k (BASELINE.json `metric`).

One "step" = one batch of Bq (default 2368 = 16 x 148 SMs) composed queries (reference image row + 32-token caption) through the
hot path: Q-Former fusion (two passes) -> similarity scan over the whole gallery -> top-50.
Workload at N=1: BASELINE.json configs[1] model (ViT-L BLIP-2, full depth, synthetic weights) with the
gallery enlarged to the 50k rows the metric is quoted on; the gallery index (16-bit features + 16-bit raw
embeds) is built by our own ViT/Q-Former from synthetic images before the timed region.  `e2e` starts from caption
STRINGS (C++ tokenizer inside the timed region).  Extra keys: roofline (+ ours-vs-cuBLAS points on the path's shapes),
roofline_scan(_hbm), parity (in-process check against the reference's golden), cpu_baseline (the unmodified reference on
the host cores), eager_gpu, rerank (C5), vit_g (C3 / C4), index_build, index_feed (indexing from PNG files), rank_skew.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  (N > 1: launched by torch.distributed.run, one rank per GPU; gallery rows sharded; per batch one all-gather of the
   query vectors, ONE scan launch of the shard for all ranks' queries, one all-to-all of the per-shard top-k and a merge
   of each rank's own queries over NCCL; per-GPU query work is fixed => weak scaling)

Prints ONE JSON line (rank 0).  Timing: CUDA events on the launching stream, barrier + synchronize on
both sides, max over ranks; inputs of consecutive steps differ and the gallery (819 MB) + weights exceed
the 126 MB L2, so no explicit L2 flush is needed ("l2": "inputs_exceed_l2").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "composed_queries_per_sec_gallery50k"
UNIT = "queries/s"
FLOP_PER_QUERY = {"clip_L": 27.50e9, "eva_clip_g": 29.32e9}   # SURVEY.md §8d (resident reference)
FLOP_PER_IMAGE = {"clip_L": 166.2e9, "eva_clip_g": 533.5e9}   # SURVEY.md §8d (ViT + Q-Former gallery pass)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--vit", default="clip_L", choices=["clip_L", "eva_clip_g"])
    ap.add_argument("--gallery", type=int, default=50000)
    ap.add_argument("--batch", type=int, default=2368,
                    help="composed queries per step per GPU (2368 = 16 x 148 SMs: every Q-Former GEMM is a whole "
                         "number of 256-row pair-tile waves; 592 / 1184 / 2368 measure 37.9k / 38.5k / 39.3k q/s and "
                         "36.5k / 38.1k / 39.1k end to end, profiles/r01o_*)")
    ap.add_argument("--k", type=int, default=50)
    ap.add_argument("--index-batch", type=int, default=128)
    ap.add_argument("--index-images", type=int, default=0,
                    help="encode only this many gallery rows per GPU with the ViT (0 = all); the remaining rows get "
                         "synthetic unit-norm features (profiling runs; query-step kernels are unchanged)")
    ap.add_argument("--cpu-sample", type=int, default=64,
                    help="composed queries per CPU step (reference arm) / per repetition of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rerank", action="store_true", help="skip the C5 rerank row")
    ap.add_argument("--no-vitg", action="store_true", help="skip the EVA-ViT-g rows (C3 / C4 shapes, ViT-g index rate)")
    ap.add_argument("--no-eager-gpu", action="store_true", help="skip the PyTorch-eager-on-this-GPU reference row")
    ap.add_argument("--no-index-feed", action="store_true", help="skip the indexing-from-PNG-files row")
    ap.add_argument("--no-gemm-points", action="store_true",
                    help="skip the sustained ours-vs-cuBLAS GEMM points inside the roofline object")
    ap.add_argument("--act-dtype", default="fp16", choices=["bf16", "fp16"],
                    help="16-bit tensor-core operand format: fp16 (default) = the reference's own autocast precision and "
                         "the mode whose embeddings meet the 1e-3 parity bar; bf16 runs at the same speed "
                         "(profiles/r02a_bench_*.log) with 8x coarser operands")
    ap.add_argument("--profile-dump", default="", help="write per-shape kernel timings (CSV) of the profiling pass")
    return ap.parse_args()


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu captures of this workload
    (profiles/ncu_traffic.json, written by tools/summarize_ncu.py from `ncu --metrics dram__bytes_*` /
    `ncu --set full` runs of the same query step); None when no capture has been summarised."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained",
                                                                               d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.rows = []
        self.proc = None
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            sel = "GPU-" + uuid if not uuid.startswith("GPU-") else uuid
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", sel], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}

    def stop(self):
        if self.proc:
            self.proc.terminate()


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port (torch fp32 restatement of the reference) on host cores
# ---------------------------------------------------------------------------------------------------
def cpu_query_sample(vit, n_queries, gallery_cpu_f32, sd, steps=1, warmup=0, min_seconds=0.0, batch=16):
    """Times fusion (two Q-Former passes) + similarity + full argsort ranking for `n_queries` composed
    queries against the whole gallery, as blip2_qformer_cir_align_prompt.py:312-361 and
    validate_blip.py:253-254 do; reference raw embeds are resident (synthetic LayerNorm-like rows)."""
    from oracle import restatement as R
    from oracle import synth

    torch.set_num_threads(os.cpu_count() or 1)
    Dv = synth.VIT_DIMS[vit][0]
    g = torch.Generator().manual_seed(77)
    times = []
    it = 0
    while it < warmup + steps or sum(times) < min_seconds:
        ref = torch.randn(n_queries, 257, Dv, generator=g)
        ids, mask = synth.make_token_ids(n_queries, seed=1000 + it)
        t0 = time.perf_counter()
        with torch.no_grad():
            # the reference's own batching: FashionIQ loop batch size 16 (validate_blip.py:149-207)
            for b0 in range(0, n_queries, batch):
                sim = R.inference(sd, ref[b0:b0 + batch], gallery_cpu_f32, ids[b0:b0 + batch], mask[b0:b0 + batch])
                order = torch.argsort(1 - sim, dim=-1)  # noqa: F841  (validate_blip.py:253-254)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        it += 1
    return n_queries * len(times) / sum(times), sum(times) / len(times)


def staged_reference_src():
    """Where the UNMODIFIED reference model sources are: /root/reference in the build container, the byte-identical
    staged copies under baseline/_ref/src (tools/stage_reference_scripts.py, written by build()) on the GPU box."""
    for p in ("/root/reference/src", os.path.join(ROOT, "baseline", "_ref", "src")):
        if os.path.isdir(os.path.join(p, "lavis", "models", "blip2_models")):
            return p
    return None


def reference_model(vit, sd, device="cpu"):
    """The reference's own Blip2QformerCirAlignPrompt (lavis/models/blip2_models/blip2_qformer_cir_align_prompt.py),
    loaded through oracle/ref_loader.py (import shims only, no arithmetic replaced), with the synthetic checkpoint and
    transformers' BertTokenizer over the synthetic vocabulary (blip2.py:30-34 adds [DEC])."""
    src = staged_reference_src()
    if src is None:
        return None
    os.environ["SPRC_REFERENCE_SRC"] = src
    import transformers as tr

    from oracle import ref_loader as RL
    from sprc_b200 import synth

    RL.REFERENCE_SRC = src
    model = RL.build_reference_model(vit=vit, seed=0)
    missing = [k for k in model.load_state_dict(sd, strict=False).missing_keys if not k.startswith("Qformer.cls")
               and "position_ids" not in k]
    assert not missing, missing[:5]
    tok = tr.BertTokenizer(vocab={t: i for i, t in enumerate(synth.make_vocab())})
    tok.add_special_tokens({"bos_token": "[DEC]"})
    model.tokenizer = tok
    return model.to(device).eval()


def reference_query_sample(model, vit, n_queries, gallery_f32, steps=1, warmup=0, min_seconds=0.0, device="cpu",
                           autocast=False):
    """One step = what the reference does for one query batch: `model.inference(reference_embeds, target_feats,
    captions)` (tokenisation, two Q-Former passes, its own broadcast matmul + max, align_prompt.py:312-361), then
    `torch.argsort(1 - sim).cpu()` (validate_blip.py:253-254).  Caption strings and reference embeds are synthetic."""
    from sprc_b200 import synth

    Dv = synth.VIT_DIMS[vit][0]
    g = torch.Generator().manual_seed(77)
    vocab = synth.make_vocab()
    times, it = [], 0
    while it < warmup + steps or sum(times) < min_seconds:
        ref = torch.randn(n_queries, 257, Dv, generator=g).to(device)
        caps = synth.make_captions(n_queries, seed=1000 + it, vocab=vocab)
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16, enabled=autocast):
            sim = model.inference(ref, gallery_f32, caps)
            order = torch.argsort(1 - sim.reshape(n_queries, -1).float(), dim=-1).cpu()  # noqa: F841
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        it += 1
    return n_queries * len(times) / sum(times), sum(times) / len(times)


def reference_batch_for_memory(n_gallery, want, avail_bytes):
    """The reference's broadcast matmul materialises Bq copies of the gallery (Bq*N*32*256*4 bytes):
    pick the largest query batch <= `want` whose expansion stays under 30 % of the available memory."""
    per_query = n_gallery * 32 * 256 * 4 * 1.1
    return int(max(1, min(want, (0.30 * avail_bytes) // per_query)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import psutil

    from sprc_b200 import synth

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.make_state_dict(args.vit, None, 12, seed=0)
    gal = synth.make_gallery_features(args.gallery, seed=99)
    model = reference_model(args.vit, sd)
    if model is not None:
        nq = reference_batch_for_memory(args.gallery, min(args.cpu_sample, 16), psutil.virtual_memory().available)
        qps, sec = reference_query_sample(model, args.vit, nq, gal, steps=args.steps, warmup=args.warmup)
        kind = "reference"
        sample = (f"each step = ONE query batch of {nq} composed queries (caption strings + resident reference "
                  f"embeds) through the UNMODIFIED reference Blip2QformerCirAlignPrompt.inference (its tokeniser call, "
                  f"two fp32 Q-Former passes, its own broadcast matmul + max over the {args.gallery}-row fp32 "
                  f"gallery) and torch.argsort(1 - sim).cpu(), torch CPU eager on {cores} threads; sources: "
                  f"{staged_reference_src()} (byte-identical staged copies, oracle/ref_loader.py shims); the "
                  f"reference's loops use batches of 16 (FashionIQ) / 32 (CIRR): the batch is capped so that the "
                  f"Bq-fold gallery expansion of its matmul fits host memory")
    else:
        nq = args.cpu_sample
        qps, sec = cpu_query_sample(args.vit, nq, gal, sd, steps=args.steps, warmup=args.warmup)
        kind = "port"
        sample = (f"{nq} composed queries/step in reference-sized batches of 16, fp32 torch restatement of the "
                  f"reference (oracle port: the staged reference sources were not found), gallery {args.gallery}, "
                  f"similarity as ONE matmul, full argsort ranking, {cores} threads")
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.vit}_blip2_cirr_shape_gallery{args.gallery}", "gallery": args.gallery,
                       "queries_per_step": nq, "k": args.k},
            "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# in-process checks and extra rows of our arm
# ---------------------------------------------------------------------------------------------------
def parity_check(model, dev, index_batch):
    """Parity of THIS process's model (full-depth synthetic checkpoint, the benchmarked operand mode) against the
    reference's own outputs in tests/golden/recall_full_L.pt (oracle/make_golden.py:run_recall_full: the unmodified
    reference class, fp32, 256 structured images, 512 composed queries): embeddings (relative Frobenius), similarity
    (max abs), Recall@K on the labels planted from the reference's ranking (stored in the file), and a bit-exact top-k
    check of the scan kernel on dyadic inputs against torch's stable sort.  Reads a data file; no oracle code."""
    from sprc_b200 import retrieval as RT
    from sprc_b200 import synth

    path = os.path.join(ROOT, "tests", "golden", "recall_full_L.pt")
    if not os.path.exists(path):
        return {"error": "tests/golden/recall_full_L.pt missing"}
    g = torch.load(path)
    c = g["case"]
    if c["vit"] != model.vit_name:
        return {"skipped": f"golden is {c['vit']}, bench model is {model.vit_name}"}
    images = synth.make_structured_images(c["n_images"], seed=c["image_seed"])
    f16, r16, f32, r32 = [], [], [], []
    for s in range(0, c["n_images"], index_batch):
        o = model.encode_gallery(images[s:s + index_batch].to(dev), want_f32=True, want_bf16=True, want_raws_f32=True,
                                 want_raws_bf16=True)
        f16.append(o["feats_bf16"]), r16.append(o["raws_bf16"]), f32.append(o["feats"][:4]), r32.append(o["raws"][:4])
    feats, raws = torch.cat(f16), torch.cat(r16)
    rel = lambda a, b: float((a.float().cpu() - b).norm() / b.norm())  # noqa: E731
    e_feats = rel(f32[0][:4], g["feats_rows"])
    e_raws = rel(r32[0][:4][:, g["raw_rows"]], g["raws_rows"])
    ids, mask, ref = g["input_ids"], g["attention_mask"], g["ref_rows"]
    Q = ids.shape[0]
    fusion = model.encode_query(raws, ids, mask, ref_rows=ref.to(dev))
    sc, ix, full = model.sim_topk(fusion, feats, k=51, want_full=True)
    sub = model.gather_scores(fusion, feats, g["members"].to(torch.int32))
    torch.cuda.synchronize()
    e_sim = float((full.cpu() - g["sim"]).abs().max())
    rec = RT.cirr_recalls_from_topk(ix.cpu(), ref, g["target"], g["members"], sub.cpu())
    rec_ref = tuple(float(x) for x in g["recalls_ref"])
    d_rec = max(abs(a - b) for a, b in zip(rec, rec_ref))
    # scan kernel, integer part: dyadic inputs (exact dot products) -> rows must equal torch's stable descending sort
    gen = torch.Generator().manual_seed(3)
    qd = (torch.randint(-4, 5, (160, 256), generator=gen).float() / 16).to(dev)
    gd = (torch.randint(-4, 5, (3000, 32, 256), generator=gen).float() / 16).to(dev)
    _, ixd, _ = model.sim_topk(qd.to(model.act_torch_dtype), gd.to(model.act_torch_dtype).contiguous(), k=50)
    simd = torch.einsum("qd,ntd->qnt", qd, gd).max(dim=-1).values
    want = torch.argsort(-simd, dim=1, stable=True)[:, :50]
    tol_e, tol_r = 1e-3, 0.05
    return {"golden": "tests/golden/recall_full_L.pt (the reference's own fp32 outputs: full-depth ViT-L + 12-layer "
                      "Q-Former, %d structured images, %d composed queries)" % (c["n_images"], Q),
            "dtype": model.act_dtype, "relF_raws": e_raws, "relF_feats": e_feats, "max_abs_dsim": e_sim,
            "recalls_reference": rec_ref, "recalls_ours": tuple(float(x) for x in rec), "max_recall_delta": d_rec,
            "label_margin": float(g["margin"]), "topk_bit_exact_dyadic": bool(torch.equal(ixd.long(), want)),
            "tolerance": {"embeddings_rel": tol_e, "recall_abs": tol_r},
            "pass": bool(e_raws < tol_e and e_feats < tol_e and e_sim < tol_e and d_rec <= tol_r
                         and torch.equal(ixd.long(), want))}


def rerank_probe(args, sd, raws, feats, dev, pk, clocks, R_=8, T_=100, steps=5):
    """C5: pairs/s of `inference_rerank` for R_ queries x their top-T_ candidates from the resident index (per-image
    cross-attention K/V hoisted out of the pair loop), CUDA events, max over nothing (per GPU; queries split over ranks
    at N > 1 with no data-path collective)."""
    from sprc_b200 import synth
    from sprc_b200.model import Blip2QformerCirRerank

    m = Blip2QformerCirRerank(vit_model=args.vit, device=dev, max_images=8, max_queries=R_, max_pairs=R_ * T_,
                              act_dtype=args.act_dtype, vit_depth=1)
    m.load_state_dict({k_: v for k_, v in sd.items()}, strict=False)
    ids, mask = synth.make_token_ids(R_, seed=99)
    n = raws.shape[0]
    ref = torch.randint(0, n, (R_,), generator=torch.Generator().manual_seed(5)).to(torch.int32).to(dev)
    fusion = m.encode_query(raws, ids, mask, ref_rows=ref)
    _, cand, _ = m.sim_topk(fusion, feats, k=T_)
    cand = cand.reshape(-1).contiguous()
    for _ in range(2):
        p = m.rerank_rows(raws, ref, cand, ids, mask, T_)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    a.record()
    for _ in range(steps):
        p = m.rerank_rows(raws, ref, cand, ids, mask, T_)
    b.record()
    torch.cuda.synchronize()
    t1 = time.time()
    ms = a.elapsed_time(b) / steps
    naive = {"clip_L": 21.48e9, "eva_clip_g": 25.12e9}[args.vit]
    hoisted = 11.8e9 * naive / 25.12e9 if args.vit == "clip_L" else 11.8e9
    pairs_s = R_ * T_ / (ms / 1e3)
    out = {"pairs_per_s_per_gpu": pairs_s, "reranked_queries_per_s_per_gpu": R_ / (ms / 1e3), "R": R_, "T": T_,
           "ms_per_call": ms, "finite": bool(torch.isfinite(p).all()),
           "tflops_reference_count": pairs_s * naive / 1e12, "tflops_executed_hoisted": pairs_s * hoisted / 1e12,
           "frac_of_sustained_peak_executed": pairs_s * hoisted / 1e12 / pk["tf_sust"],
           "clocks": clocks.window(t0, t1),
           "note": "FLOPs per pair: the reference recomputes the K/V projection of all 514 image tokens for every pair "
                   "(SURVEY 8d: %.2f GF); executed work after hoisting per-image K/V ~ %.1f GF" % (naive / 1e9,
                                                                                                  hoisted / 1e9)}
    del m
    torch.cuda.empty_cache()
    return out


def eager_gpu_rows(args, sd, feats, dev):
    """SURVEY 8d "PyTorch eager on the same B200": the UNMODIFIED reference `inference` (+ argsort) with torch's stock
    CUDA kernels, fp32 and under fp16 autocast, one reference-sized batch of 16 queries per step."""
    rows = {}
    try:
        model = reference_model(args.vit, sd, device=str(dev))
        if model is None:
            return {"unavailable": "staged reference sources not found (baseline/_ref/src)"}
        gal = feats.float()
        for name, ac in (("fp32", False), ("fp16_autocast", True)):
            qps, sec = reference_query_sample(model, args.vit, 16, gal, steps=3, warmup=1, device=str(dev), autocast=ac)
            rows[name] = {"value": qps, "unit": UNIT, "s_per_batch_of_16": sec}
        del model, gal
        torch.cuda.empty_cache()
    except Exception as e:  # the row is informative, never fatal
        rows["error"] = f"{type(e).__name__}: {e}"[:300]
    return rows


def gemm_same_shape_points(lib, L, dev, adt, secs=0.5):
    """What the board's power cap allows on THIS box, on the shapes the path runs: our CTA-pair GEMM (bias / activation /
    16-bit conversion / fp32 residual fused) and cuBLAS (plain torch.matmul, no epilogue: its best case) launched back to
    back for `secs` each, timed with CUDA events over the second half.  The denominator of `roofline.frac`
    (MEASURED_PEAKS.json: cuBLAS on 8192^3) is not reachable on these shapes at this cap by either library."""
    shapes = [("vitL_fc1_32896x4096x1024_quickgelu", 32896, 4096, 1024, 2, 0),
              ("qformer_qkv_112184x2304x768", 112184, 2304, 768, 0, 0),
              ("qformer_ffn2_112184x768x3072_fp32_residual", 112184, 768, 3072, 0, 1),
              ("cublas_reference_shape_8192x8192x8192", 8192, 8192, 8192, 0, 0)]
    out = {}

    def sustained(fn, flops):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            fn()
        b.record()
        torch.cuda.synchronize()
        n = max(5, int(secs / 2 / (a.elapsed_time(b) / 5 / 1e3)))
        for _ in range(n):
            fn()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return flops / (a.elapsed_time(b) / n / 1e3) / 1e12

    try:
        for name, M, N, K, act, res in shapes:
            A = torch.randn(M, K, device=dev).to(adt)
            W = (torch.randn(N, K, device=dev) * 0.03).to(adt)
            bias = torch.randn(N, device=dev)
            if res:
                o = torch.zeros(M, N, device=dev)
                ours = lambda: L.check(lib.sprc_op_gemm(L.ptr(A), L.ptr(W), M, N, K, K, K, 0, 0, L.ptr(bias), L.ptr(o),  # noqa: E731
                                                        L.ptr(o), None, N, act, 0, L.cur_stream()))
            else:
                o = torch.empty(M, N, device=dev, dtype=adt)
                ours = lambda: L.check(lib.sprc_op_gemm(L.ptr(A), L.ptr(W), M, N, K, K, K, 0, 0, L.ptr(bias), None, None,  # noqa: E731
                                                        L.ptr(o), N, act, 0, L.cur_stream()))
            co = torch.empty(M, N, device=dev, dtype=adt)
            Wt = W.t()
            cub = lambda: torch.matmul(A, Wt, out=co)  # noqa: E731
            fl = 2.0 * M * N * K
            out[name] = {"ours_tflops": sustained(ours, fl), "cublas_plain_tflops": sustained(cub, fl)}
            del A, W, o, co
            torch.cuda.empty_cache()
        out["how"] = (f"each kernel launched back to back for {secs} s (board at its power cap), CUDA events over the "
                      "second half; ours fuses bias / activation / 16-bit conversion / the fp32 residual, cuBLAS is a "
                      "plain matmul")
    except Exception as e:  # informative, never fatal
        out["error"] = f"{type(e).__name__}: {e}"[:300]
    return out


def index_feed_rows(args, model, dev, n_files=1024, w=640, h=480):
    """SURVEY 8f N2, input side: gallery indexing FROM FILES.  `n_files` synthetic photo-like PNGs (640 x 480 RGB) on
    local disk, indexed (a) through the native feed - file names -> C++ PNG decode into a pinned arena (all host cores)
    -> one H2D copy -> GPU TargetPad / bicubic resize / normalise -> ViT + Q-Former; (b) through the reference's feed -
    `DataLoader(num_workers=2)` whose workers run PIL decode + `targetpad_transform` (utils.py:54-64,
    data_utils.py:91-105) in front of the same encoder.  Plus the two decoders alone (host only)."""
    import shutil
    import tempfile

    import numpy as np
    import PIL.Image
    from torch.utils.data import DataLoader

    from sprc_b200 import retrieval as RT
    from sprc_b200.preprocess import PngBatchDecoder, PngIndexFeeder, TargetPadPreprocessor

    rows = {}
    tmp = tempfile.mkdtemp(prefix="sprc_feed_")
    try:
        rng = np.random.default_rng(0)
        yy, xx = np.mgrid[0:h, 0:w]
        base = []
        for _ in range(16):   # 16 distinct photo-like images, saved n_files / 16 times each under different names
            ch = []
            for _c in range(3):
                a, b, ph = rng.uniform(0.005, 0.08, 3)
                ch.append(np.clip(127 + 80 * np.sin(a * xx + 40 * ph) * np.cos(b * yy) + rng.normal(0, 10, (h, w)), 0, 255))
            base.append(np.stack(ch, -1).astype(np.uint8))
        files = []
        for i in range(n_files):
            f = os.path.join(tmp, f"im{i:05d}.png")
            if i < 16:
                PIL.Image.fromarray(base[i], "RGB").save(f)
            else:
                shutil.copyfile(files[i % 16], f)
            files.append(f)
        mb = sum(os.path.getsize(f) for f in files) / 1e6
        cores = os.cpu_count() or 1

        class Folder:   # classic-mode protocol of the reference's datasets (data_utils.py:253-270)
            def __init__(self, preprocess):
                self.preprocess = preprocess

            def __len__(self):
                return len(files)

            def __getitem__(self, i):
                return os.path.basename(files[i]), self.preprocess(PIL.Image.open(files[i]))

        # decoders alone
        dec = PngBatchDecoder(threads=0)
        for _ in range(2):            # both arenas grow to the batch's size (and get pinned) outside the timed loop
            dec.decode(files[:128])
        t0 = time.perf_counter()
        for s_ in range(0, n_files, 128):
            b = dec.decode(files[s_:s_ + 128])
            assert int(b.status.max()) == 0
        t_nat = time.perf_counter() - t0
        t0 = time.perf_counter()
        for f in files[:64]:
            np.asarray(PIL.Image.open(f).convert("RGB"))
        t_pil = (time.perf_counter() - t0) / 64
        rows["decode_only"] = {"native_images_per_s": n_files / t_nat, "native_threads": min(cores, 32),
                               "pillow_images_per_s_one_core": 1.0 / t_pil}
        # whole feed in front of the encoder
        feeder = PngIndexFeeder(TargetPadPreprocessor(1.25, 224, device=str(dev)), threads=0)
        RT.build_index(Folder(RT.image_path), model, batch_size=128, png_feeder=feeder, keep_raws=False)   # warm
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ia = RT.build_index(Folder(RT.image_path), model, batch_size=128, png_feeder=feeder, keep_raws=False)
        torch.cuda.synchronize()
        t_a = time.perf_counter() - t0
        rows["native_feed"] = {"images_per_s": n_files / t_a,
                               "path": "file names -> sprc_png_decode_files (pinned arena) -> H2D -> "
                                       "sprc_preprocess_targetpad -> sprc_encode_gallery, decode of batch i+1 under the "
                                       "GPU work of batch i"}
        ref_tf = None
        src = staged_reference_src()
        if src:
            sys.path.insert(0, src)
            try:
                import importlib

                du = importlib.import_module("data_utils")
                ref_tf = du.targetpad_transform(1.25, 224)
            except Exception:  # noqa: BLE001
                ref_tf = None
            finally:
                sys.path.remove(src)
        if ref_tf is not None:
            n_ref = min(n_files, 192)
            sub = torch.utils.data.Subset(Folder(ref_tf), range(n_ref))
            t0 = time.perf_counter()
            ib = RT.build_index(sub, model, batch_size=32, num_workers=2, keep_raws=False)   # utils.py:54: batch 32, 2 workers
            torch.cuda.synchronize()
            t_b = time.perf_counter() - t0
            rows["reference_feed"] = {"images_per_s": n_ref / t_b,
                                      "path": "the reference's DataLoader(batch_size=32, num_workers=2) with PIL decode + "
                                              "targetpad_transform in the workers, same encoder"}
            rows["same_index"] = bool(torch.equal(ia.feats[:n_ref], ib.feats))
        rows["files"] = f"{n_files} PNG files {w}x{h} RGB, {mb / n_files:.2f} MB each, local disk (page cache warm)"
        rows["host_cores"] = cores
    except Exception as e:  # informative row, never fatal
        rows["error"] = f"{type(e).__name__}: {e}"[:300]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return rows


def vitg_rows(args, world, rank, dev, pk, dist):
    """BASELINE.json configs[2] / [3] with the EVA-ViT-g model (Dv 1408, 39 blocks): index-build rate, and the query job
    of C3 (N = 1: 6 000 composed queries x 75 000-row gallery, batches of 2 000) or of C4's shape (N > 1: 200 000-row
    gallery sharded row-wise over the ranks, the 6 000 queries split over them, all-to-all of candidates).  Gallery
    features for the scan are synthetic unit-norm rows (SURVEY 8d allows that for scan-side measurements); reference
    raw embeds come from images this model encodes.  Device-resident inputs, CUDA events, max over ranks."""
    from sprc_b200 import _lib as L
    from sprc_b200 import synth
    from sprc_b200.model import Blip2QformerCirAlignPrompt

    lib = L.load()
    k, IB = args.k, 64
    Qtot = 6000
    nb = 3 if world == 1 else 1
    Bq = Qtot // (nb * world)
    n_rows = 75000 if world == 1 else 200000 // world
    lo = 0 if world == 1 else rank * n_rows
    m = Blip2QformerCirAlignPrompt(vit_model="eva_clip_g", device=dev, max_images=IB, max_queries=Bq,
                                   act_dtype=args.act_dtype)
    t0 = time.time()
    sd = synth.make_state_dict("eva_clip_g", None, 12, seed=0)
    assert m.load_state_dict(sd, strict=False).missing_keys == []
    del sd
    t_load = time.time() - t0
    h, adt = m._h, m.act_torch_dtype
    st = lambda: L.c_void_p(torch.cuda.current_stream(dev).cuda_stream)  # noqa: E731
    n_img = 1024
    raws = torch.empty(n_img, 257, 1408, device=dev, dtype=adt)
    ftmp = torch.empty(n_img, 32, 256, device=dev, dtype=adt)
    img = torch.empty(IB, 3, 224, 224, device=dev)
    gen = torch.Generator(device=dev).manual_seed(4242 + rank)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):                      # first pass warms up, second is timed
        a.record()
        for s in range(0, n_img, IB):
            img.normal_(generator=gen).clamp_(-2.2, 2.2)
            L.check(lib.sprc_encode_gallery(h, L.ptr(img), IB, None, L.ptr(ftmp[s:]), None, L.ptr(raws[s:]), st()))
        b.record()
        torch.cuda.synchronize()
    ips = n_img / (a.elapsed_time(b) / 1e3)
    feats = synth.make_gallery_features(n_rows, seed=99 + rank, device=dev, dtype=adt)
    ids, lens, rows = [], [], []
    for j in range(nb):
        i_, m_ = synth.make_token_ids(Bq, seed=777 + 10 * rank + j)
        ids.append(i_.to(dev))
        lens.append(m_.sum(dim=1).to(torch.int32).contiguous())
        rows.append(torch.randint(0, n_img, (Bq,), generator=torch.Generator().manual_seed(j + rank)).to(torch.int32).to(dev))
    fusion = torch.empty(Bq, 256, device=dev, dtype=adt)
    fusion_all = torch.empty(world * Bq, 256, device=dev, dtype=adt)
    sc = torch.empty(Bq, k, device=dev)
    ix = torch.empty(Bq, k, device=dev, dtype=torch.int32)
    send = torch.empty(world, 2, Bq, k, device=dev, dtype=torch.int32)
    recv = torch.empty_like(send)

    def job():
        for j in range(nb):
            L.check(lib.sprc_encode_query_lens(h, L.ptr(raws), L.BF16, L.ptr(rows[j]), L.ptr(ids[j]), L.ptr(lens[j]), Bq,
                                               None, L.ptr(fusion), st()))
            if world == 1:
                L.check(lib.sprc_sim_topk(h, L.ptr(fusion), Bq, L.ptr(feats), n_rows, 0, k, L.ptr(sc), L.ptr(ix), None,
                                          st()))
            else:
                dist.all_gather_into_tensor(fusion_all, fusion)
                L.check(lib.sprc_sim_topk_grouped(h, L.ptr(fusion_all), world * Bq, L.ptr(feats), n_rows, lo, k,
                                                  L.ptr(send[0, 0]), L.ptr(send[0, 1]), Bq, 2 * Bq * k, st()))
                dist.all_to_all_single(recv, send)
                L.check(lib.sprc_topk_merge_packed(h, L.ptr(recv), world, Bq, k, L.ptr(sc), L.ptr(ix), st()))

    job()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    reps = 3
    a.record()
    for _ in range(reps):
        job()
    b.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    if world > 1:
        t = torch.tensor([ms, -ips], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ips = t[0].item(), -t[1].item()
    ok = bool(torch.isfinite(sc).all() and (ix >= 0).all())
    out = {"model": "eva_clip_g (1408 x 39 blocks), synthetic seed 0", "load_s": t_load,
           "index_images_per_s_per_gpu": ips,
           "index_frac_of_vit_gemm_roofline": ips * FLOP_PER_IMAGE["eva_clip_g"] / (pk["tf_sust"] * 1e12),
           "config": ("C3: 6000 queries x 75000-row gallery, 1 GPU" if world == 1 else
                      "C4 shape: 6000 queries x 200000-row gallery sharded row-wise over %d GPUs (%d rows each), "
                      "one all-to-all of candidates" % (world, n_rows)),
           "queries_per_s": Qtot / (ms / 1e3), "ms_per_6000_queries": ms, "results_finite": ok}
    del m, raws, feats, ftmp
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist

    from sprc_b200 import _lib as L
    from sprc_b200 import synth
    from sprc_b200.model import Blip2QformerCirAlignPrompt

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = L.load()
    pk = peaks()

    Bq, k, N = args.batch, args.k, args.gallery
    # caption STRINGS over the generated vocabulary (no bert-base-uncased offline): the library's C++ WordPiece
    # tokenizer maps them back to exactly the ids the device-resident loop uses (checked below)
    import tempfile

    from sprc_b200.tokenizer import OfflineBertTokenizer

    vocab = synth.make_vocab()
    vocab_file = synth.write_vocab(os.path.join(tempfile.mkdtemp(prefix="sprc_vocab_"), "vocab.txt"))
    model = Blip2QformerCirAlignPrompt(vit_model=args.vit, device=dev, max_images=args.index_batch,
                                       max_queries=Bq, act_dtype=args.act_dtype,
                                       tokenizer=OfflineBertTokenizer(vocab_file))
    adt = model.act_torch_dtype
    sd = synth.make_state_dict(args.vit, None, 12, seed=0)
    assert model.load_state_dict(sd, strict=False).missing_keys == []
    Dv = model.vit_width

    # ---- gallery index shard: rows [lo, hi) of the N-row gallery, encoded by our ViT + Q-Former ----
    lo, hi = rank * N // world, (rank + 1) * N // world
    n_local = hi - lo
    feats = torch.empty(n_local, 32, 256, device=dev, dtype=adt)
    raws = torch.empty(n_local, 257, Dv, device=dev, dtype=adt)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    h = model._h
    st = lambda: L.c_void_p(torch.cuda.current_stream(dev).cuda_stream)  # noqa: E731
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    IB = args.index_batch
    img = torch.empty(IB, 3, 224, 224, device=dev)
    # warm one batch, then time the whole index build
    img.normal_(generator=gen).clamp_(-2.2, 2.2)
    L.check(lib.sprc_encode_gallery(h, L.ptr(img), min(IB, n_local), None, L.ptr(feats), None, L.ptr(raws), st()))
    torch.cuda.synchronize()
    launches_index0 = lib.sprc_launch_count()
    n_index = n_local if args.index_images <= 0 else min(n_local, args.index_images)
    if n_index < n_local:
        feats[n_index:] = synth.make_gallery_features(n_local - n_index, seed=99 + rank, device=dev,
                                                      dtype=adt)
        raws[n_index:].zero_()
    e0.record()
    for s in range(0, n_index, IB):
        b = min(IB, n_index - s)
        img.normal_(generator=gen).clamp_(-2.2, 2.2)
        L.check(lib.sprc_encode_gallery(h, L.ptr(img), b, None, L.ptr(feats[s:]), None, L.ptr(raws[s:]), st()))
    e1.record()
    torch.cuda.synchronize()
    index_s = e0.elapsed_time(e1) / 1e3
    index_ips = n_index / index_s
    launches_index = int(lib.sprc_launch_count() - launches_index0)
    # one more batch under the library profiler (per-launch CUDA events): what the tcgen05 GEMMs of an index batch reach
    # on their own, next to the whole-pipeline images/s above
    import ctypes as _ct

    _pr = (_ct.c_double * 20)()
    lib.sprc_profile(1)
    L.check(lib.sprc_encode_gallery(h, L.ptr(img), min(IB, n_local), None, L.ptr(feats), None, L.ptr(raws), st()))
    torch.cuda.synchronize()
    L.check(lib.sprc_profile_read(_pr, 5))
    lib.sprc_profile(0)
    index_prof = {"gemm_ms": _pr[0], "gemm_tflops": (_pr[1] / (_pr[0] / 1e3) / 1e12) if _pr[0] > 0 else 0.0,
                  "attention_ms": _pr[4], "layernorm_ms": _pr[8], "images": min(IB, n_local)}

    # ---- query pool (host pinned + device copies); reference rows come from the local shard ----
    pool = max(4, min(16, args.steps + args.warmup))
    ids_h = torch.empty(pool, Bq, 32, dtype=torch.int64).pin_memory()
    mask_h = torch.empty(pool, Bq, 32, dtype=torch.int64).pin_memory()
    rows_h = torch.empty(pool, Bq, dtype=torch.int32).pin_memory()
    caps_pool = []
    for p_ in range(pool):
        i, m = synth.make_token_ids(Bq, seed=4321 + 100 * rank + p_)
        ids_h[p_], mask_h[p_] = i, m
        caps_pool.append([" ".join(vocab[int(t)] for t in row[1:int(n_) - 1]) for row, n_ in zip(i.tolist(),
                                                                                                  m.sum(dim=1).tolist())])
        rows_h[p_] = torch.randint(0, n_index, (Bq,), generator=torch.Generator().manual_seed(7 + 100 * rank + p_))
    tb = model.tokenizer(caps_pool[0])
    strings_equal_ids = bool(torch.equal(tb.input_ids, ids_h[0]) and torch.equal(tb.attention_mask, mask_h[0]))
    assert strings_equal_ids, "the caption strings must tokenise to the ids of the device-resident loop"
    ids_d, mask_d, rows_d = ids_h.to(dev), mask_h.to(dev), rows_h.to(dev)
    # caption lengths stay on the host, where the tokenizer produced them (sprc_encode_query_lens)
    lens_h = mask_h.sum(dim=2).to(torch.int32).contiguous()
    ragged = os.environ.get("SPRC_RAGGED", "1") != "0"
    out_sc_h = torch.empty(Bq, k, dtype=torch.float32).pin_memory()
    out_ix_h = torch.empty(Bq, k, dtype=torch.int32).pin_memory()
    out_sc_h2 = [torch.empty(Bq, k, dtype=torch.float32).pin_memory() for _ in range(3)]   # pipelined e2e: 2 in flight
    out_ix_h2 = [torch.empty(Bq, k, dtype=torch.int32).pin_memory() for _ in range(3)]
    # end to end FROM STRINGS (what `inference` receives): batch i+1 is tokenised (C++ threads) and enqueued while the
    # GPU works on batch i (two batches in flight through sprc_query_topk_strings_submit / sprc_query_topk_host_wait).
    # SPRC_E2E_IDS=1: the round-1 path from pre-tokenised host ids (serial sprc_query_topk_host)
    e2e_from_ids = os.environ.get("SPRC_E2E_IDS", "0") == "1"
    inflight = [0]
    ids_stage = [torch.empty(Bq, 32, dtype=torch.int64).pin_memory() for _ in range(2)]    # N > 1: tokenizer output
    mask_stage = [torch.empty(Bq, 32, dtype=torch.int64).pin_memory() for _ in range(2)]
    lens_stage = [torch.empty(Bq, dtype=torch.int32) for _ in range(2)]

    fusion = torch.empty(Bq, 256, device=dev, dtype=adt)
    fusion_all = torch.empty(world * Bq, 256, device=dev, dtype=adt)
    sc = torch.empty(Bq, k, device=dev)
    ix = torch.empty(Bq, k, device=dev, dtype=torch.int32)
    # N > 1: ONE exchange buffer [dest rank][scores | rows][Bq][k]; the scan of rank r's queries writes straight into
    # slot r, one all-to-all delivers to every rank the `world` candidate lists of ITS OWN Bq queries, and each rank
    # merges only those (sprc_topk_merge_packed reads the received buffer in place)
    cand_send = torch.empty(world, 2, Bq, k, device=dev, dtype=torch.int32)
    cand_recv = torch.empty(world, 2, Bq, k, device=dev, dtype=torch.int32)
    msc = torch.empty(Bq, k, device=dev)
    mix = torch.empty(Bq, k, device=dev, dtype=torch.int32)
    # N > 1, two streams: the fusion of batch i + 1 (main stream) runs while batch i is exchanged, scanned and merged on
    # a second stream, so the two collectives - and the wait for the slowest rank they imply - leave the critical path.
    # Two buffers of everything a batch owns; SPRC_BENCH_LOCKSTEP=1 selects the one-stream lockstep step.
    pipelined = world > 1 and os.environ.get("SPRC_BENCH_LOCKSTEP", "0") != "1"
    pipe_allowed = pipelined
    pipe = None
    if pipelined:
        pipe = {"n": 0, "s_ex": torch.cuda.Stream(device=dev),
                "fusion": [fusion, torch.empty_like(fusion)], "fusion_all": [fusion_all, torch.empty_like(fusion_all)],
                "send": [cand_send, torch.empty_like(cand_send)], "recv": [cand_recv, torch.empty_like(cand_recv)],
                "msc": [msc, torch.empty_like(msc)], "mix": [mix, torch.empty_like(mix)],
                "ev_enc": [torch.cuda.Event(), torch.cuda.Event()], "ev_gath": [torch.cuda.Event(), torch.cuda.Event()],
                "ev_done": [torch.cuda.Event(), torch.cuda.Event()], "last": 0}

    pipe_state = {"on": pipelined}   # the profiling pass switches to the one-stream step (per-kernel times undisturbed)

    def step_device(i, ids_src=None, lens_src=None, mask_src=None, host_out=None):
        p_ = i % pool
        ids_ = ids_d[p_] if ids_src is None else ids_src
        fus = fusion
        if pipe_state["on"]:
            b_ = pipe["n"] & 1
            pipe["n"] += 1
            pipe["slot"] = b_
            fus = pipe["fusion"][b_]
            if pipe["n"] > 2:   # the all-gather of the batch that last used this buffer has read it
                torch.cuda.current_stream(dev).wait_event(pipe["ev_gath"][b_])
        if ragged:
            L.check(lib.sprc_encode_query_lens(h, L.ptr(raws), L.BF16, L.ptr(rows_d[p_]), L.ptr(ids_),
                                               L.ptr(lens_h[p_] if lens_src is None else lens_src), Bq, None,
                                               L.ptr(fus), st()))
        else:
            L.check(lib.sprc_encode_query(h, L.ptr(raws), L.BF16, L.ptr(rows_d[p_]), L.ptr(ids_),
                                          L.ptr(mask_d[p_] if mask_src is None else mask_src), Bq, None, L.ptr(fus),
                                          st()))
        if world == 1:
            L.check(lib.sprc_sim_topk(h, L.ptr(fusion), Bq, L.ptr(feats), n_local, 0, k, L.ptr(sc), L.ptr(ix), None,
                                      st()))
        elif pipe_state["on"]:
            b_ = pipe["slot"]
            s_enc, s_ex = torch.cuda.current_stream(dev), pipe["s_ex"]
            pipe["ev_enc"][b_].record(s_enc)
            with torch.cuda.stream(s_ex):
                s_ex.wait_event(pipe["ev_enc"][b_])
                dist.all_gather_into_tensor(pipe["fusion_all"][b_], pipe["fusion"][b_])
                pipe["ev_gath"][b_].record(s_ex)     # fusion[b_] may be overwritten by the encode two batches on
                sx = L.c_void_p(s_ex.cuda_stream)
                snd, rcv = pipe["send"][b_], pipe["recv"][b_]
                L.check(lib.sprc_sim_topk_grouped(h, L.ptr(pipe["fusion_all"][b_]), world * Bq, L.ptr(feats), n_local, lo,
                                                  k, L.ptr(snd[0, 0]), L.ptr(snd[0, 1]), Bq, 2 * Bq * k, sx))
                dist.all_to_all_single(rcv, snd)
                L.check(lib.sprc_topk_merge_packed(h, L.ptr(rcv), world, Bq, k, L.ptr(pipe["msc"][b_]),
                                                   L.ptr(pipe["mix"][b_]), sx))
                if host_out is not None:
                    host_out[0].copy_(pipe["msc"][b_], non_blocking=True)
                    host_out[1].copy_(pipe["mix"][b_], non_blocking=True)
                pipe["ev_done"][b_].record(s_ex)
            pipe["last"] = b_
        else:
            dist.all_gather_into_tensor(fusion_all, fusion)
            # ONE scan launch for all world * Bq queries against this rank's shard; the [Bq, k] block of rank r's queries
            # lands in slot r of the exchange buffer (sprc_sim_topk_grouped)
            L.check(lib.sprc_sim_topk_grouped(h, L.ptr(fusion_all), world * Bq, L.ptr(feats), n_local, lo, k,
                                              L.ptr(cand_send[0, 0]), L.ptr(cand_send[0, 1]), Bq, 2 * Bq * k, st()))
            dist.all_to_all_single(cand_recv, cand_send)
            L.check(lib.sprc_topk_merge_packed(h, L.ptr(cand_recv), world, Bq, k, L.ptr(msc), L.ptr(mix), st()))

    ids_dev_stage = torch.empty(Bq, 32, dtype=torch.int64, device=dev)
    mask_dev_stage = torch.empty(Bq, 32, dtype=torch.int64, device=dev)

    # e2e diagnostics: host seconds spent inside submit / wait, and an event pair around every submitted batch (the GPU's
    # busy time per batch and the idle gaps between consecutive batches)
    e2e_diag = {"submit_s": 0.0, "wait_s": 0.0, "ev": []}

    from concurrent.futures import ThreadPoolExecutor

    tok_pool = ThreadPoolExecutor(max_workers=1)
    tok_ahead = {}

    h2d_ev = [None, None]   # pinned staging slot -> the H2D copy that last read it

    def tokenize_step(j):
        if h2d_ev[j & 1] is not None:
            h2d_ev[j & 1].synchronize()
        model.tokenizer.tokenize_into(caps_pool[j % pool], ids_stage[j & 1], mask_stage[j & 1], lens_stage[j & 1])

    def step_host(i):
        p_ = i % pool
        if world == 1 and not e2e_from_ids:
            # strings in: tokenise + enqueue step i (H2D + kernels + D2H), THEN wait for step i-1; every step's
            # tokenisation and copies are inside the timed region
            t0 = time.perf_counter()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            model.query_topk_strings_submit(raws, feats, rows_h[p_], caps_pool[p_], k, out_sc_h2[i % 3], out_ix_h2[i % 3])
            eb.record()
            e2e_diag["ev"].append((ea, eb))
            t1 = time.perf_counter()
            inflight[0] += 1
            if inflight[0] == 2:
                L.check(lib.sprc_query_topk_host_wait(h))
                inflight[0] -= 1
            e2e_diag["submit_s"] += t1 - t0
            e2e_diag["wait_s"] += time.perf_counter() - t1
        elif world == 1:
            L.check(lib.sprc_query_topk_host(h, L.ptr(raws), L.ptr(feats), n_local, L.ptr(rows_h[p_]),
                                             L.ptr(ids_h[p_]), L.ptr(mask_h[p_]), Bq, k, L.ptr(out_sc_h),
                                             L.ptr(out_ix_h), st()))
        else:
            # strings in on every rank: C++ tokenizer into pinned staging, H2D, the device step (all-gather of query
            # vectors, local scans, all-to-all of candidates, merge of this rank's queries), D2H of its top-k
            # The tokenizer works one batch ahead on a helper thread (C++ workers, no interpreter lock): batch i + 1 is
            # tokenised while the GPU runs batch i, as in the single-GPU submit / wait pair.
            s_ = i & 1
            if tok_ahead.get("step") != i:
                tok_ahead["fut"] = tok_pool.submit(tokenize_step, i)
            tok_ahead["fut"].result()
            tok_ahead["step"], tok_ahead["fut"] = i + 1, tok_pool.submit(tokenize_step, i + 1)
            ids_dev_stage.copy_(ids_stage[s_], non_blocking=True)
            if not ragged:
                mask_dev_stage.copy_(mask_stage[s_], non_blocking=True)
            rows_d[p_].copy_(rows_h[p_], non_blocking=True)
            h2d_ev[s_] = torch.cuda.Event()
            h2d_ev[s_].record()
            if pipe_state["on"]:
                # results of batch i land in the host buffers of its slot on the exchange stream; the host waits for
                # batch i - 1 (two batches in flight, as on one GPU), the drain waits for the last one
                step_device(i, ids_src=ids_dev_stage, lens_src=lens_stage[s_], mask_src=mask_dev_stage,
                            host_out=(out_sc_h2[pipe["n"] & 1], out_ix_h2[pipe["n"] & 1]))
                if pipe["n"] > 1:
                    pipe["ev_done"][pipe["last"] ^ 1].synchronize()
            else:
                step_device(i, ids_src=ids_dev_stage, lens_src=lens_stage[s_], mask_src=mask_dev_stage)
                out_sc_h.copy_(msc, non_blocking=True)
                out_ix_h.copy_(mix, non_blocking=True)
                torch.cuda.current_stream(dev).synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def drain_pipe():
        if pipe_state["on"]:   # the exchange stream's last batch is part of the timed work
            torch.cuda.current_stream(dev).wait_stream(pipe["s_ex"])

    def drain_host():
        drain_pipe()
        while inflight[0] > 0:
            L.check(lib.sprc_query_topk_host_wait(h))
            inflight[0] -= 1
        if tok_ahead.get("fut") is not None:   # the batch tokenised ahead of the last step is not used
            tok_ahead["fut"].result()
            tok_ahead.clear()

    def timed(fn, steps, warmup, drain=None):
        for i in range(warmup):
            fn(i)
        if drain:
            drain()
        barrier()
        t_wall0 = time.time()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            fn(warmup + i)
        if drain:
            drain()   # the last step's results are on the host before the clock stops
        b.record()
        barrier()
        t_wall1 = time.time()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, t_wall0, t_wall1

    K, W = args.steps, max(args.warmup, 3)
    clocks = ClockSampler(local)
    time.sleep(0.25)
    l0 = lib.sprc_launch_count()
    ms, tw0, tw1 = timed(step_device, K, W, drain=drain_pipe)
    launches = (lib.sprc_launch_count() - l0) * K // (K + W)
    clk = clocks.window(tw0, tw1)
    value = world * Bq * K / (ms / 1e3)

    ms_e2e, _, _ = timed(step_host, K, W, drain=drain_host)
    e2e_value = world * Bq * K / (ms_e2e / 1e3)
    e2e_host = None
    if e2e_diag["ev"]:
        ev = e2e_diag["ev"][-K:]
        busy = [a_.elapsed_time(b_) for a_, b_ in ev]
        gaps = [ev[j][1].elapsed_time(ev[j + 1][0]) for j in range(len(ev) - 1)]
        n_all = len(e2e_diag["ev"])
        e2e_host = {"host_submit_ms_per_step": e2e_diag["submit_s"] / n_all * 1e3,
                    "host_wait_ms_per_step": e2e_diag["wait_s"] / n_all * 1e3,
                    "gpu_busy_ms_per_step": sum(busy) / len(busy),
                    "gpu_gap_ms_per_step": sum(gaps) / max(1, len(gaps)),
                    "note": "event pair around every submitted batch: busy = first copy to last copy of a batch on the "
                            "GPU, gap = idle time between consecutive batches; host_* = wall time inside submit "
                            "(tokenizer + enqueue) and wait"}
    # the device-resident loop once more AFTER the e2e loop: the GPU is power-capped in this workload, and how much of
    # the value/e2e gap is the host path and how much the power state of a longer run shows in this repeat
    ms_rep, _, _ = timed(step_device, K, W, drain=drain_pipe)
    value_repeat = world * Bq * K / (ms_rep / 1e3)

    # ---- N > 1: the sharded pipeline must give what ONE GPU gives on the whole gallery (bit-exact, query sample) ----
    sharded_equals_single = None
    if world > 1 and N % world == 0:
        step_device(0)
        drain_pipe()
        torch.cuda.synchronize()
        fus_l = pipe["fusion"][pipe["last"]] if pipe_state["on"] else fusion
        mix_l = pipe["mix"][pipe["last"]] if pipe_state["on"] else mix
        msc_l = pipe["msc"][pipe["last"]] if pipe_state["on"] else msc
        whole = torch.empty(N, 32, 256, device=dev, dtype=adt)
        dist.all_gather_into_tensor(whole, feats)
        nq = min(256, Bq)
        sc1 = torch.empty(nq, k, device=dev)
        ix1 = torch.empty(nq, k, device=dev, dtype=torch.int32)
        L.check(lib.sprc_sim_topk(h, L.ptr(fus_l), nq, L.ptr(whole), N, 0, k, L.ptr(sc1), L.ptr(ix1), None, st()))
        torch.cuda.synchronize()
        ok = torch.tensor([int(torch.equal(ix1, mix_l[:nq]) and torch.equal(sc1, msc_l[:nq]))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        sharded_equals_single = bool(ok.item())
        assert sharded_equals_single, "sharded scan + all-to-all + merge differs from the single-GPU scan"
        del whole
        torch.cuda.empty_cache()

    # ---- roofline: a second pass of the same K steps with per-launch CUDA events (library profiler) ----
    import ctypes

    prof = (ctypes.c_double * 20)()
    drain_pipe()
    barrier()
    pipe_state["on"] = False   # one stream: a kernel's CUDA-event time is its own
    lib.sprc_profile(1)
    for i in range(K):
        step_device(W + i)
    torch.cuda.synchronize()
    pipe_state["on"] = pipe_allowed
    L.check(lib.sprc_profile_read(prof, 5))
    if args.profile_dump and rank == 0:
        L.check(lib.sprc_profile_dump((args.profile_dump + ".query.csv").encode()))
        lib.sprc_profile(1)
        L.check(lib.sprc_encode_gallery(h, L.ptr(img), min(IB, n_local), None, L.ptr(feats), None, L.ptr(raws), st()))
        torch.cuda.synchronize()
        L.check(lib.sprc_profile_dump((args.profile_dump + ".index.csv").encode()))
    lib.sprc_profile(0)
    cat = lambda c: dict(ms=prof[c * 4], flops=prof[c * 4 + 1], bytes=prof[c * 4 + 2], n=prof[c * 4 + 3])  # noqa: E731
    gemm, attn, lnorm, scan, merge = (cat(c) for c in range(5))
    prof_total = sum(x["ms"] for x in (gemm, attn, lnorm, scan, merge))
    # N > 1: the step is two collectives long in lockstep, so it is paced by the slowest rank; show every rank's own
    # kernel time per step and SM clock under load (boards of one box differ under their power caps)
    rank_skew = None
    if world > 1:
        mine = torch.tensor([prof_total / max(K, 1), float(clk.get("sm_mhz") or 0.0)], device=dev)
        allr = torch.empty(world, 2, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        rank_skew = {"kernel_ms_per_step": [round(v, 2) for v in allr[:, 0].tolist()],
                     "sm_mhz": [round(v) for v in allr[:, 1].tolist()],
                     "note": "per rank: sum of its own kernels' CUDA-event times per step (profiling pass) and its median "
                             "SM clock in the timed loop; ms_per_step above is the lockstep step (max over ranks)"}
    gemm_tf = gemm["flops"] / (gemm["ms"] / 1e3) / 1e12 if gemm["ms"] > 0 else 0.0
    scan_gbs = scan["bytes"] / (scan["ms"] / 1e3) / 1e9 if scan["ms"] > 0 else 0.0
    scan_tf = scan["flops"] / (scan["ms"] / 1e3) / 1e12 if scan["ms"] > 0 else 0.0
    roofline = {"kernel": "gemm_bf16_tcgen05_2cta_kernel (CTA pairs; only the small last-layer [CLS]-row GEMMs of the "
                          "step run the single-CTA gemm_bf16_tcgen05_kernel)", "bound": "tensor", "achieved": gemm_tf,
                "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": gemm_tf / pk["tf_sust"],
                "traffic": ncu_traffic().get("gemm_bytes_per_launch"),
                "traffic_source": ncu_traffic().get("gemm_source"),
                "peak_source": pk["src"] + ", sustained bf16 (kernel timed inside a long step)",
                "launches_per_step": gemm["n"] / K, "avg_launch_us": gemm["ms"] * 1e3 / max(gemm["n"], 1),
                "share_of_step": gemm["ms"] / prof_total if prof_total else None,
                "how": "algorithmic 2*M*N*K per launch / CUDA-event time per launch, second pass of the same steps",
                "same_shapes_at_the_power_cap": (gemm_same_shape_points(lib, L, dev, adt)
                                                 if (rank == 0 and world == 1 and not args.no_gemm_points) else None)}
    roofline_scan = {"kernel": "scan_topk_kernel", "bound": "hbm" if world * Bq <= 128 else "tensor",
                     "achieved_gbs": scan_gbs, "peak_gbs": pk["hbm"], "frac_hbm": scan_gbs / pk["hbm"],
                     "achieved_tflops": scan_tf, "frac_tensor": scan_tf / pk["tf_burst"],
                     "avg_launch_us": scan["ms"] * 1e3 / max(scan["n"], 1),
                     "share_of_step": scan["ms"] / prof_total if prof_total else None,
                     "note": "gallery bytes N*32*256*2 per launch; queries/launch = %d" % (world * Bq)}
    # ---- scan in its HBM-bound regime (one 128-query tile per gallery pass: every gallery byte is read once) ----
    # a kernel timed ALONE (MEASURED_PEAKS' burst regime): let the power state of the long loops above settle first
    torch.cuda.synchronize()
    time.sleep(1.0)
    q128 = torch.nn.functional.normalize(torch.randn(128, 256, device=dev), dim=-1).to(adt)
    sc128 = torch.empty(128, k, device=dev)
    ix128 = torch.empty(128, k, device=dev, dtype=torch.int32)
    sa, sb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib.sprc_profile(1)
    for i in range(W + K):
        if i == W:
            lib.sprc_profile(1)
        L.check(lib.sprc_sim_topk(h, L.ptr(q128), 128, L.ptr(feats), n_local, lo, k, L.ptr(sc128), L.ptr(ix128), None,
                                  st()))
    torch.cuda.synchronize()
    prof2 = (ctypes.c_double * 20)()
    L.check(lib.sprc_profile_read(prof2, 5))
    lib.sprc_profile(0)
    s_ms, s_bytes, s_n = prof2[3 * 4], prof2[3 * 4 + 2], prof2[3 * 4 + 3]
    hbm_gbs = s_bytes / (s_ms / 1e3) / 1e9 if s_ms > 0 else 0.0
    roofline_scan_hbm = {"kernel": "scan_topk_kernel", "bound": "hbm", "achieved": hbm_gbs, "peak": pk["hbm"],
                         "unit": "GB/s", "frac": hbm_gbs / pk["hbm"],
                         "traffic": ncu_traffic().get("scan_bytes_per_launch"),
                         "avg_launch_us": s_ms * 1e3 / max(s_n, 1),
                         "note": "128 queries (one UMMA M tile) x the %d-row gallery shard, top-%d; gallery "
                                 "(%.0f MB) exceeds L2; algorithmic bytes N*32*256*2 + Q*512 per launch" % (
                                     n_local, k, n_local * 32 * 256 * 2 / 1e6)}

    breakdown = {"gemm_ms": gemm["ms"] / K, "attention_ms": attn["ms"] / K, "layernorm_ms": lnorm["ms"] / K,
                 "scan_ms": scan["ms"] / K, "merge_ms": merge["ms"] / K}

    # ---- parity of this very model against the reference's golden outputs; rerank (C5) row ----
    parity = parity_check(model, dev, args.index_batch) if rank == 0 else None
    rerank = None
    if not args.no_rerank:
        rerank = rerank_probe(args, sd, raws, feats, dev, pk, clocks)
        if world > 1:
            t = torch.tensor([rerank["pairs_per_s_per_gpu"]], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            rerank["pairs_per_s_all_gpus"] = world * t.item()   # slowest rank x ranks: queries split contiguously
    index_feed = None
    if rank == 0 and world == 1 and not args.no_index_feed:
        index_feed = index_feed_rows(args, model, dev)
    vitg = None
    if not args.no_vitg and args.vit == "clip_L":
        del cand_send, cand_recv
        torch.cuda.empty_cache()
        vitg = vitg_rows(args, world, rank, dev, pk, dist)
    clocks.stop()

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload on the host cores ----
    cpu = None
    eager = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import psutil

        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        gal_cpu = feats.float().cpu()
        t_cpu0 = time.perf_counter()
        ref_model = reference_model(args.vit, sd)
        if ref_model is not None:
            nq = reference_batch_for_memory(N, 16, psutil.virtual_memory().available)
            qps, sec = reference_query_sample(ref_model, args.vit, nq, gal_cpu, steps=1, warmup=1, min_seconds=12.0)
            cpu = {"value": qps, "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": f"repetitions of ONE reference-sized batch of {nq} composed queries (caption strings) "
                             f"through the UNMODIFIED reference Blip2QformerCirAlignPrompt.inference (tokeniser, two "
                             f"fp32 Q-Former passes, its broadcast matmul + max) + argsort(1 - sim).cpu() against the "
                             f"same {N}-row gallery (our index as fp32) for >= 12 s after one warm-up "
                             f"({time.perf_counter() - t_cpu0:.1f} s in all incl. model construction), torch CPU "
                             f"eager on {cores} host threads, {sec:.2f} s per batch"}
            # the indexing half of the path on the same host cores: the unmodified reference's
            # extract_target_features (ViT + Q-Former gallery pass, fp32) on one batch of 8 synthetic images
            try:
                imgs = synth.make_images(8)
                t_i0 = time.perf_counter()
                with torch.no_grad():
                    ref_model.extract_target_features(imgs, mode="mean")
                cpu["index_images_per_s"] = 8 / (time.perf_counter() - t_i0)
                cpu["index_sample"] = ("ONE batch of 8 synthetic images through the unmodified reference "
                                       "extract_target_features (no warm-up), same host threads")
            except Exception as e:  # informative
                cpu["index_error"] = f"{type(e).__name__}: {e}"[:200]
            del ref_model
        else:
            qps, sec = cpu_query_sample(args.vit, args.cpu_sample, gal_cpu, sd, steps=1, warmup=1, min_seconds=12.0)
            cpu = {"value": qps, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"repetitions of {args.cpu_sample} composed queries (batch 16) vs the same {N}-row gallery "
                             f"for >= 12 s, fp32 torch restatement (oracle port; staged reference sources not found) "
                             f"on {cores} host threads, {sec:.2f} s per repetition"}
        del gal_cpu
        if not args.no_eager_gpu:
            eager = eager_gpu_rows(args, sd, feats, dev)

    if world == 1 and ragged:
        # what sprc_query_topk_host* copies: ids int64 [Bq,32], ref rows int32 [Bq], the ragged row tables built from
        # the host mask (toff | len | cls [Bq] each, row->sample [T8], pair table int4 [Bq/2]); the mask stays on the host
        lp = lens_h[0].clamp_min(1).view(-1, 2).sum(dim=1) if Bq % 2 == 0 else lens_h[0].clamp_min(1)
        t8 = int(((lp + 7) // 8 * 8).sum())
        h2d_bytes = Bq * 32 * 8 + Bq * 4 + 4 * (3 * Bq + t8 + 4 * ((Bq + 1) // 2))
        e2e_api = ("sprc_query_topk_host (pinned host ids/mask/ref rows -> top-k on host)" if e2e_from_ids else
                   "model.query_topk_strings_submit / query_topk_host_wait = sprc_query_topk_strings_submit + "
                   "sprc_query_topk_host_wait: caption STRINGS + host reference rows in (C++ WordPiece tokenizer "
                   "inside the timed region), top-k on the host out; two batches in flight so batch i+1 is tokenised "
                   "and enqueued while the GPU works on batch i")
    else:
        h2d_bytes = world * (Bq * (32 * 8 * (1 if ragged else 2) + 4))
        e2e_api = ("caption STRINGS -> C++ tokenizer (one batch ahead, helper thread) -> pinned ids -> device step (all-gather of query vectors, local "
                   "scan in ONE grouped launch, ONE all-to-all of candidates, merge of this rank's queries) -> top-k rows "
                   "of this rank on host; two batches in flight (the fusion of batch i + 1 on the main stream while batch i "
                   "is exchanged, scanned and merged on a second stream)" + ("" if pipelined else " - disabled: lockstep"))
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.act_dtype, "data": "synthetic",
            "config": {"workload": f"{args.vit}_blip2_cirr_shape_gallery{N}", "gallery": N,
                       "queries_per_step_per_gpu": Bq, "k": k, "gallery_rows_per_gpu": n_local,
                       "l2": "inputs_exceed_l2 (gallery %.0f MB + weights; query batches rotate)" % (
                           n_local * 32 * 256 * 2 / 1e6),
                       "parallelism": "gallery rows sharded x%d, queries data-parallel" % world + (
                           "; two-stream step (fusion of batch i + 1 overlaps the exchange / scan / merge of batch i)"
                           if pipelined else ""),
                       "weights": "synthetic seed 0, full depth (sprc_b200/synth.py)",
                       "captions": "32-token rows, %.1f live tokens on average (SURVEY 8d: L~U{3..20} + [CLS],[SEP]); "
                                   "%s" % (float(lens_h.float().mean()),
                                           "query passes over live text rows only (ragged layout)" if ragged
                                           else "all 64 padded rows per query computed"),
                       "text": "e2e starts from caption strings over a generated 30 522-token vocabulary "
                               "(strings_equal_ids: %s)" % strings_equal_ids},
            "clocks": clk,
            "value_repeat_after_e2e": value_repeat,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": world * Bq * k * 8, "ms_per_step": ms_e2e / K,
                    "api": e2e_api, "pipeline": e2e_host},
            "gpu_launches": int(launches),
            "roofline": roofline, "roofline_scan": roofline_scan, "roofline_scan_hbm": roofline_scan_hbm,
            "step_breakdown_ms": breakdown,
            # queries/s x the REFERENCE's FLOPs per composed query (all 64 padded rows) / sustained peak: an
            # "effective" figure - the ragged passes execute fewer FLOPs than that; roofline.achieved counts executed ones
            "frac_of_qformer_gemm_roofline": value / world * FLOP_PER_QUERY[args.vit] / (pk["tf_sust"] * 1e12),
            "index_build": {"images_per_s_per_gpu": index_ips, "seconds": index_s,
                            "frac_of_vit_gemm_roofline": index_ips * FLOP_PER_IMAGE[args.vit] / (pk["tf_sust"] * 1e12),
                            "images_encoded_per_gpu": n_index, "launches": launches_index,
                            "gemm_kernels": {"achieved_tflops": index_prof["gemm_tflops"],
                                             "frac_of_sustained_peak": index_prof["gemm_tflops"] / pk["tf_sust"],
                                             "ms_per_batch": index_prof["gemm_ms"],
                                             "attention_ms_per_batch": index_prof["attention_ms"],
                                             "layernorm_ms_per_batch": index_prof["layernorm_ms"],
                                             "batch_images": index_prof["images"],
                                             "how": "one index batch under the library profiler: algorithmic 2*M*N*K of "
                                                    "every tcgen05 GEMM launch / its CUDA-event time (the whole-pipeline "
                                                    "fraction above also pays for attention, LayerNorm and the "
                                                    "fp32-residual epilogues)"}},
            "cpu_baseline": cpu,
            "parity": parity,
            "rerank": rerank,
            "eager_gpu": eager,
            "index_feed": index_feed,
            "sharded_equals_single": sharded_equals_single,
            "rank_skew": rank_skew,
            "roofline_vit": {args.vit: {"images_per_s_per_gpu": index_ips,
                                        "frac_of_vit_gemm_roofline": index_ips * FLOP_PER_IMAGE[args.vit] / (
                                            pk["tf_sust"] * 1e12)},
                             **({"eva_clip_g": {"images_per_s_per_gpu": vitg["index_images_per_s_per_gpu"],
                                                "frac_of_vit_gemm_roofline": vitg["index_frac_of_vit_gemm_roofline"]}}
                                if vitg else {}),
                             "note": "images/s x the reference's FLOPs per image (ViT + Q-Former gallery pass, SURVEY "
                                     "8d) / sustained bf16 peak"},
            "vit_g": vitg,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
