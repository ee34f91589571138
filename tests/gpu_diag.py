"""Verbose GPU diagnostics for the single-op kernels (run under gpurun; not collected by pytest).

Prints, per case, the error of each CUDA kernel against a torch fp32 reference computed from the
same bf16-rounded inputs, and on a mismatch an error map by tile so descriptor / layout bugs can be
located from one run.  `python tests/gpu_diag.py [gemm] [ln] [attn] [perf]`
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sprc_b200 import _lib as L  # noqa: E402

lib = L.load()
dev = torch.device("cuda:0")
torch.manual_seed(0)


def gemm(A, W, bias=None, residual=None, out_dtype=torch.float32, act=0, impl=0, grp_rows=0, grp_stride=0,
         out=None, M=None):
    M = A.shape[0] if M is None else M
    N, K = W.shape
    if out is None:
        out = torch.empty(M, N, device=dev, dtype=out_dtype)
    of = out if out.dtype == torch.float32 else None
    ob = out if out.dtype == torch.bfloat16 else None
    L.check(lib.sprc_op_gemm(L.ptr(A), L.ptr(W), M, N, K, A.stride(-2), W.stride(0), grp_rows, grp_stride,
                             L.ptr(bias), L.ptr(residual), L.ptr(of), L.ptr(ob), out.stride(-2), act, impl,
                             L.cur_stream()))
    return out


def report(name, got, ref, tol):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    rel = err.max().item() / denom
    ok = bool(torch.isfinite(got).all().item()) and rel < tol
    print(f"[{'OK ' if ok else 'BAD'}] {name}: max|err|={err.max().item():.3e} rel={rel:.3e} "
          f"ref_absmax={denom:.3e} nan={int((~torch.isfinite(got)).sum().item())}", flush=True)
    if not ok and got.dim() == 2:
        M, N = got.shape
        bm, bn = 32, 32
        e = err[: M // bm * bm, : N // bn * bn]
        if e.numel():
            tiles = e.reshape(M // bm, bm, N // bn, bn).amax(dim=(1, 3))
            bad = (tiles > tol * denom)
            print(f"      bad 32x32 tiles: {int(bad.sum())}/{bad.numel()}; per-row-block counts (first 16): "
                  f"{bad.sum(1)[:16].tolist()}; per-col-block counts (first 16): {bad.sum(0)[:16].tolist()}")
            r0 = err[:8, :8]
            print("      got[0:4,0:8]=", got[:4, :8].tolist())
            print("      ref[0:4,0:8]=", ref[:4, :8].tolist())
            # within-tile pattern of the first bad tile
            idx = bad.nonzero()
            if len(idx):
                i, j = idx[0].tolist()
                sub = err[i * bm:(i + 1) * bm, j * bn:(j + 1) * bn] > tol * denom
                print(f"      first bad tile ({i},{j}): bad rows {sub.any(1).nonzero().flatten().tolist()[:32]} "
                      f"bad cols {sub.any(0).nonzero().flatten().tolist()[:32]}")
    return ok


def run_gemm():
    ok = True
    cases = [
        # (M, N, K, bias, act, residual, out_dtype)
        (128, 128, 64, False, 0, False, torch.float32),
        (128, 128, 128, False, 0, False, torch.float32),
        (128, 256, 256, True, 0, False, torch.float32),
        (256, 128, 768, True, 0, False, torch.bfloat16),
        (300, 384, 1024, True, 1, False, torch.bfloat16),
        (257 * 4, 3072, 1024, True, 0, False, torch.bfloat16),
        (257 * 4, 1024, 4096, True, 0, True, torch.float32),
        (257 * 3, 1408, 1408, True, 2, False, torch.bfloat16),
        (64, 768, 3072, True, 0, True, torch.float32),
        (256 * 2, 1024, 592, False, 0, False, torch.float32),
        (40000, 256, 768, True, 0, False, torch.float32),  # many tiles per CTA (persistent loop, both TMEM stages)
        (19000, 4096, 1024, True, 1, False, torch.bfloat16),  # BN=256 path
    ]
    for (M, N, K, has_bias, act, has_res, odt) in cases:
        A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        bias = torch.randn(N, device=dev) if has_bias else None
        res = torch.randn(M, N, device=dev) if has_res else None
        ref = A.float() @ W.float().t()
        if has_bias:
            ref = ref + bias
        if act == 1:
            ref = torch.nn.functional.gelu(ref)
        elif act == 2:
            ref = ref * torch.sigmoid(1.702 * ref)
        if has_res:
            ref = ref + res
        tol = 2e-2 if odt == torch.bfloat16 else 2e-3
        for impl, nm in ((1, "simt"), (0, "tcgen05")):
            out = gemm(A, W, bias, res, odt, act, impl)
            torch.cuda.synchronize()
            ok &= report(f"gemm[{nm}] M={M} N={N} K={K} bias={has_bias} act={act} res={has_res} {odt}", out, ref,
                         tol)
    # grouped rows: first 32 rows of every 64-row sample (Q-Former query rows), in-place residual
    B = 37
    H = torch.randn(B * 64, 768, device=dev)
    Hb = H.bfloat16()
    W = (torch.randn(768, 768, device=dev) * 0.05).bfloat16()
    bias = torch.randn(768, device=dev)
    for off, nm in ((0, "query rows"), (32, "text rows")):
        ref = H.clone()
        rows = Hb.view(B, 64, 768)[:, off:off + 32].float()
        ref.view(B, 64, 768)[:, off:off + 32] += rows @ W.float().t() + bias
        for impl, inm in ((1, "simt"), (0, "tcgen05")):
            out = H.clone()
            gemm(Hb[off:], W, bias, out[off:], torch.float32, 0, impl, grp_rows=32, grp_stride=64, out=out[off:],
                 M=B * 32)
            torch.cuda.synchronize()
            ok &= report(f"gemm[{inm}] grouped {nm} B={B}", out, ref, 2e-3)
    return ok


def run_ln():
    ok = True
    for rows, width, eps in ((257 * 5, 1408, 1e-6), (257 * 5, 1024, 1e-5), (64 * 7, 768, 1e-12)):
        x = torch.randn(rows, width, device=dev) * 3 + 1
        g = torch.randn(width, device=dev)
        b = torch.randn(width, device=dev)
        ref = torch.nn.functional.layer_norm(x, (width,), g, b, eps)
        of = torch.empty_like(x)
        ob = torch.empty(rows, width, device=dev, dtype=torch.bfloat16)
        L.check(lib.sprc_op_layernorm(L.ptr(x), rows, width, L.ptr(g), L.ptr(b), eps, 0, 0, L.ptr(of), L.ptr(ob),
                                      L.cur_stream()))
        torch.cuda.synchronize()
        ok &= report(f"layernorm f32 {rows}x{width}", of, ref, 1e-5)
        ok &= report(f"layernorm bf16 {rows}x{width}", ob, ref, 1e-2)
    return ok


def attn_ref(q, k, v, scale, mask=None):
    s = torch.einsum("bhqd,bhkd->bhqk", q.float(), k.float()) * scale
    if mask is not None:
        s = s + mask[:, None, None, :]
    p = s.softmax(-1)
    return torch.einsum("bhqk,bhkd->bhqd", p, v.float())


def run_attn():
    ok = True
    # ViT: packed qkv [B*257, 3*D]
    for (B, H, dh) in ((3, 16, 64), (2, 16, 88)):
        D = H * dh
        qkv = (torch.randn(B * 257, 3 * D, device=dev)).bfloat16()
        out = torch.zeros(B * 257, D, device=dev, dtype=torch.bfloat16)
        scale = dh ** -0.5
        L.check(lib.sprc_op_attention(L.ptr(qkv), L.ptr(qkv[:, D:]), L.ptr(qkv[:, 2 * D:]), L.ptr(out), B, H, dh,
                                      257, 257, 3 * D, 3 * D, 3 * D, D, 257, 257, None, scale, L.cur_stream()))
        torch.cuda.synchronize()
        t = qkv.view(B, 257, 3, H, dh).permute(2, 0, 3, 1, 4)
        ref = attn_ref(t[0], t[1], t[2], scale).permute(0, 2, 1, 3).reshape(B * 257, D)
        ok &= report(f"attention ViT B={B} H={H} dh={dh}", out, ref, 2e-2)
    # Q-Former self-attention with pad mask, S=64 and S=32
    for S in (64, 32):
        B, H, dh = 5, 12, 64
        D = 768
        qkv = torch.randn(B * S, 3 * D, device=dev).bfloat16()
        mask = torch.zeros(B, S, device=dev)
        if S == 64:
            for b in range(B):
                mask[b, 32 + 5 + b:] = -10000.0
        out = torch.zeros(B * S, D, device=dev, dtype=torch.bfloat16)
        L.check(lib.sprc_op_attention(L.ptr(qkv), L.ptr(qkv[:, D:]), L.ptr(qkv[:, 2 * D:]), L.ptr(out), B, H, dh, S,
                                      S, 3 * D, 3 * D, 3 * D, D, S, S, L.ptr(mask), 0.125, L.cur_stream()))
        torch.cuda.synchronize()
        t = qkv.view(B, S, 3, H, dh).permute(2, 0, 3, 1, 4)
        ref = attn_ref(t[0], t[1], t[2], 0.125, mask).permute(0, 2, 1, 3).reshape(B * S, D)
        ok &= report(f"attention QF self S={S}", out, ref, 2e-2)
    # Q-Former cross-attention: 32 query rows out of 64-row samples, K/V from a packed [B*257, 6*1536] buffer
    B, H, dh, D = 4, 12, 64, 768
    q = torch.randn(B * 64, D, device=dev).bfloat16()
    kv = torch.randn(B * 257, 6 * 1536, device=dev).bfloat16()
    layer = 3
    out = torch.zeros(B * 64, D, device=dev, dtype=torch.bfloat16)
    kp = kv[:, layer * 1536:]
    vp = kv[:, layer * 1536 + 768:]
    L.check(lib.sprc_op_attention(L.ptr(q), L.ptr(kp), L.ptr(vp), L.ptr(out), B, H, dh, 32, 257, D, 9216, 9216, D,
                                  64, 257, None, 0.125, L.cur_stream()))
    torch.cuda.synchronize()
    qh = q.view(B, 64, H, dh)[:, :32].permute(0, 2, 1, 3)
    kh = kv[:, layer * 1536: layer * 1536 + 768].reshape(B, 257, H, dh).permute(0, 2, 1, 3)
    vh = kv[:, layer * 1536 + 768: layer * 1536 + 1536].reshape(B, 257, H, dh).permute(0, 2, 1, 3)
    ref = attn_ref(qh, kh, vh, 0.125).permute(0, 2, 1, 3).reshape(B, 32, D)
    ok &= report("attention QF cross 32x257", out.view(B, 64, D)[:, :32].reshape(B * 32, D), ref.reshape(B * 32, D),
                 2e-2)
    ok &= report("attention QF cross untouched rows", out.view(B, 64, D)[:, 32:].reshape(B * 32, D),
                 torch.zeros(B * 32, D, device=dev), 1.0)
    return ok


def bench(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run_perf():
    print("--- GEMM throughput (tcgen05 kernel vs torch.matmul/cuBLAS on the same shapes) ---")
    B = 128
    shapes = [
        ("vitL qkv", B * 257, 3072, 1024), ("vitL proj", B * 257, 1024, 1024), ("vitL fc1", B * 257, 4096, 1024),
        ("vitL fc2", B * 257, 1024, 4096), ("vitg qkv", B * 257, 4224, 1408), ("vitg fc1", B * 257, 6144, 1408),
        ("vitg fc2", B * 257, 1408, 6144), ("qf ffn1", 256 * 64, 3072, 768), ("qf out", 256 * 64, 768, 768),
        ("8k cube", 8192, 8192, 8192),
    ]
    for nm, M, N, K in shapes:
        A = torch.randn(M, K, device=dev).bfloat16()
        W = torch.randn(N, K, device=dev).bfloat16()
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        t = bench(lambda: gemm(A, W, None, None, torch.bfloat16, 0, 0, out=out))
        t2 = bench(lambda: torch.matmul(A, W.t(), out=out))
        fl = 2.0 * M * N * K
        print(f"  {nm:10s} M={M} N={N} K={K}: ours {t*1e3:8.1f} us = {fl/t/1e9:7.1f} TFLOP/s | cuBLAS {t2*1e3:8.1f} us "
              f"= {fl/t2/1e9:7.1f} TFLOP/s", flush=True)
    print("--- attention ---")
    for (B, H, dh) in ((128, 16, 64), (128, 16, 88)):
        D = H * dh
        qkv = torch.randn(B * 257, 3 * D, device=dev).bfloat16()
        out = torch.zeros(B * 257, D, device=dev, dtype=torch.bfloat16)
        t = bench(lambda: L.check(lib.sprc_op_attention(L.ptr(qkv), L.ptr(qkv[:, D:]), L.ptr(qkv[:, 2 * D:]),
                                                        L.ptr(out), B, H, dh, 257, 257, 3 * D, 3 * D, 3 * D, D, 257,
                                                        257, None, dh ** -0.5, L.cur_stream())))
        fl = 4.0 * B * H * 257 * 257 * dh
        print(f"  ViT attn B={B} dh={dh}: {t*1e3:8.1f} us = {fl/t/1e9:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "ln", "attn", "perf"]
    print("device:", torch.cuda.get_device_name(0), "| lib:", L.LIB_PATH, flush=True)
    allok = True
    t0 = time.time()
    if "gemm" in which:
        allok &= run_gemm()
    if "ln" in which:
        allok &= run_ln()
    if "attn" in which:
        allok &= run_attn()
    if "perf" in which:
        run_perf()
    print(f"ALL {'OK' if allok else 'BAD'} in {time.time()-t0:.1f}s")
    sys.exit(0 if allok else 1)
