"""LayerNorm fold, second attempt: correctness + sustained timing of the two GEMM epilogues (csrc/gemm2_fold.cu) at the
query step's shapes, against what they would replace (run under gpurun; not collected by pytest):
    python tests/gpu_probe_fold.py [seconds]
  producer   s' = A W^T + b + LN(s)  (fp32 in place + raw 16-bit copy + statistics)   vs   reduce-add GEMM + LayerNorm kernel
  consumer   act(rstd (s16 Wf^T - mean c) + d)                                         vs   the default GEMM on LN'd rows"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sprc_b200 import _lib as L  # noqa: E402

lib = L.load()
L.check(lib.sprc_set_act_dtype(1))
SECS = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
dev = torch.device("cuda:0")
h16 = torch.float16


def stats(x):   # [M, N] fp32 -> [N/64, M, 2] (mean, M2) partials, part-major
    M, N = x.shape
    xs = x.view(M, N // 64, 64)
    m = xs.mean(-1)
    return torch.stack([m, ((xs - m[..., None]) ** 2).sum(-1)], -1).permute(1, 0, 2).contiguous()


def fold_struct(**kw):
    f = L.SprcGemmFold()
    f.split, f.eps = kw.pop("split", 0), kw.pop("eps", 1e-12)
    keep = []
    for k, t in kw.items():
        if t is not None:
            setattr(f, k, t.data_ptr())
            keep.append(t)
    return f, keep


def sustained(fn, secs=SECS):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    n = max(10, int(secs / 2 / (e0.elapsed_time(e1) / 10 / 1e3)))
    for _ in range(n):
        fn()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3   # us


def check_and_time(M, N, K, normed, name):
    g = torch.Generator(device=dev).manual_seed(M + K + N)
    A = torch.randn(M, K, device=dev, generator=g).to(h16)
    W = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).to(h16)
    b = torch.randn(N, device=dev, generator=g)
    x0 = torch.randn(M, N, device=dev, generator=g) * 1.7 + torch.randn(M, 1, device=dev, generator=g)
    ga, be = 1 + 0.3 * torch.randn(N, device=dev, generator=g), 0.2 * torch.randn(N, device=dev, generator=g)
    st_res = stats(x0)
    # ---- correctness (first 4096 rows in fp32 on the device) ----
    x = x0.clone()
    out16 = torch.zeros(M, N, device=dev, dtype=h16)
    st_out = torch.full((N // 64, M, 2), float("nan"), device=dev)
    f, keep = fold_struct(resid=x, out16=out16, st_out=st_out, st_res=st_res if normed else None,
                          res_g=ga if normed else None, res_b=be if normed else None)
    L.check(lib.sprc_op_gemm_fold(L.ptr(A), L.ptr(W), None, M, 0, N, K, L.ptr(b), None, 0, L.ptr(x), None, f,
                                  L.cur_stream()))
    torch.cuda.synchronize()
    R = min(M, 4096)
    for lo in (0, M - R):
        sl = slice(lo, lo + R)
        r = torch.nn.functional.layer_norm(x0[sl], (N,), ga, be, 1e-12) if normed else x0[sl]
        ref = A[sl].float() @ W.float().T + b + r
        e32 = (x[sl] - ref).abs().max().item()
        e16 = (out16[sl].float() - ref).abs().max().item()
        so = st_out[:, sl].permute(1, 0, 2)
        m = so[:, :, 0].mean(-1)
        var = (so[:, :, 1].sum(-1) + 64 * ((so[:, :, 0] - m[:, None]) ** 2).sum(-1)) / N
        em = (m - ref.mean(-1)).abs().max().item()
        ev = ((var - ref.var(-1, unbiased=False)) / ref.var(-1, unbiased=False)).abs().max().item()
        assert e32 < 2e-3 and e16 < 2e-2 and em < 1e-4 and ev < 1e-3, (name, lo, e32, e16, em, ev)
    # ---- timing ----
    xa = x0.clone()
    xb = torch.empty(M, N, device=dev, dtype=h16)

    def old():
        L.check(lib.sprc_op_gemm(L.ptr(A), L.ptr(W), M, N, K, K, K, 0, 0, L.ptr(b), L.ptr(xa), L.ptr(xa), None, N, 0, 0,
                                 L.cur_stream()))
        L.check(lib.sprc_op_layernorm(L.ptr(xa), M, N, L.ptr(ga), L.ptr(be), 1e-12, 0, 0, L.ptr(xa) if normed else None,
                                      L.ptr(xb), L.cur_stream()))

    def old_gemm_only():
        L.check(lib.sprc_op_gemm(L.ptr(A), L.ptr(W), M, N, K, K, K, 0, 0, L.ptr(b), L.ptr(xa), L.ptr(xa), None, N, 0, 0,
                                 L.cur_stream()))

    def new():
        L.check(lib.sprc_op_gemm_fold(L.ptr(A), L.ptr(W), None, M, 0, N, K, L.ptr(b), None, 0, L.ptr(x), None, f,
                                      L.cur_stream()))

    t_old, t_g, t_new = sustained(old), sustained(old_gemm_only), sustained(new)
    print(f"producer {name:14s} M{M} N{N} K{K}: reduce-add GEMM {t_g:7.1f} us + LayerNorm = {t_old:7.1f} us | "
          f"fold producer {t_new:7.1f} us  ({t_old / t_new:.2f}x)", flush=True)


def consumer(M, N, K, act, name):
    g = torch.Generator(device=dev).manual_seed(M + N + 7)
    s = torch.randn(M, K, device=dev, generator=g) * 1.7 + 0.4 * torch.randn(M, 1, device=dev, generator=g)
    s16 = s.to(h16)
    st = stats(s)
    W = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).to(h16)
    b = torch.randn(N, device=dev, generator=g)
    ga, be = 1 + 0.3 * torch.randn(K, device=dev, generator=g), 0.2 * torch.randn(K, device=dev, generator=g)
    Wf, c, d = torch.empty_like(W), torch.empty(N, device=dev), torch.empty(N, device=dev)
    L.check(lib.sprc_op_fold_weight(L.ptr(W), L.ptr(ga), L.ptr(be), L.ptr(b), N, K, L.ptr(Wf), L.ptr(c), L.ptr(d),
                                    L.cur_stream()))
    out = torch.zeros(M, N, device=dev, dtype=h16)
    f, keep = fold_struct(st_in=st, c=c)

    def new():
        L.check(lib.sprc_op_gemm_fold(L.ptr(s16), L.ptr(Wf), None, M, 0, N, K, L.ptr(d), None, act, None, L.ptr(out), f,
                                      L.cur_stream()))

    new()
    torch.cuda.synchronize()
    R = min(M, 4096)
    sl = slice(M - R, M)
    y = torch.nn.functional.layer_norm(s[sl], (K,), ga, be, 1e-12) @ W.float().T + b
    ref = torch.nn.functional.gelu(y) if act == 1 else (y * torch.sigmoid(1.702 * y) if act == 2 else y)
    err = (out[sl].float() - ref).abs().max().item()
    assert err < 6e-2, (name, err)
    xn = torch.nn.functional.layer_norm(s, (K,), ga, be, 1e-12).to(h16)
    out2 = torch.empty_like(out)

    def old():
        L.check(lib.sprc_op_gemm(L.ptr(xn), L.ptr(W), M, N, K, K, K, 0, 0, L.ptr(b), None, None, L.ptr(out2), N, act, 0,
                                 L.cur_stream()))

    t_old, t_new = sustained(old), sustained(new)
    print(f"consumer {name:14s} M{M} N{N} K{K}: default GEMM {t_old:7.1f} us | fold consumer {t_new:7.1f} us "
          f"({t_old / t_new:.2f}x)   max|err| {err:.2e}", flush=True)


check_and_time(1000, 768, 768, True, "small")          # ragged M, correctness of the edge
check_and_time(112184, 768, 768, True, "qf out")
check_and_time(112184, 768, 3072, True, "qf ffn2")
check_and_time(32896, 1024, 1024, False, "vitL proj")
check_and_time(32896, 1024, 4096, False, "vitL fc2")
check_and_time(32896, 1408, 1408, False, "vitg proj")   # ragged last N block
consumer(112184, 2304, 768, 0, "qf qkv")
consumer(112184, 3072, 768, 1, "qf ffn1 gelu")
consumer(32896, 3072, 1024, 0, "vitL qkv")
consumer(32896, 4096, 1024, 2, "vitL fc1")
