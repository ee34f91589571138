"""GPU checks of the LayerNorm fold (csrc/gemm2_fold.cu, csrc/ln_fold.cu; SPRC_LN_FOLD selects the schedule).

Reference semantics: Qformer.py:291-295, 373-381 (post-LN sublayers  y = LayerNorm(dense(a) + x)) and the ViT's pre-LN
blocks (eva_vit.py:173-176, clip_vit.py:132-139).  Op level: the producer and consumer GEMM epilogues against torch
fp32, including two row ranges with different LayerNorms / weights and ragged edges.  Model level: the parity tests of
tests/test_parity_gpu.py (reference goldens, 1e-3 gates, full-depth Recall@K and rerank) re-run in a child process with
the OTHER schedule than the default one, so both stay pinned whichever is the default.
"""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from sprc_b200 import _lib as L

    so = L.load()
    L.check(so.sprc_set_act_dtype(1))
    return L, so


def _stats(x):
    """[M, N] fp32 -> part-major (mean, M2) partials [N / 64, M, 2]."""
    M, N = x.shape
    xs = x.view(M, N // 64, 64)
    m = xs.mean(-1)
    return torch.stack([m, ((xs - m[..., None]) ** 2).sum(-1)], -1).permute(1, 0, 2).contiguous()


def _merge(st):
    """part-major partials [P, M, 2] -> (mean [M], var [M])."""
    m = st[:, :, 0].mean(0)
    var = (st[:, :, 1].sum(0) + 64 * ((st[:, :, 0] - m[None]) ** 2).sum(0)) / (64 * st.shape[0])
    return m, var


def _fold_struct(L, **kw):
    f = L.SprcGemmFold()
    f.split, f.eps, f.st_stride = kw.pop("split", 0), kw.pop("eps", 1e-12), kw.pop("st_stride", 0)
    keep = []
    for k, t in kw.items():
        if t is not None:
            setattr(f, k, t.data_ptr())
            keep.append(t)
    return f, keep


def test_fold_weight_kernel(lib):
    L, so = lib
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    W = (torch.randn(3072, 768, device=dev, generator=g) * 0.03).half()
    bias = torch.randn(3072, device=dev, generator=g)
    gamma, beta = 1 + 0.3 * torch.randn(768, device=dev, generator=g), 0.2 * torch.randn(768, device=dev, generator=g)
    Wf = torch.empty_like(W)
    c, d = torch.empty(3072, device=dev), torch.empty(3072, device=dev)
    L.check(so.sprc_op_fold_weight(L.ptr(W), L.ptr(gamma), L.ptr(beta), L.ptr(bias), 3072, 768, L.ptr(Wf), L.ptr(c),
                                   L.ptr(d), L.cur_stream()))
    torch.cuda.synchronize()
    wf = (W.float() * gamma).half()
    assert torch.equal(Wf, wf)
    assert (c - wf.float().sum(-1)).abs().max() < 1e-4
    assert (d - (W.float() @ beta + bias)).abs().max() < 1e-4


@pytest.mark.parametrize("M,split,K,dual,normed,N", [(512, 0, 768, False, False, 768), (1000, 0, 768, False, True, 768),
                                                    (27912, 18944, 768, False, True, 768),
                                                    (27912, 18944, 3072, True, True, 768),
                                                    (2570, 0, 4096, False, False, 1024),     # ViT-L fc2, raw residual
                                                    (2570, 0, 1408, False, False, 1408)])    # ViT-g proj, ragged N block
def test_producer_gemm(lib, M, split, K, dual, normed, N):
    """s' = A W^T + b + LN(resid) in place, raw 16-bit copy, N / 64 row-statistics partials (part-major)."""
    L, so = lib
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(M + K)
    A = torch.randn(M, K, device=dev, generator=g).half()
    W1 = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).half()
    W2 = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).half()
    b1, b2 = torch.randn(N, device=dev, generator=g), torch.randn(N, device=dev, generator=g)
    x = torch.randn(M, N, device=dev, generator=g) * 1.7 + torch.randn(M, 1, device=dev, generator=g)
    g1, be1 = 1 + 0.3 * torch.randn(N, device=dev, generator=g), 0.2 * torch.randn(N, device=dev, generator=g)
    g2, be2 = 1 + 0.3 * torch.randn(N, device=dev, generator=g), 0.2 * torch.randn(N, device=dev, generator=g)
    stride = M + 40                                     # statistics planes wider than the launch (as in the model)
    st_res = torch.zeros(N // 64, stride, 2, device=dev)
    st_res[:, :M] = _stats(x)
    hi = torch.arange(M, device=dev) >= split if split else torch.zeros(M, dtype=torch.bool, device=dev)
    r = x
    if normed:
        n1 = torch.nn.functional.layer_norm(x, (N,), g1, be1, 1e-12)
        n2 = torch.nn.functional.layer_norm(x, (N,), g2, be2, 1e-12)
        r = torch.where(hi[:, None], n2, n1)
    y1 = A.float() @ W1.float().T + b1
    y2 = A.float() @ W2.float().T + b2 if dual else y1
    ref = torch.where(hi[:, None], y2, y1) + r
    out16 = torch.zeros(M, N, device=dev, dtype=torch.float16)
    st_a = torch.full((N // 64, stride, 2), float("nan"), device=dev)
    st_b = torch.full((N // 64, stride, 2), float("nan"), device=dev)
    f, keep = _fold_struct(L, split=split, st_stride=stride, resid=x, out16=out16, st_out=st_a,
                           st_out2=st_b if split else None, st_res=st_res if normed else None,
                           st_res2=st_res if (normed and split) else None, res_g=g1 if normed else None,
                           res_b=be1 if normed else None, res_g2=g2 if (normed and split) else None,
                           res_b2=be2 if (normed and split) else None)
    L.check(so.sprc_op_gemm_fold(L.ptr(A), L.ptr(W1), L.ptr(W2) if dual else None, M, split if dual else 0, N, K,
                                 L.ptr(b1), L.ptr(b2) if dual else None, 0, L.ptr(x), None, f, L.cur_stream()))
    torch.cuda.synchronize()
    assert torch.isfinite(x).all()
    assert (x - ref).abs().max().item() < 2e-3
    assert (out16.float() - ref).abs().max().item() < 2e-2
    st = torch.where(hi[None, :, None], st_b[:, :M], st_a[:, :M]) if split else st_a[:, :M]
    m, var = _merge(st)
    assert (m - ref.mean(-1)).abs().max() < 1e-4
    assert ((var - ref.var(-1, unbiased=False)) / ref.var(-1, unbiased=False)).abs().max() < 1e-3


@pytest.mark.parametrize("M,split,N,act,dual,K", [(512, 0, 768, 0, False, 768), (1000, 0, 2304, 0, False, 768),
                                                  (27912, 18944, 3072, 1, True, 768),
                                                  (27912, 18944, 2304, 0, True, 768),
                                                  (2570, 0, 4224, 0, False, 1408),    # ViT-g qkv after norm1
                                                  (2570, 0, 4096, 2, False, 1024)])   # ViT-L fc1 + QuickGELU after ln_2
def test_consumer_gemm(lib, M, split, N, act, dual, K):
    """act(LN(s) W^T + b) from the raw 16-bit rows, the folded weight and the row statistics."""
    L, so = lib
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(M + N)
    s = torch.randn(M, K, device=dev, generator=g) * 1.7 + 0.4 * torch.randn(M, 1, device=dev, generator=g)
    s16 = s.half()
    st = _stats(s)
    refs, keepw = [], []
    for i in range(2 if dual else 1):
        W = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).half()
        b = torch.randn(N, device=dev, generator=g)
        ga, be = 1 + 0.3 * torch.randn(K, device=dev, generator=g), 0.2 * torch.randn(K, device=dev, generator=g)
        Wf, c, d = torch.empty_like(W), torch.empty(N, device=dev), torch.empty(N, device=dev)
        L.check(so.sprc_op_fold_weight(L.ptr(W), L.ptr(ga), L.ptr(be), L.ptr(b), N, K, L.ptr(Wf), L.ptr(c), L.ptr(d),
                                       L.cur_stream()))
        keepw.append((Wf, c, d))
        y = torch.nn.functional.layer_norm(s, (K,), ga, be, 1e-12) @ W.float().T + b
        refs.append(torch.nn.functional.gelu(y) if act == 1 else (y * torch.sigmoid(1.702 * y) if act == 2 else y))
    hi = torch.arange(M, device=dev) >= split if split else torch.zeros(M, dtype=torch.bool, device=dev)
    ref = torch.where(hi[:, None], refs[-1], refs[0])
    out = torch.zeros(M, N, device=dev, dtype=torch.float16)
    f, keep = _fold_struct(L, split=split, st_in=st, st_in2=st if split else None, c=keepw[0][1],
                           c2=keepw[-1][1] if dual else None)
    L.check(so.sprc_op_gemm_fold(L.ptr(s16), L.ptr(keepw[0][0]), L.ptr(keepw[-1][0]) if dual else None, M,
                                 split if dual else 0, N, K, L.ptr(keepw[0][2]),
                                 L.ptr(keepw[-1][2]) if dual else None, act, None, L.ptr(out), f, L.cur_stream()))
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    assert torch.isfinite(out.float()).all() and err < 2e-2, err   # fp16 operands of |values| <~ 8, one output rounding


def test_reference_parity_with_the_other_schedule(lib):
    """Every reference-golden parity test (stage tensors at 1e-3, cir_cat, ragged == padded, batch invariance, full-depth
    Recall@K within +-0.05, full-depth rerank, C1) in a child process with the schedule that is NOT this process's
    default: the fold and the LayerNorm-kernel schedule are both pinned to the reference."""
    L, so = lib
    env = dict(os.environ, SPRC_LN_FOLD="0" if so.sprc_ln_fold_enabled() else "1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_parity_gpu.py"), "-m", "gpu",
                        "-q", "-x", "-k", "not bench_batch"], env=env, cwd=ROOT, capture_output=True, text=True,
                       timeout=1500)
    tail = "\n".join(r.stdout.splitlines()[-15:])
    print(tail)
    assert r.returncode == 0, tail + r.stderr[-2000:]
