"""Is the TMA reduce-add epilogue what paces the residual GEMMs?  Same shape with (a) fp32 reduce-add into the residual
(what the model runs), (b) plain fp32 TMA store, (c) plain 16-bit store; sustained (python tests/gpu_probe_residual.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sprc_b200 import _lib as L  # noqa: E402

lib = L.load()
L.check(lib.sprc_set_act_dtype(1))
dev = torch.device("cuda:0")


def sustained(fn, secs=0.8):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        fn()
    b.record()
    torch.cuda.synchronize()
    n = max(10, int(secs / 2 / (a.elapsed_time(b) / 10 / 1e3)))
    for _ in range(n):
        fn()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for name, M, N, K in (("vitL proj", 32896, 1024, 1024), ("vitL fc2", 32896, 1024, 4096), ("qf out", 112184, 768, 768),
                      ("qf ffn2", 112184, 768, 3072)):
    A = torch.randn(M, K, device=dev).half()
    W = (torch.randn(N, K, device=dev) * 0.03).half()
    bias = torch.randn(N, device=dev)
    x = torch.zeros(M, N, device=dev)
    o32 = torch.empty(M, N, device=dev)
    o16 = torch.empty(M, N, device=dev, dtype=torch.float16)
    red = lambda: L.check(lib.sprc_op_gemm(L.ptr(A), L.ptr(W), M, N, K, K, K, 0, 0, L.ptr(bias), L.ptr(x), L.ptr(x), None,  # noqa: E731
                                           N, 0, 0, L.cur_stream()))
    st32 = lambda: L.check(lib.sprc_op_gemm(L.ptr(A), L.ptr(W), M, N, K, K, K, 0, 0, L.ptr(bias), None, L.ptr(o32), None,  # noqa: E731
                                            N, 0, 0, L.cur_stream()))
    st16 = lambda: L.check(lib.sprc_op_gemm(L.ptr(A), L.ptr(W), M, N, K, K, K, 0, 0, L.ptr(bias), None, None, L.ptr(o16),  # noqa: E731
                                            N, 0, 0, L.cur_stream()))
    t = [sustained(f) for f in (red, st32, st16)]
    fl = 2.0 * M * N * K
    print(f"{name:10s} M{M} N{N} K{K}: reduce-add fp32 {t[0]:7.1f} us ({fl / t[0] / 1e6:6.0f} TF/s) | plain fp32 store "
          f"{t[1]:7.1f} us ({fl / t[1] / 1e6:6.0f}) | plain 16-bit store {t[2]:7.1f} us ({fl / t[2] / 1e6:6.0f})", flush=True)
