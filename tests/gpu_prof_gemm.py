"""Single-shape GEMM driver for ncu captures: python tests/gpu_prof_gemm.py M N K act res f32 [grp] [iters]"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sprc_b200 import _lib as L
lib = L.load()
M, N, K, act, res, f32 = [int(x) for x in sys.argv[1:7]]
grp = int(sys.argv[7]) if len(sys.argv) > 7 else 0
iters = int(sys.argv[8]) if len(sys.argv) > 8 else 3
rows = M * 2 if grp else M
A = torch.randn(rows, K, device="cuda").bfloat16()
W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
bias = torch.randn(N, device="cuda")
out = torch.zeros(rows, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
R = torch.randn(rows, N, device="cuda") if res else None
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for i in range(iters):
    flush.zero_()
    e0.record()
    L.check(lib.sprc_op_gemm(L.ptr(A), L.ptr(W), M, N, K, K, K, grp, 2 * grp, L.ptr(bias), L.ptr(R),
                             L.ptr(out) if f32 else None, None if f32 else L.ptr(out), N, act, 0, L.cur_stream()))
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(f"M{M} N{N} K{K} act{act} res{res} f32{f32} grp{grp}: {min(ts):.1f} us = {2.0*M*N*K/min(ts)/1e6:.1f} TFLOP/s")
