"""Sustained (power-capped) throughput of our CTA-pair GEMM against cuBLAS on the SAME shape, one after the other
(run under gpurun; not collected by pytest):  python tests/gpu_sustained_gemm.py [seconds]

A kernel timed alone runs at boost clocks; inside an index batch or a query step the board sits at its power cap and
the SM clock drops to whatever the kernel mix allows.  This loop launches one shape back to back for `seconds`, times
the second half with CUDA events and samples SM clock + board power through NVML meanwhile, so the two libraries are
compared in the regime the product runs in."""
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sprc_b200 import _lib as L  # noqa: E402

lib = L.load()
SECS = float(sys.argv[1]) if len(sys.argv) > 1 else 1.5
dt16 = torch.float16   # the library's default operand format (act_dtype fp16); sprc_op_gemm follows SPRC_ACT_DTYPE
try:
    import pynvml
    pynvml.nvmlInit()
    H = pynvml.nvmlDeviceGetHandleByIndex(0)
except Exception:  # noqa: BLE001
    H = None


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.stop = False
        self.mhz, self.w = [], []

    def run(self):
        while not self.stop and H is not None:
            self.mhz.append(pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM))
            self.w.append(pynvml.nvmlDeviceGetPowerUsage(H) / 1e3)
            time.sleep(0.02)


def sustained(fn, flops):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / 20 / 1e3
    n = max(20, int(SECS / 2 / per))
    for _ in range(n):   # heat-up half
        fn()
    s = Sampler()
    s.start()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    s.stop = True
    s.join()
    t = e0.elapsed_time(e1) / n / 1e3
    med = lambda v: sorted(v)[len(v) // 2] if v else float("nan")  # noqa: E731
    return flops / t / 1e12, t * 1e6, med(s.mhz), med(s.w)


SHAPES = [  # name, M, N, K, act, residual(in place, fp32 out)
    ("vitL qkv", 32896, 3072, 1024, 0, 0),
    ("vitL proj+res", 32896, 1024, 1024, 0, 1),
    ("vitL fc1 qgelu", 32896, 4096, 1024, 2, 0),
    ("vitL fc2+res", 32896, 1024, 4096, 0, 1),
    ("qf qkv", 112184, 2304, 768, 0, 0),
    ("qf out+res", 112184, 768, 768, 0, 1),
    ("qf ffn1 gelu", 112184, 3072, 768, 1, 0),
    ("qf ffn2+res", 112184, 768, 3072, 0, 1),
    ("kv proj", 608576, 9216, 1024, 0, 0),
    ("cublas-ref 8192^3", 8192, 8192, 8192, 0, 0),
]
only = os.environ.get("SPRC_SHAPES")
for name, M, N, K, act, res in SHAPES:
    if only and not any(o in name for o in only.split(",")):
        continue
    A = torch.randn(M, K, device="cuda").to(dt16)
    W = (torch.randn(N, K, device="cuda") * 0.03).to(dt16)
    bias = torch.randn(N, device="cuda")
    flops = 2.0 * M * N * K
    if res:
        out = torch.zeros(M, N, device="cuda")
        ours = lambda: L.check(lib.sprc_op_gemm(L.ptr(A), L.ptr(W), M, N, K, K, K, 0, 0, L.ptr(bias), L.ptr(out),  # noqa: E731
                                                L.ptr(out), None, N, act, 0, L.cur_stream()))
    else:
        out = torch.empty(M, N, device="cuda", dtype=dt16)
        ours = lambda: L.check(lib.sprc_op_gemm(L.ptr(A), L.ptr(W), M, N, K, K, K, 0, 0, L.ptr(bias), None, None,  # noqa: E731
                                                L.ptr(out), N, act, 0, L.cur_stream()))
    co = torch.empty(M, N, device="cuda", dtype=dt16)
    Wt = W.t()
    cub = lambda: torch.matmul(A, Wt, out=co)   # noqa: E731   (no bias / activation / residual: cuBLAS's best case)
    r_c = sustained(cub, flops)
    r_o = sustained(ours, flops)
    r_c2 = sustained(cub, flops)
    print(f"{name:18s} M{M} N{N} K{K}: ours {r_o[0]:7.1f} TF/s ({r_o[1]:7.1f} us, {r_o[2]} MHz, {r_o[3]:.0f} W) | "
          f"cuBLAS plain {r_c[0]:7.1f} / {r_c2[0]:7.1f} TF/s ({r_c[1]:7.1f} us, {r_c[2]} MHz, {r_c[3]:.0f} W)", flush=True)
    del A, W, out, co
    torch.cuda.empty_cache()
