"""Single-shape ViT attention driver for ncu captures: python tests/gpu_prof_attn.py B dh [iters]"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sprc_b200 import _lib as L
lib = L.load()
B, dh = int(sys.argv[1]), int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
H = 16
D = H * dh
qkv = torch.randn(B * 257, 3 * D, device="cuda").bfloat16()
out = torch.zeros(B * 257, D, device="cuda", dtype=torch.bfloat16)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for i in range(iters):
    e0.record()
    L.check(lib.sprc_op_attention(L.ptr(qkv), L.ptr(qkv[:, D:]), L.ptr(qkv[:, 2 * D:]), L.ptr(out), B, H, dh, 257, 257,
                                  3 * D, 3 * D, 3 * D, D, 257, 257, None, dh ** -0.5, L.cur_stream()))
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(f"ViT attention B={B} dh={dh}: {min(ts):.1f} us = {4.0*B*H*257*257*dh/min(ts)/1e6:.1f} TFLOP/s")
