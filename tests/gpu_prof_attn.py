"""ViT attention timing, first vs second generation kernel (run under gpurun):
python tests/gpu_prof_attn.py [B] [dh]      SPRC_VIT_ATTN_V1=1 selects the first-generation kernel"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sprc_b200 import _lib as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dh = int(sys.argv[2]) if len(sys.argv) > 2 else 64
H, T = 16, 257
D = H * dh
lib = L.load()
L.check(lib.sprc_set_act_dtype(1))
qkv = torch.randn(B * T, 3 * D, device="cuda").half()
out = torch.zeros(B * T, D, device="cuda", dtype=torch.float16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run():
    L.check(lib.sprc_op_attention(L.ptr(qkv), L.ptr(qkv[:, D:]), L.ptr(qkv[:, 2 * D:]), L.ptr(out), B, H, dh, T, T, 3 * D,
                                  3 * D, 3 * D, D, T, T, None, dh ** -0.5, L.cur_stream()))


for _ in range(3):
    run()
ts = []
for _ in range(10):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run()
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
us = ts[len(ts) // 2] * 1e3
fl = 4.0 * B * H * T * T * dh
by = B * T * 4 * D * 2
print(f"vit attention B={B} dh={dh} {'v1' if os.environ.get('SPRC_VIT_ATTN_V1') else 'v2'}: {us:.1f} us  "
      f"{fl / us / 1e6:.1f} TFLOP/s  {by / us / 1e3:.0f} GB/s (qkv in + out, L2 flushed)")
