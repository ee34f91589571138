"""Index-build rate against the image batch size (run under gpurun; not collected by pytest):
python tests/gpu_index_sweep.py [vit] [batches...]   e.g.  clip_L 64 128 192 256

Every batch size gets its own model instance (the workspaces are sized by max_images), one warm batch, then
`n_img` images timed with CUDA events; the last size also writes the per-shape profile of ONE batch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sprc_b200 import _lib as L  # noqa: E402
from sprc_b200 import synth  # noqa: E402
from sprc_b200.model import Blip2QformerCirAlignPrompt  # noqa: E402

vit = sys.argv[1] if len(sys.argv) > 1 else "clip_L"
sizes = [int(x) for x in sys.argv[2:]] or [64, 128, 192, 256]
dev = torch.device("cuda:0")
lib = L.load()
FLOPS = {"clip_L": 155.3e9, "eva_clip_g": 520.7e9}[vit]   # per image, ViT + Q-Former gallery pass
PEAK = 1401.2e12
sd = synth.make_state_dict(vit, None, 12, seed=0) if hasattr(synth, "make_state_dict") else None
for IB in sizes:
    m = Blip2QformerCirAlignPrompt(vit_model=vit, device=dev, max_images=IB, max_queries=8)
    m.load_state_dict(sd)
    Dv = m.vit_width
    n_img = max(2048, 8 * IB) // IB * IB
    feats = torch.empty(n_img, 32, 256, device=dev, dtype=m.act_torch_dtype)
    raws = torch.empty(n_img, 257, Dv, device=dev, dtype=m.act_torch_dtype)
    img = torch.randn(IB, 3, 224, 224, device=dev).clamp_(-2.2, 2.2)
    st = L.cur_stream
    L.check(lib.sprc_encode_gallery(m._h, L.ptr(img), IB, None, L.ptr(feats), None, L.ptr(raws), st()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(2):
        e0.record()
        for s in range(0, n_img, IB):
            L.check(lib.sprc_encode_gallery(m._h, L.ptr(img), IB, None, L.ptr(feats[s:]), None, L.ptr(raws[s:]), st()))
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 1e3)
    ips = n_img / best
    print(f"{vit} index batch {IB}: {ips:.0f} img/s  ({best / (n_img / IB) * 1e3:.2f} ms per batch, "
          f"{ips * FLOPS / PEAK:.3f} of the sustained bf16 peak)", flush=True)
    if IB == sizes[-1] or os.environ.get("SPRC_SWEEP_PROFILE_ALL"):
        lib.sprc_profile(1)
        L.check(lib.sprc_encode_gallery(m._h, L.ptr(img), IB, None, L.ptr(feats), None, L.ptr(raws), st()))
        torch.cuda.synchronize()
        lib.sprc_profile_dump(f"gpurun_out/index_sweep_{vit}_b{IB}.csv".encode())
        lib.sprc_profile(0)
    del m, feats, raws, img
    torch.cuda.empty_cache()
