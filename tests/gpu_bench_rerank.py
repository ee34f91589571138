"""Rerank throughput probe (run under gpurun; not collected by pytest): python tests/gpu_bench_rerank.py [R] [T]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from sprc_b200 import _lib as L  # noqa: E402
from sprc_b200.model import Blip2QformerCirRerank  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda:0")
lib = L.load()
m = Blip2QformerCirRerank(vit_model="clip_L", device=dev, max_images=8, max_queries=8, max_pairs=R * T, vit_depth=1)
m.load_state_dict(synth.make_state_dict("clip_L", 1, 12, seed=0))
N = 4096
raws = torch.randn(N, 257, 1024, device=dev).to(m.act_torch_dtype)
ids, mask = synth.make_token_ids(R, seed=1)
ref = torch.randint(0, N, (R,), dtype=torch.int32, device=dev)
cand = torch.randint(0, N, (R * T,), dtype=torch.int32, device=dev)
lib.sprc_profile(1)
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.time()
    p = m.rerank_rows(raws, ref, cand, ids, mask, T)
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(f"rerank R={R} T={T}: {dt * 1e3:.1f} ms = {R * T / dt:.0f} pairs/s = {R / dt:.1f} queries/s "
          f"({R * T * 21.48e9 / dt / 1e12:.0f} TFLOP/s naive, {R * T * 11.8e9 / dt / 1e12:.0f} hoisted)", flush=True)
lib.sprc_profile_dump(b"gpurun_out/rerank_shapes.csv")
