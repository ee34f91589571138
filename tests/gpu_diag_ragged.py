"""Diagnostics (run under gpurun; not collected by pytest): ragged vs padded composed-query fusion per query."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from sprc_b200 import _lib as L  # noqa: E402
from sprc_b200.model import Blip2QformerCirAlignPrompt  # noqa: E402

lib = L.load()
dev = torch.device("cuda:0")
qf = int(sys.argv[1]) if len(sys.argv) > 1 else 2
m = Blip2QformerCirAlignPrompt(vit_model="clip_L", device=dev, max_images=8, max_queries=16, vit_depth=1, qf_layers=qf)
m.load_state_dict(synth.make_state_dict("clip_L", 1, qf, seed=0))
torch.manual_seed(0)


def run(lens, tag):
    B = len(lens)
    ids = torch.zeros(B, 32, dtype=torch.int64)
    mask = torch.zeros(B, 32, dtype=torch.int64)
    for b, n in enumerate(lens):
        ids[b, 0] = 101
        ids[b, 1:n - 1] = torch.randint(1000, 29999, (n - 2,))
        ids[b, n - 1] = 102
        mask[b, :n] = 1
    raws = torch.randn(B, 257, 1024, device=dev).bfloat16()
    ids_d, mask_d = ids.to(dev), mask.to(dev)
    a = torch.empty(B, 256, device=dev)
    b_ = torch.empty(B, 256, device=dev)
    L.check(lib.sprc_encode_query(m._h, L.ptr(raws), L.BF16, None, L.ptr(ids_d), L.ptr(mask_d), B, L.ptr(a), None,
                                  L.cur_stream()))
    lens_t = torch.tensor(lens, dtype=torch.int32)
    L.check(lib.sprc_encode_query_lens(m._h, L.ptr(raws), L.BF16, None, L.ptr(ids_d), L.ptr(lens_t), B, L.ptr(b_), None,
                                       L.cur_stream()))
    torch.cuda.synchronize()
    d = ((a - b_).norm(dim=1) / a.norm(dim=1)).tolist()
    print(f"{tag:28s} lens={lens} per-query rel diff: " + " ".join(f"{x:.1e}" for x in d), flush=True)


run([8], "B=1 L=8")
run([5], "B=1 L=5")
run([8, 8], "B=2 L=8,8")
run([16, 16], "B=2 L=16,16")
run([32, 32], "B=2 full")
run([5, 9], "B=2 L=5,9")
run([9, 5], "B=2 L=9,5")
run([12, 20, 7], "B=3")
run([3, 17, 22, 9, 14, 6, 11, 8], "B=8")
