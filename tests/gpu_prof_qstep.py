"""One composed-query step (Q-Former fusion + text pass) for ncu captures of its kernels:
python tests/gpu_prof_qstep.py [Bq] [iters]   (ViT depth 1: only the query path is exercised)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from sprc_b200 import _lib as L  # noqa: E402
from sprc_b200.model import Blip2QformerCirAlignPrompt  # noqa: E402

Bq = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
lib = L.load()
model = Blip2QformerCirAlignPrompt(vit_model="clip_L", device=dev, max_images=8, max_queries=Bq, vit_depth=1)
sd = synth.make_state_dict("clip_L", 1, 12, seed=0)
model.load_state_dict(sd, strict=False)
N = 2048
raws = torch.randn(N, 257, 1024, device=dev).to(model.act_torch_dtype)
ids, mask = synth.make_token_ids(Bq, seed=1)
lens = mask.sum(dim=1).to(torch.int32).contiguous()   # host: caption lengths for the ragged passes
ids, mask = ids.to(dev), mask.to(dev)
rows = torch.randint(0, N, (Bq,), dtype=torch.int32).to(dev)
fusion = torch.empty(Bq, 256, device=dev, dtype=model.act_torch_dtype)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(iters):
    e0.record()
    L.check(lib.sprc_encode_query_lens(model._h, L.ptr(raws), L.BF16, L.ptr(rows), L.ptr(ids), L.ptr(lens), Bq, None,
                                       L.ptr(fusion), L.cur_stream()))
    e1.record()
    torch.cuda.synchronize()
    print(f"encode_query Bq={Bq}: {e0.elapsed_time(e1):.3f} ms = {Bq / e0.elapsed_time(e1) * 1e3:.0f} q/s")
