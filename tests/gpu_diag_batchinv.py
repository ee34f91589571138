"""Batch invariance / run-to-run determinism of encode_gallery (run under gpurun; not collected by pytest).

Encodes the same 64 images (a) twice as one batch of 64, (b) as two batches of 32, (c) as 64 + 5 extra images, and
prints how many raws / feats elements differ bitwise.  A nonzero (a) is a race; nonzero (b)/(c) is a kernel whose
arithmetic depends on where an image sits in the batch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sprc_b200 import synth  # noqa: E402
from sprc_b200.model import Blip2QformerCirAlignPrompt  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 2
qf = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
m = Blip2QformerCirAlignPrompt(vit_model="clip_L", device=dev, max_images=96, max_queries=8, vit_depth=depth, qf_layers=qf)
m.load_state_dict(synth.make_state_dict("clip_L", depth, qf, seed=0))
img = torch.randn(69, 3, 224, 224, generator=torch.Generator().manual_seed(5)).clamp_(-2.2, 2.2).to(dev)


def enc(x):
    o = m.encode_gallery(x, want_f32=True, want_bf16=False, want_raws_f32=True)
    torch.cuda.synchronize()
    return o["raws"].clone(), o["feats"].clone()


def diff(tag, a, b):
    for name, x, y in (("raws", a[0], b[0]), ("feats", a[1], b[1])):
        ne = (x != y)
        print(f"{tag:28s} {name:5s}: {int(ne.sum())} of {x.numel()} differ, max|d| {float((x - y).abs().max()):.3e}, "
              f"images touched {sorted(set(ne.flatten(1).any(dim=1).nonzero().flatten().tolist()))[:12]}", flush=True)


a1 = enc(img[:64])
a2 = enc(img[:64])
diff("same batch twice", a1, a2)
b = [enc(img[:32]), enc(img[32:64])]
diff("2 x 32 vs 64", a1, (torch.cat([b[0][0], b[1][0]]), torch.cat([b[0][1], b[1][1]])))
c = enc(img)
diff("first 64 of 69 vs 64", a1, (c[0][:64], c[1][:64]))
d = [enc(img[:33]), enc(img[33:64])]
diff("33 + 31 vs 64", a1, (torch.cat([d[0][0], d[1][0]]), torch.cat([d[0][1], d[1][1]])))
for _ in range(3):
    diff("repeat", a1, enc(img[:64]))
