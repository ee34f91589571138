"""Diagnostics: ragged Q-Former self-attention kernel vs torch, per sample / row group."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sprc_b200 import _lib as L  # noqa: E402

lib = L.load()
dev = torch.device("cuda:0")
torch.manual_seed(0)


def run(lens):
    B = len(lens)
    toff, T8 = [], 0
    for g in range((B + 1) // 2):
        used = 0
        for b in range(2 * g, min(2 * g + 2, B)):
            toff.append(T8 + used)
            used += lens[b]
        T8 += (used + 7) // 8 * 8
    rows = 32 * B + T8
    qkv = (torch.randn(rows, 2304, device=dev) * 0.5).bfloat16()
    out = torch.full((rows, 768), float("nan"), device=dev).bfloat16()
    pairs = []
    for g in range((B + 1) // 2):
        b0, b1 = 2 * g, 2 * g + 1
        pairs += [toff[b0], lens[b0], toff[b0] + lens[b0], lens[b1] if b1 < B else 0]
    pd = torch.tensor(pairs, dtype=torch.int32, device=dev)
    L.check(lib.sprc_op_attention_ragged(L.ptr(qkv), 2304, L.ptr(out), 768, B, rows, L.ptr(pd), 0.125, L.cur_stream()))
    torch.cuda.synchronize()
    q, k, v = qkv[:, :768].float(), qkv[:, 768:1536].float(), qkv[:, 1536:].float()
    msg = []
    for b in range(B):
        idx = torch.cat([torch.arange(32 * b, 32 * b + 32), torch.arange(32 * B + toff[b], 32 * B + toff[b] + lens[b])]).to(dev)
        qq = q[idx].view(-1, 12, 64).transpose(0, 1)
        kk = k[idx].view(-1, 12, 64).transpose(0, 1)
        vv = v[idx].view(-1, 12, 64).transpose(0, 1)
        ref = (torch.softmax(qq @ kk.transpose(1, 2) * 0.125, -1) @ vv).transpose(0, 1).reshape(-1, 768)
        got = out[idx].float()
        eq = (got[:32] - ref[:32]).abs().max().item()
        et = (got[32:] - ref[32:]).abs().max().item()
        # per head error on the query rows
        eh = (got[:32] - ref[:32]).abs().view(32, 12, 64).amax(dim=(0, 2))
        msg.append(f"b{b}(L={lens[b]}): q {eq:.1e} txt {et:.1e} heads>1e-2: {[i for i, e in enumerate(eh.tolist()) if e > 1e-2]}")
    print(f"lens={lens}: " + " | ".join(msg), flush=True)


run([8])
run([8, 8])
run([32, 32])
run([5, 9])
run([12, 20, 7])
run([3, 17, 22, 9, 14, 6, 11, 8])
