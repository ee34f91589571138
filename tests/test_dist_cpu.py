"""world_size-2 gloo test (CPU) of the multi-rank plumbing in sprc_b200/retrieval.py: row-shard ranges and
global row offsets, owner-computes routing of the fusion step, the single packed all-gather of per-shard
top-k candidates, the subset-score exchange — everything around the CUDA calls.  The CUDA calls themselves are
replaced by a CPU checker built from the oracle (tests only; the product backend is the CUDA model), so the
sharded result must equal the single-process oracle ranking exactly."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import restatement as R
from sprc_b200 import retrieval as RT


class OracleBackend:
    """CPU stand-in with the methods retrieval.query_topk / rerank_topk use."""
    device = torch.device("cpu")
    max_queries = 8
    max_pairs = 8      # rerank chunk = max_pairs // T queries

    def rerank_rows(self, table, ref_rows, cand_rows, input_ids, attention_mask, T):
        """Stand-in pair score: a function of the VALUES of the reference / candidate raw embeds and the caption only
        (never of where the rows sit in `table`), like inference_rerank."""
        r = table[ref_rows.long()].float().mean(dim=1)                       # [R,Dv]
        c = table[cand_rows.long()].float().mean(dim=1).view(r.shape[0], T, -1)
        txt = (input_ids.float() * attention_mask.float()).sum(dim=1, keepdim=True) / 3e5
        return torch.sigmoid(40 * (r[:, None] * c).sum(-1) + txt).reshape(-1)

    def __init__(self, proj):
        self.proj = proj  # fixed random projection playing the role of the Q-Former fusion

    def encode_query(self, raws_table, input_ids, attention_mask, ref_rows=None, out_dtype=torch.bfloat16):
        ref = raws_table[ref_rows.long()].float().mean(dim=1)          # [Q, Dv]
        txt = (input_ids.float() * attention_mask.float()).sum(dim=1, keepdim=True) / 30000.0
        f = torch.nn.functional.normalize(ref @ self.proj + txt, dim=-1)
        return f.to(torch.bfloat16)

    def sim_topk(self, fusion, feats, k=0, row_offset=0, want_full=False):
        sim = R.similarity(fusion.float(), feats.float())
        n = feats.shape[0]
        order = R.ranking(sim, min(k, n))
        sc = torch.full((fusion.shape[0], k), float("-inf"))
        ix = torch.full((fusion.shape[0], k), -1, dtype=torch.int32)
        sc[:, : order.shape[1]] = torch.gather(sim, 1, order)
        ix[:, : order.shape[1]] = (order + row_offset).int()
        return sc, ix, None

    def gather_scores(self, fusion, feats, rows):
        sim = R.similarity(fusion.float(), feats.float())
        out = torch.gather(sim, 1, rows.clamp_min(0).long())
        out[rows < 0] = float("-inf")
        return out

    def topk_merge(self, cs, ci):
        P, Q, k = cs.shape
        s = cs.permute(1, 0, 2).reshape(Q, P * k)
        i = ci.permute(1, 0, 2).reshape(Q, P * k).long()
        key = torch.where(i < 0, torch.full_like(i, 1 << 40), i)
        o1 = torch.argsort(key, dim=1, stable=True)                    # ties -> lower row
        s, i = torch.gather(s, 1, o1), torch.gather(i, 1, o1)
        o2 = torch.argsort(-s, dim=1, stable=True)[:, :k]
        return torch.gather(s, 1, o2), torch.gather(i, 1, o2).int()


def _problem():
    g = torch.Generator().manual_seed(0)
    N, Q, Dv = 37, 6, 16
    feats = torch.nn.functional.normalize(torch.randn(N, 32, 256, generator=g), dim=-1).to(torch.bfloat16)
    raws = torch.randn(N, 257, Dv, generator=g).to(torch.bfloat16)
    proj = torch.randn(Dv, 256, generator=g)
    ref_rows = torch.tensor([0, 36, 18, 19, 5, 30])
    ids = torch.randint(1000, 30000, (Q, 32), generator=g)
    mask = torch.ones(Q, 32, dtype=torch.long)
    subset = torch.randint(0, N, (Q, 6), generator=g)
    subset[0, 0] = -1
    return N, feats, raws, proj, ref_rows, ids, mask, subset


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, feats, raws, proj, ref_rows, ids, mask, subset = _problem()
        lo, hi = RT.shard_range(N, rank, world)
        names = [f"img{i:03d}" for i in range(N)]
        index = RT.GalleryIndex(feats=feats[lo:hi].contiguous(), raws=raws[lo:hi].contiguous(), names=names, lo=lo,
                                hi=hi, n_total=N)
        sc, ix, sub = RT.query_topk(OracleBackend(proj), index, ref_rows, ids, mask, k=10, subset_rows=subset)
        # rerank of the first 4 candidates: queries split over ranks, raw embeds fetched from their owner ranks
        rr = RT.rerank_topk(OracleBackend(proj), index, ix.long(), ref_rows, ids, mask, 4)
        need = torch.tensor([36, 0, 18, 19, 0, 5][: 3 + 3 * rank])           # ragged, repeated, cross-shard requests
        got = RT.fetch_raw_rows(index, need)
        assert torch.equal(got, raws[need]), "fetch_raw_rows must return the owners' rows in request order"
        assert RT.fetch_raw_rows(index, torch.empty(0, dtype=torch.long)).shape[0] == 0
        torch.save((sc, ix, sub, rr), f"{out_path}.{rank}")
    finally:
        dist.destroy_process_group()


def test_sharded_query_topk_equals_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    N, feats, raws, proj, ref_rows, ids, mask, subset = _problem()
    be = OracleBackend(proj)
    index = RT.GalleryIndex(feats=feats, raws=raws, names=[f"img{i:03d}" for i in range(N)])
    sc1, ix1, sub1 = RT.query_topk(be, index, ref_rows, ids, mask, k=10, subset_rows=subset)
    # single-process result == plain oracle ranking
    fusion = be.encode_query(raws, ids, mask, ref_rows=ref_rows)
    assert torch.equal(ix1.long(), R.ranking(R.similarity(fusion.float(), feats.float()), 10))
    rr1 = RT.rerank_topk(be, index, ix1.long(), ref_rows, ids, mask, 4)
    # brute force, the reference's loop (cirr_test_submission.py:87-112) one query at a time
    for q in range(len(ref_rows)):
        cand = ix1[q, :4].long()
        table = torch.cat([raws[ref_rows[q:q + 1]], raws[cand]])
        p = be.rerank_rows(table, torch.tensor([0]), torch.arange(1, 5), ids[q:q + 1], mask[q:q + 1], 4)
        assert torch.equal(rr1[q, :4], cand[torch.argsort(1 - p, stable=True)]) and torch.equal(rr1[q, 4:], ix1[q, 4:].long())
    assert not torch.equal(rr1, ix1.long())            # the stand-in scores do re-order something
    for r in range(2):
        sc, ix, sub, rr = torch.load(f"{out}.{r}")
        assert torch.equal(ix, ix1) and torch.equal(sc, sc1), f"rank {r}"
        assert torch.equal(sub, sub1)
        assert torch.equal(rr, rr1), f"rank {r}: sharded rerank must equal the single-process rerank"
