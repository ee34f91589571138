"""Host-side retrieval drivers: the reference's eval loops (src/utils.py:46-77, src/validate_blip.py:24-57,
149-207,232-285,359-410, src/cirr_test_submission.py:61-132) re-expressed over integer gallery rows and the
fused CUDA scan/top-k, with optional row-sharding of the gallery over ranks.

What changes relative to the reference (results are identical; see tests/test_dropin_gpu.py):
  * the gallery index keeps bf16 features [N,32,256] and bf16 raw embeds [N,257,Dv] (half of the
    reference's fp32 residency);
  * queries are ranked with top-(k+1) + the 6 subset scores instead of a full argsort of N scores and
    O(Q*N) string comparisons (validate_blip.py:253-271); labels are integer row ids;
  * with world_size > 1 each rank indexes and scans rows [lo, hi); per-shard top-k candidates are
    exchanged with ONE all-gather (scores+ids) and merged; query vectors are combined with one all-reduce.

`backend` is the CUDA model (sprc_b200.model.Blip2QformerCirAlignPrompt).  Tests may inject a CPU
checker with the same five methods; the product never does.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row shard of rank `rank`: [rank*n//world, (rank+1)*n//world)."""
    return rank * n // world, (rank + 1) * n // world


def owner_of(row: torch.Tensor, n: int, world: int) -> torch.Tensor:
    """Rank that owns global gallery row(s) `row` under `shard_range`."""
    # smallest r with (r+1)*n//world > row
    his = torch.tensor([(r + 1) * n // world for r in range(world)], dtype=torch.int64)
    return torch.bucketize(row.to(torch.int64), his, right=True).clamp_(0, world - 1)


@dataclass
class GalleryIndex:
    """Row shard [lo, hi) of an N-row gallery.  `names` always lists ALL N rows (names are tiny)."""
    feats: torch.Tensor               # bf16 [hi-lo, 32, 256]
    raws: Optional[torch.Tensor]      # bf16 [hi-lo, 257, Dv] (None if raw embeds were not kept)
    names: List[str]
    lo: int = 0
    hi: int = 0
    n_total: int = 0
    name_to_row: Dict[str, int] = field(default_factory=dict)
    bounds: Optional[List[int]] = None   # world+1 row boundaries of ALL ranks' shards (None: `shard_range` split)

    def __post_init__(self):
        if not self.n_total:
            self.n_total = len(self.names)
        if not self.hi:
            self.hi = self.lo + self.feats.shape[0]
        if not self.name_to_row:
            self.name_to_row = {n: i for i, n in enumerate(self.names)}

    def owner(self, rows: torch.Tensor, world: int) -> torch.Tensor:
        """Rank holding each global row: by the recorded shard boundaries when the index was built from a dataset
        that dropped unreadable images (shards then differ from `shard_range`), else by `shard_range`."""
        if self.bounds is None:
            return owner_of(rows, self.n_total, world)
        his = torch.tensor(self.bounds[1:], dtype=torch.int64)
        return torch.bucketize(rows.to(torch.int64), his, right=True).clamp_(0, world - 1)

    def rows_of(self, names: Sequence[str]) -> torch.Tensor:
        return torch.tensor([self.name_to_row[n] for n in names], dtype=torch.int64)


def _dist():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


# ---------------------------------------------------------------------------------------------------
# indexing  (utils.py:46-77)
# ---------------------------------------------------------------------------------------------------
@torch.no_grad()
def raw_rgb(image):
    """`preprocess` callable for the reference's dataset classes (data_utils.py CIRRDataset / FashionIQDataset take
    any callable): hands the decoded image to `build_index(..., gpu_preprocess=...)` as an RGB uint8 array instead
    of running the PIL transform chain in the DataLoader worker."""
    if image.mode != "RGB":
        raise NotImplementedError(f"image mode {image.mode!r}: the GPU preprocessor takes RGB images")
    return np.asarray(image)


def image_path(image):
    """`preprocess` callable for the reference's dataset classes that hands back the file NAME: `PIL.Image.open` is lazy
    (only the header has been read when the dataset calls `preprocess`), so `build_index(..., png_feeder=...)` can decode
    the whole batch with the native threaded decoder (csrc/png.cpp) instead of one PIL image per DataLoader worker."""
    return image.filename


def _path_batches(dataset, batch_size):
    names, paths = [], []
    for i in range(len(dataset)):
        item = dataset[i]
        if item is None:          # the reference's datasets return None for files that fail to open
            continue
        names.append(item[0])
        paths.append(item[1])
        if len(names) == batch_size:
            yield names, paths
            names, paths = [], []
    if names:
        yield names, paths


def _collate_raw(batch):
    batch = [b for b in batch if b is not None]      # utils.py:141-148 drops unreadable images
    return [b[0] for b in batch], [b[1] for b in batch]


def build_index(dataset, backend, batch_size: int = 64, num_workers: int = 2, collate_fn=None,
                keep_raws: bool = True, progress: bool = False, gpu_preprocess=None, png_feeder=None) -> GalleryIndex:
    """Encode this rank's row shard of `dataset` ('classic' mode: items are (name, image)).
    gpu_preprocess: a `sprc_b200.preprocess.TargetPadPreprocessor`; the dataset must then be built with
    `preprocess=retrieval.raw_rgb` so that items are (name, uint8 [H,W,3]) and resize/crop/normalise run on the GPU
     — workers only decode.
    png_feeder: a `sprc_b200.preprocess.PngIndexFeeder`; the dataset must then be built with
    `preprocess=retrieval.image_path` so that items are (name, file name): no DataLoader workers at all — a helper
    thread decodes batch i+1 natively (C++ workers, pinned arena) while the GPU resizes and encodes batch i."""
    from torch.utils.data import DataLoader, Subset

    if gpu_preprocess is not None:
        collate_fn = _collate_raw

    dist, rank, world = _dist()
    n = len(dataset)
    lo, hi = shard_range(n, rank, world)
    sub = Subset(dataset, range(lo, hi)) if world > 1 else dataset
    feats, raws, names = [], [], []
    if png_feeder is not None:
        from concurrent.futures import ThreadPoolExecutor

        def _fed():
            with ThreadPoolExecutor(max_workers=1) as ex:
                pending = None
                for bn, bp in _path_batches(sub, batch_size):
                    fut = ex.submit(png_feeder.decode, bp)
                    if pending is not None:
                        yield pending[0], pending[1].result()
                    pending = (bn, fut)
                if pending is not None:
                    yield pending[0], pending[1].result()

        it = _fed()
    else:
        it = DataLoader(dataset=sub, batch_size=batch_size, num_workers=num_workers,
                        pin_memory=gpu_preprocess is None, collate_fn=collate_fn)
    if progress:
        from tqdm import tqdm

        it = tqdm(it)
    for batch_names, images in it:
        if png_feeder is not None:
            images, keep = png_feeder.finish(images)
            batch_names = [batch_names[i] for i in keep]   # files Pillow refuses are dropped, as the reference does
            if not batch_names:
                continue
        elif gpu_preprocess is not None:
            images = gpu_preprocess(images)
        o = backend.encode_gallery(images.to(backend.device, non_blocking=True), want_f32=False, want_bf16=True,
                                   want_raws_f32=False, want_raws_bf16=keep_raws)
        feats.append(o["feats_bf16"])
        if keep_raws:
            raws.append(o["raws_bf16"])
        names.extend(batch_names)
    f = torch.cat(feats) if feats else torch.empty(0, 32, 256, dtype=getattr(backend, 'act_torch_dtype', torch.bfloat16), device=backend.device)
    r = (torch.cat(raws) if raws else None) if keep_raws else None
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, names)
        all_names = [x for part in gathered for x in part]
    else:
        all_names = names
    # the reference drops unreadable images silently (data_utils.py:191-192 + collate_fn); keep that:
    # rows are numbered by position among the images that were actually encoded
    bounds = None
    if world > 1:
        counts = [len(p) for p in gathered]
        lo = sum(counts[:rank])
        bounds = [sum(counts[:i]) for i in range(world + 1)]   # real [lo, hi) of every rank (rows were dropped)
    return GalleryIndex(feats=f.contiguous(), raws=None if r is None else r.contiguous(), names=all_names, lo=lo,
                        hi=lo + f.shape[0], n_total=len(all_names), bounds=bounds)


# ---------------------------------------------------------------------------------------------------
# query -> top-k (+ subset scores)
# ---------------------------------------------------------------------------------------------------
@torch.no_grad()
def query_topk(backend, index: GalleryIndex, ref_rows: torch.Tensor, input_ids: torch.Tensor,
               attention_mask: torch.Tensor, k: int, subset_rows: Optional[torch.Tensor] = None):
    """ref_rows: global gallery rows of the reference images [Q]; ids/mask [Q,32];
    subset_rows (optional) global rows [Q,m] (-1 = skip).
    Returns (scores [Q,k] fp32, rows [Q,k] int32 global, subset_scores [Q,m] or None), identical on all ranks."""
    dist, rank, world = _dist()
    dev = backend.device
    Q = ref_rows.shape[0]
    ref_rows = ref_rows.to(torch.int64)
    if world == 1:
        fusion = backend.encode_query(index.raws, input_ids, attention_mask, ref_rows=(ref_rows - index.lo))
    else:
        # owner-computes: the rank holding the reference row's raw embeds runs the Q-Former fusion
        mine = (ref_rows >= index.lo) & (ref_rows < index.hi)
        fusion32 = torch.zeros(Q, 256, dtype=torch.float32, device=dev)
        if bool(mine.any()):
            sel = mine.nonzero().flatten()
            f = backend.encode_query(index.raws, input_ids[sel], attention_mask[sel],
                                     ref_rows=(ref_rows[sel] - index.lo))
            fusion32[sel.to(dev)] = f.float()
        dist.all_reduce(fusion32)  # exactly one rank contributes each row: x + 0 + ... is exact
        fusion = fusion32.to(index.feats.dtype)
    send = None
    if index.feats.shape[0] == 0:
        # empty shard (fewer images than ranks): no candidates from this rank
        sc = torch.full((Q, k), float("-inf"), dtype=torch.float32, device=dev)
        ix = torch.full((Q, k), -1, dtype=torch.int32, device=dev)
    elif world > 1 and hasattr(backend, "sim_topk_grouped"):
        # one launch writes every destination rank's [c, k] block straight into the exchange buffer
        send = _exchange_buffer(Q, k, world, dev)
        backend.sim_topk_grouped(fusion, index.feats, k, index.lo, send)
        sc = ix = None
    else:
        sc, ix, _ = backend.sim_topk(fusion, index.feats, k=k, row_offset=index.lo)
    sub = None
    if subset_rows is not None and index.feats.shape[0] == 0:
        sub = torch.full(tuple(subset_rows.shape), float("-inf"), dtype=torch.float32, device=dev)
    elif subset_rows is not None:
        local = subset_rows.to(torch.int64) - index.lo
        local = torch.where((subset_rows >= index.lo) & (subset_rows < index.hi), local, torch.full_like(local, -1))
        sub = backend.gather_scores(fusion, index.feats, local.to(torch.int32))
    if world > 1:
        sc, ix = exchange_and_merge(backend, sc, ix, dist, rank, world, send=send, Q=Q)
        if sub is not None:
            dist.all_reduce(sub, op=dist.ReduceOp.MAX)  # non-owners hold -inf
    return sc, ix, sub


def _exchange_buffer(Q: int, k: int, world: int, dev) -> torch.Tensor:
    """int32 [world, 2, c, k] (c = ceil(Q / world)) pre-filled with the padding candidate (-inf, -1)."""
    c = (Q + world - 1) // world
    send = torch.empty(world, 2, c, k, dtype=torch.int32, device=dev)
    send[:, 0] = torch.tensor(float("-inf")).view(torch.int32).item()
    send[:, 1] = -1
    return send


def exchange_and_merge(backend, sc, ix, dist, rank: int, world: int, send=None, Q=None):
    """Per-shard candidates (sc fp32 / ix int32 [Q,k], identical query order on every rank) -> merged [Q,k] on every
    rank.  ONE all-to-all of one packed buffer delivers to rank r only the candidates of ITS slice of the queries
    (c = ceil(Q / world) queries, P lists of k), rank r merges those c queries, and one all-gather of the merged
    [c, 2k] block makes the (small) result identical everywhere: bytes received per rank 2 * Q * k * 8, flat in the
    number of ranks (an all-gather of all candidates + a replicated merge would be world * Q * k * 8 and Q merges)."""
    if send is None:   # candidates as [Q, k] tensors: pack them into the exchange buffer here
        Q, k = sc.shape
        dev = sc.device
        c = (Q + world - 1) // world
        send = _exchange_buffer(Q, k, world, dev)
        flat_s, flat_i = send[:, 0].reshape(world * c, k), send[:, 1].reshape(world * c, k)   # copies when strided
        flat_s[:Q] = sc.contiguous().view(torch.int32)
        flat_i[:Q] = ix
        send[:, 0] = flat_s.view(world, c, k)
        send[:, 1] = flat_i.view(world, c, k)
    else:              # the scan already wrote it (Model.sim_topk_grouped)
        _, _, c, k = send.shape
        dev = send.device
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)                                       # recv[p] = rank p's lists for MY queries
    if hasattr(backend, "topk_merge_packed"):
        msc, mix = backend.topk_merge_packed(recv)
    else:
        msc, mix = backend.topk_merge(recv[:, 0].contiguous().view(torch.float32), recv[:, 1].contiguous())
    mine = torch.cat([msc.contiguous().view(torch.int32), mix.to(torch.int32)], dim=1)       # [c, 2k]
    allm = torch.empty(world * c, 2 * k, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(allm, mine.contiguous())
    return allm[:Q, :k].contiguous().view(torch.float32), allm[:Q, k:].contiguous()


# ---------------------------------------------------------------------------------------------------
# rerank of the first T candidates  (cirr_test_submission.py:87-112, validate_blip_rerank.py:48-71,196-221)
# ---------------------------------------------------------------------------------------------------
def fetch_raw_rows(index: GalleryIndex, rows: torch.Tensor) -> torch.Tensor:
    """Raw embeds [n,257,Dv] of GLOBAL gallery rows `rows` on this rank's device.  One rank: a gather from the
    resident table.  Row-sharded index (SURVEY §8e "replicas + row fetch"): every rank tells every rank which rows
    it needs (one small all-gather of row ids), owners send exactly those rows point to point (batched isend/irecv:
    NVLink under NCCL, also valid under gloo).  Collective: every rank must call it the same number of times; a rank
    with nothing to fetch passes an empty `rows`."""
    dist, rank, world = _dist()
    rows = rows.to(torch.int64).cpu()
    dev = index.raws.device
    if world == 1:
        return index.raws[(rows - index.lo).to(dev)]
    n = torch.tensor([rows.numel()], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c) for c in counts]
    width = max(max(counts), 1)
    mine = torch.full((width,), -1, dtype=torch.int64, device=dev)
    mine[: rows.numel()] = rows.to(dev)
    req = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(req, mine)
    req = [r[:c].cpu() for r, c in zip(req, counts)]                      # what each rank needs
    own = index.owner(rows, world)                                        # who holds what I need
    out = torch.empty((rows.numel(),) + tuple(index.raws.shape[1:]), dtype=index.raws.dtype, device=dev)
    ops, recv_bufs, keep_alive = [], {}, []
    for p in range(world):
        wanted = req[p][index.owner(req[p], world) == rank]               # rows rank p needs from me, in p's order
        if p == rank:
            out[(own == rank).nonzero().flatten().to(dev)] = index.raws[(wanted - index.lo).to(dev)]
            continue
        if wanted.numel():
            buf = index.raws[(wanted - index.lo).to(dev)].contiguous()
            keep_alive.append(buf)
            ops.append(dist.P2POp(dist.isend, buf, p))
        n_from_p = int((own == p).sum())
        if n_from_p:
            recv_bufs[p] = torch.empty((n_from_p,) + tuple(index.raws.shape[1:]), dtype=index.raws.dtype, device=dev)
            ops.append(dist.P2POp(dist.irecv, recv_bufs[p], p))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for p, buf in recv_bufs.items():
        out[(own == p).nonzero().flatten().to(dev)] = buf
    return out


@torch.no_grad()
def rerank_topk(backend, index: GalleryIndex, top_rows: torch.Tensor, ref_rows: torch.Tensor, input_ids: torch.Tensor,
                attention_mask: torch.Tensor, T: int) -> torch.Tensor:
    """Re-order the first T entries of each query's ranking `top_rows` [Q,K] (global rows, identical on all ranks) by
    `inference_rerank` probability, descending, ties keeping the first-stage order — the reference's loop
    (cirr_test_submission.py:87-112) on integer rows.  Queries are independent, so ranks split them contiguously
    (SURVEY §8e); the raw embeds of each chunk's reference + candidate rows come from their owner ranks
    (`fetch_raw_rows`), and one all-gather of the re-ordered rows makes the result identical everywhere."""
    dist, rank, world = _dist()
    top_rows = top_rows.cpu().to(torch.int64).clone()
    ref_rows = ref_rows.cpu().to(torch.int64)
    Q = top_rows.shape[0]
    max_pairs = int(getattr(backend, "max_pairs", 0))
    if max_pairs < T:
        raise ValueError("rerank needs a model built with max_pairs >= top (Blip2QformerCirRerank)")
    step = max(1, max_pairs // T)
    qlo, qhi = shard_range(Q, rank, world)
    longest = max(shard_range(Q, r, world)[1] - shard_range(Q, r, world)[0] for r in range(world))
    dev = index.raws.device
    for it in range((longest + step - 1) // step):                        # same trip count on every rank
        s, e = min(qlo + it * step, qhi), min(qlo + (it + 1) * step, qhi)
        cand = top_rows[s:e, :T]
        need = torch.cat([ref_rows[s:e], cand.reshape(-1)])
        uniq, inv = torch.unique(need, return_inverse=True)
        table = fetch_raw_rows(index, uniq)
        if e == s:
            continue
        R_ = e - s
        p = backend.rerank_rows(table, inv[:R_].to(torch.int32).to(dev), inv[R_:].to(torch.int32).to(dev),
                                input_ids[s:e], attention_mask[s:e], T).reshape(R_, T).cpu()
        order = torch.argsort(1 - p, dim=-1, stable=True)
        top_rows[s:e, :T] = torch.gather(cand, 1, order)
    if world > 1:
        mine = torch.full((longest, T), -1, dtype=torch.int64, device=dev)
        mine[: qhi - qlo] = top_rows[qlo:qhi, :T].to(dev)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        for r in range(world):
            lo_r, hi_r = shard_range(Q, r, world)
            top_rows[lo_r:hi_r, :T] = parts[r][: hi_r - lo_r].cpu()
    return top_rows


def _subset_scores_after_rerank(ranked: torch.Tensor, group_rows: torch.Tensor, sub_scores: torch.Tensor,
                                T: int) -> torch.Tensor:
    """Subset scores that reproduce "the re-ordered global list restricted to the group members"
    (cirr_test_submission.py:121-123, validate_blip_rerank.py:222-228): a member found at position p < T of the
    re-ordered list scores 1e6 - p (above any similarity), the others keep their first-stage similarity."""
    ranked, group_rows = ranked.cpu().to(torch.int64), group_rows.cpu().to(torch.int64)
    hit = (ranked[:, :T, None] == group_rows[:, None, :]) & (group_rows[:, None, :] >= 0)     # [Q,T,m]
    pos = torch.where(hit.any(dim=1), hit.float().argmax(dim=1), torch.full(group_rows.shape, -1))
    return torch.where(pos >= 0, 1e6 - pos.float(), sub_scores.cpu().float())


def _tokenize(backend, captions):
    tok = backend.tokenizer(list(captions), padding="max_length", truncation=True, max_length=32, return_tensors="pt")
    return tok.input_ids, tok.attention_mask


# ---------------------------------------------------------------------------------------------------
# metrics on integer rows  (validate_blip.py:44-55, 253-285)
# ---------------------------------------------------------------------------------------------------
def cirr_recalls_from_topk(top_rows: torch.Tensor, ref_rows: torch.Tensor, tgt_rows: torch.Tensor,
                           group_rows: torch.Tensor, group_scores: torch.Tensor):
    """top_rows [Q,51] ranking; group_rows/scores [Q,6].  Returns the 7 recalls in percent, as
    compute_cirr_val_metrics does: the reference row is deleted from the ranking first (:258-261)."""
    top_rows = top_rows.cpu().to(torch.int64)
    ref_rows, tgt_rows = ref_rows.cpu().to(torch.int64), tgt_rows.cpu().to(torch.int64)
    Q, K1 = top_rows.shape
    if bool((ref_rows == tgt_rows).any()):
        raise AssertionError("a query's target equals its reference: no positive left after reference removal")
    keep = top_rows != ref_rows[:, None]
    # stable compaction of each row to its first K1-1 kept entries
    pos = torch.cumsum(keep.to(torch.int64), dim=1) - 1
    ranked = torch.full((Q, K1), -1, dtype=torch.int64)
    qi = torch.arange(Q)[:, None].expand(Q, K1)
    ranked[qi[keep], pos[keep]] = top_rows[keep]
    # no truncation: a row that held its reference ends in one -1 slot (never a label); a row that did not (the
    # reference ranked below K1, or the caller removed it already, as the rerank path does) keeps all K1 entries
    labels = (ranked == tgt_rows[:, None]) & (ranked >= 0)
    rec = lambda kk: (labels[:, :kk].any(dim=1).sum().item() / Q) * 100.0  # noqa: E731
    # subset ranking: group members without the reference, by (score desc, row asc) like the global order
    g_rows = group_rows.cpu().to(torch.int64)
    g_sc = group_scores.cpu().float().clone()
    g_sc[g_rows == ref_rows[:, None]] = float("-inf")
    g_sc[g_rows < 0] = float("-inf")
    order = torch.argsort(g_rows, dim=1, stable=True)
    g_rows_s, g_sc_s = torch.gather(g_rows, 1, order), torch.gather(g_sc, 1, order)
    order2 = torch.argsort(-g_sc_s, dim=1, stable=True)
    g_ranked = torch.gather(g_rows_s, 1, order2)
    g_valid = torch.gather(g_sc_s, 1, order2) > float("-inf")
    g_labels = (g_ranked == tgt_rows[:, None]) & g_valid
    if not bool((g_labels.sum(dim=1) == 1).all()):
        raise AssertionError("each query needs exactly one positive among its group members "
                             "(validate_blip.py:274)")
    grec = lambda kk: (g_labels[:, :kk].any(dim=1).sum().item() / Q) * 100.0  # noqa: E731
    return grec(1), grec(2), grec(3), rec(1), rec(5), rec(10), rec(50)


def fiq_recalls_from_topk(top_rows: torch.Tensor, tgt_rows: torch.Tensor):
    top_rows = top_rows.cpu().to(torch.int64)
    labels = top_rows == tgt_rows.cpu().to(torch.int64)[:, None]
    Q = top_rows.shape[0]
    return (labels[:, :10].any(dim=1).sum().item() / Q) * 100.0, (labels[:, :50].any(dim=1).sum().item() / Q) * 100.0


# ---------------------------------------------------------------------------------------------------
# drop-in entry points (same signatures as the reference's)
# ---------------------------------------------------------------------------------------------------
def as_index(index_features, index_names) -> GalleryIndex:
    """Accept what `extract_index_blip_features` returns: a GalleryIndex, or the reference's
    (features, raw_features) tuple of tensors."""
    if isinstance(index_features, GalleryIndex):
        return index_features
    feats, raws = index_features[0], index_features[-1]
    return GalleryIndex(feats=feats.contiguous(), raws=raws.contiguous(), names=list(index_names))


class IndexFeatures(tuple):
    """What the fast `extract_index_blip_features` returns as `index_features`: behaves like the reference's
    (features, raw_features) tuple (validate_blip.py:169,377 index it with [0], [1], [-1]) and carries the
    sharded GalleryIndex for the fast compute_* functions."""

    def __new__(cls, index: GalleryIndex):
        obj = super().__new__(cls, (index.feats, index.raws))
        obj.index = index
        return obj


def extract_index_blip_features(dataset, blip_model, save_memory=False):
    """utils.py:46-77 -> ((index_features, index_features_raw), index_names)."""
    from torch.utils.data.dataloader import default_collate

    def collate_fn(batch):  # utils.py:141-148
        return default_collate([b for b in batch if b is not None])

    index = build_index(dataset, blip_model, batch_size=64, num_workers=2, collate_fn=collate_fn)
    return IndexFeatures(index), index.names


@torch.no_grad()
def compute_cirr_val_metrics(relative_val_dataset, blip_model, index_features, index_names, txt_processors,
                             rerank_top: int = 0):
    """validate_blip.py:232-285 -> (gR@1, gR@2, gR@3, R@1, R@5, R@10, R@50) in percent.
    rerank_top = T > 0: validate_blip_rerank.py:166-239 — after the reference image is deleted from each ranking, its
    first T entries are re-ordered by `inference_rerank` (the reference hard-codes T = 200) and the subset ranking
    follows the re-ordered list."""
    index = index_features.index if isinstance(index_features, IndexFeatures) else as_index(index_features,
                                                                                            index_names)
    ref_names, tgt_names, caps, groups = [], [], [], []
    for i in range(len(relative_val_dataset)):
        item = relative_val_dataset[i]
        if item is None:
            continue
        r, t, c, g = item
        ref_names.append(r)
        tgt_names.append(t)
        caps.append(txt_processors["eval"](c))
        groups.append(list(g))
    ref_rows, tgt_rows = index.rows_of(ref_names), index.rows_of(tgt_names)
    group_rows = torch.tensor([[index.name_to_row.get(n, -1) for n in g] for g in groups], dtype=torch.int64)
    ids, mask = _tokenize(blip_model, caps)
    tops, subs = [], []
    B = max(1, blip_model.max_queries)
    for s in range(0, len(caps), B):
        sl = slice(s, s + B)
        _, ix, sub = query_topk(blip_model, index, ref_rows[sl], ids[sl], mask[sl],
                                k=min(max(50, rerank_top) + 1, max(index.n_total, 51)), subset_rows=group_rows[sl])
        tops.append(ix.cpu())
        subs.append(sub.cpu())
    top_rows, sub_scores = torch.cat(tops).to(torch.int64), torch.cat(subs)
    if rerank_top > 0:
        Q, K1 = top_rows.shape
        keep = top_rows != ref_rows[:, None]                      # :190-194 reference removed BEFORE the rerank
        ranked = torch.stack([r[k_][: K1 - 1] for r, k_ in zip(top_rows, keep)])
        T = min(rerank_top, ranked.shape[1], index.n_total - 1)
        ranked = rerank_topk(blip_model, index, ranked, ref_rows, ids, mask, T)
        sub_scores = _subset_scores_after_rerank(ranked, group_rows, sub_scores, T)
        top_rows = ranked
    return cirr_recalls_from_topk(top_rows, ref_rows, tgt_rows, group_rows, sub_scores)


@torch.no_grad()
def compute_fiq_val_metrics(relative_val_dataset, blip_model, index_features, index_names, txt_processors,
                            save_memory=False, rerank_top: int = 0):
    """validate_blip.py:24-57 -> (R@10, R@50) in percent; captions joined as :180-183.
    rerank_top = T > 0: validate_blip_rerank.py:24-96 — the first T entries of each ranking re-ordered by
    `inference_rerank` before the labels are taken (the reference hard-codes T = 40)."""
    index = index_features.index if isinstance(index_features, IndexFeatures) else as_index(index_features,
                                                                                            index_names)
    ref_names, tgt_names, caps = [], [], []
    for i in range(len(relative_val_dataset)):
        item = relative_val_dataset[i]
        if item is None:
            continue
        r, t, c = item
        ref_names.append(r)
        tgt_names.append(t)
        c0, c1 = c[0], c[1]
        caps.append(txt_processors["eval"](f"{c0.strip('.?, ').capitalize()} and {c1.strip('.?, ')}"))
    ref_rows, tgt_rows = index.rows_of(ref_names), index.rows_of(tgt_names)
    ids, mask = _tokenize(blip_model, caps)
    tops = []
    B = max(1, blip_model.max_queries)
    for s in range(0, len(caps), B):
        sl = slice(s, s + B)
        _, ix, _ = query_topk(blip_model, index, ref_rows[sl], ids[sl], mask[sl], k=max(50, rerank_top))
        tops.append(ix.cpu())
    top = torch.cat(tops)
    if rerank_top > 0:
        top = rerank_topk(blip_model, index, top, ref_rows, ids, mask, min(rerank_top, index.n_total))
    if not bool((top >= 0).all()) and index.n_total >= 50:
        raise AssertionError("top-k returned unfilled slots")
    return fiq_recalls_from_topk(top, tgt_rows)


# ---------------------------------------------------------------------------------------------------
# CIRR test-split submission on integer rows  (cirr_test_submission.py:60-132)
# ---------------------------------------------------------------------------------------------------
def cirr_submission_from_topk(top_rows: torch.Tensor, ref_rows: torch.Tensor, group_rows: torch.Tensor,
                              group_scores: torch.Tensor, index_names: Sequence[str], pairs_id: Sequence,
                              k_global: int = 50, k_group: int = 3):
    """top_rows [Q, >= k_global + 1]: ranking by similarity (ties: lower row), reference possibly included;
    group_rows / group_scores [Q, m]: the query's group members and their similarities.
    Returns (pairid_to_predictions, pairid_to_group_predictions) exactly as generate_cirr_test_dicts builds them:
    the reference row is deleted from the ranking (:115-119), the subset ranking is the global order restricted
    to the group members (:121-123), then top-50 / top-3 names (:126-129)."""
    top_rows = top_rows.cpu().to(torch.int64)
    ref_rows = ref_rows.cpu().to(torch.int64)
    names = np.asarray(list(index_names), dtype=object)
    Q = top_rows.shape[0]
    glob, grp = {}, {}
    g_rows = group_rows.cpu().to(torch.int64)
    g_sc = group_scores.cpu().float()
    for j in range(Q):
        row = top_rows[j]
        ranked = row[(row != ref_rows[j]) & (row >= 0)][:k_global]
        glob[str(int(pairs_id[j]))] = names[ranked.numpy()].tolist()
        members = [(-float(g_sc[j, i]), int(g_rows[j, i])) for i in range(g_rows.shape[1])
                   if int(g_rows[j, i]) >= 0 and int(g_rows[j, i]) != int(ref_rows[j])]
        members.sort()  # score descending, ties by lower row: the order argsort(1 - sim) gives the full ranking
        seen, sub = set(), []
        for _, r in members:
            if r not in seen:
                seen.add(r)
                sub.append(r)
        grp[str(int(pairs_id[j]))] = names[np.asarray(sub[:k_group], dtype=np.int64)].tolist()
    return glob, grp


@torch.no_grad()
def generate_cirr_test_dicts(relative_test_dataset, blip_model, index_features, index_names, txt_processors,
                             rerank=False, top: int = 50):
    """Same signature and result as cirr_test_submission.py:60-132, computed from the fused scan's integer rows:
    one pass of top-(50+1) per query batch + 6 gathered subset scores instead of a [Q, N] similarity matrix, a full
    argsort and object-array string compares.  rerank=True re-orders each query's first `top` candidates by
    `inference_rerank` probabilities (:87-112; the reference hard-codes top = 50)."""
    index = index_features.index if isinstance(index_features, IndexFeatures) else as_index(index_features,
                                                                                            index_names)
    pairs_id, ref_names, caps, groups = [], [], [], []
    for i in range(len(relative_test_dataset)):
        item = relative_test_dataset[i]
        if item is None:
            continue
        pid, r, c, g = item
        pairs_id.append(pid)
        ref_names.append(r)
        caps.append(txt_processors["eval"](c))
        groups.append(list(g))
    ref_rows = index.rows_of(ref_names)
    group_rows = torch.tensor([[index.name_to_row.get(n, -1) for n in g] for g in groups], dtype=torch.int64)
    ids, mask = _tokenize(blip_model, caps)
    k = min(max(top, 50) + 1, index.n_total)
    tops, subs = [], []
    B = max(1, blip_model.max_queries)
    for s in range(0, len(caps), B):
        sl = slice(s, s + B)
        _, ix, sub = query_topk(blip_model, index, ref_rows[sl], ids[sl], mask[sl], k=k, subset_rows=group_rows[sl])
        tops.append(ix.cpu().to(torch.int64))
        subs.append(sub.cpu())
    top_rows, sub_scores = torch.cat(tops), torch.cat(subs)
    if rerank:
        T = min(top, top_rows.shape[1])
        top_rows = rerank_topk(blip_model, index, top_rows, ref_rows, ids, mask, T)
        # :115-123 the subset ranking is the RE-ORDERED global list restricted to the group members: members inside
        # the first T take their new positions (ahead of every member outside, whose first-stage order stands)
        sub_scores = _subset_scores_after_rerank(top_rows, group_rows, sub_scores, T)
    return cirr_submission_from_topk(top_rows, ref_rows, group_rows, sub_scores, index.names, pairs_id)


# ---------------------------------------------------------------------------------------------------
# on-disk gallery index  (the reference recomputes the index on every run,
# blip_validate.py:80,119)
# ---------------------------------------------------------------------------------------------------
INDEX_MAGIC = b"SPRCIDX1"


def save_index(index: GalleryIndex, path: str) -> None:
    """One file: magic, u64 header length, JSON header, then the feature block [N,32,256] and (optionally) the raw
    embedding block [N,257,Dv], both in the 16-bit format they are resident in.  Blocks start on 4096-byte
    boundaries so a rank can map exactly its row range.  Only a whole (unsharded) index is written."""
    import json

    if index.lo != 0 or index.hi != index.n_total:
        raise ValueError("save_index needs the whole index (gather the shards first)")
    feats = index.feats.contiguous()
    raws = index.raws.contiguous() if index.raws is not None else None
    dt = {torch.bfloat16: "bf16", torch.float16: "fp16"}[feats.dtype]
    hdr = {"version": 1, "n": index.n_total, "dtype": dt, "tokens": 32, "dim": 256,
           "raw_tokens": 257 if raws is not None else 0, "raw_dim": int(raws.shape[-1]) if raws is not None else 0,
           "names": list(index.names)}
    blob = json.dumps(hdr).encode()
    off_feats = (len(INDEX_MAGIC) + 8 + len(blob) + 4095) // 4096 * 4096
    feats_bytes = feats.numel() * 2
    off_raws = (off_feats + feats_bytes + 4095) // 4096 * 4096
    with open(path, "wb") as f:
        f.write(INDEX_MAGIC)
        f.write(len(blob).to_bytes(8, "little"))
        f.write(blob)
        f.seek(off_feats)
        f.write(feats.cpu().view(torch.int16).numpy().tobytes())
        if raws is not None:
            f.seek(off_raws)
            for s0 in range(0, raws.shape[0], 1024):   # bounded host staging
                f.write(raws[s0:s0 + 1024].cpu().view(torch.int16).numpy().tobytes())


def load_index(path: str, device, rank: int = 0, world: int = 1, with_raws: bool = True) -> GalleryIndex:
    """Maps rows [lo, hi) of rank `rank` straight from the file to `device` (no full-index host copy)."""
    import json

    with open(path, "rb") as f:
        if f.read(len(INDEX_MAGIC)) != INDEX_MAGIC:
            raise ValueError(f"{path}: not a sprc-b200 gallery index")
        n_hdr = int.from_bytes(f.read(8), "little")
        hdr = json.loads(f.read(n_hdr).decode())
    if hdr.get("version") != 1:
        raise ValueError(f"{path}: unsupported index version {hdr.get('version')}")
    n = int(hdr["n"])
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[hdr["dtype"]]
    lo, hi = shard_range(n, rank, world)
    off_feats = (len(INDEX_MAGIC) + 8 + n_hdr + 4095) // 4096 * 4096
    row_f = hdr["tokens"] * hdr["dim"]
    off_raws = (off_feats + n * row_f * 2 + 4095) // 4096 * 4096

    def block(offset, row_elems, shape_tail):
        if hi == lo:
            return torch.empty((0,) + shape_tail, dtype=dt, device=device)
        m = np.memmap(path, dtype=np.int16, mode="r", offset=offset + lo * row_elems * 2, shape=((hi - lo) * row_elems,))
        import warnings

        with warnings.catch_warnings():   # the mapping is read-only on purpose: it is only the source of the copy
            warnings.simplefilter("ignore", UserWarning)
            t = torch.from_numpy(np.ascontiguousarray(m)).view(dt).reshape((hi - lo,) + shape_tail)
        out = t.to(device)
        return out.clone() if out.device.type == "cpu" else out   # never hand out a tensor aliasing the read-only map

    feats = block(off_feats, row_f, (hdr["tokens"], hdr["dim"]))
    raws = None
    if with_raws and hdr["raw_tokens"]:
        raws = block(off_raws, hdr["raw_tokens"] * hdr["raw_dim"], (hdr["raw_tokens"], hdr["raw_dim"]))
    return GalleryIndex(feats=feats, raws=raws, names=list(hdr["names"]), lo=lo, hi=hi, n_total=n)
