"""Python host mirror of the reference's model surface over the C ABI (include/sprc_b200.h).

`Blip2QformerCirAlignPrompt` keeps the names, argument meaning and return shapes of
lavis/models/blip2_models/blip2_qformer_cir_align_prompt.py (class :26, `inference` :312-361,
`extract_target_features` :364-386, `from_config` :502-529) and of `inference_rerank`
(blip2_qformer_cir_rerank.py:399-445), so src/blip_validate.py, src/cirr_test_submission.py and
src/validate_blip.py drive it unchanged.  All arithmetic happens in libsprc_b200.so (hand-written
sm_100a kernels); torch only owns device buffers.  There is no CPU path: constructing the model
without the library or without a CUDA device raises.
"""
from __future__ import annotations

import os
from collections import namedtuple
from typing import List, Optional, Sequence

import torch

from . import _lib as L
from .tokenizer import BlipCaptionProcessor, OfflineBertTokenizer, TokenBatch

_IncompatibleKeys = namedtuple("IncompatibleKeys", ["missing_keys", "unexpected_keys"])

_VIT = {"eva_clip_g": (L.VIT_EVA_G, 1408), "clip_L": (L.VIT_CLIP_L, 1024)}
_DTYPES = {torch.float32: L.F32, torch.float16: L.F16, torch.bfloat16: L.BF16}

# model_type -> constructor arguments (lavis/configs/models/blip2/blip2_pretrain.yaml:6-36 and
# blip2_pretrain_vitL.yaml:6-37; PRETRAINED_MODEL_CONFIG_DICT at align_prompt.py:38-42)
MODEL_TYPES = {
    "pretrain": dict(vit_model="eva_clip_g", img_size=224, num_query_token=32),
    "pretrain_vitL": dict(vit_model="clip_L", img_size=224, num_query_token=32),
}


class Blip2QformerCirAlignPrompt:
    PRETRAINED_MODEL_CONFIG_DICT = {
        "pretrain": "configs/models/blip2/blip2_pretrain.yaml",
        "pretrain_vitL": "configs/models/blip2/blip2_pretrain_vitL.yaml",
    }

    def __init__(self, vit_model="eva_clip_g", img_size=224, drop_path_rate=0, use_grad_checkpoint=False,
                 vit_precision="fp16", freeze_vit=True, num_query_token=32, cross_attention_freq=2, embed_dim=256,
                 max_txt_len=32, *, device=None, max_images=64, max_queries=64, max_pairs=0, vit_depth=0,
                 qf_layers=0, tokenizer=None, act_dtype=None):
        if vit_model not in _VIT:
            raise ValueError("vit model must be eva_clip_g or clip_L")
        if img_size != 224 or num_query_token != 32 or cross_attention_freq != 2 or embed_dim != 256 \
                or max_txt_len != 32:
            raise ValueError("libsprc_b200 is specialised for 224x224 images, 32 query tokens, cross-attention "
                             "every 2nd layer, 256-d ITC heads and 32-token captions")
        self._lib = L.load()  # raises if the CUDA library is missing: no fallback
        if not torch.cuda.is_available():
            raise L.SprcError("Blip2QformerCirAlignPrompt needs a CUDA device (sm_100a); there is no CPU path")
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if dev.type != "cuda":
            raise L.SprcError(f"Blip2QformerCirAlignPrompt cannot run on {dev}; there is no CPU path")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self._device = dev
        self.vit_name = vit_model
        self.vit_width = _VIT[vit_model][1]
        self.max_txt_len = max_txt_len
        self.max_images, self.max_queries, self.max_pairs = int(max_images), int(max_queries), int(max_pairs)
        self.tokenizer = tokenizer if tokenizer is not None else OfflineBertTokenizer()
        self.training = False
        # 16-bit operand format of every kernel: fp16 (default: the reference's own autocast precision, blip2.py:36-44,
        # and the mode that meets the north star's 1e-3 embedding tolerance at unchanged speed) or bf16;
        # SPRC_ACT_DTYPE overrides the default for unchanged reference scripts
        act_dtype = act_dtype or os.environ.get("SPRC_ACT_DTYPE", "fp16")
        if act_dtype not in ("bf16", "fp16"):
            raise ValueError("act_dtype must be 'bf16' or 'fp16'")
        self.act_dtype = act_dtype
        self.act_torch_dtype = torch.float16 if act_dtype == "fp16" else torch.bfloat16
        cfg = L.SprcConfig(_VIT[vit_model][0], int(vit_depth), int(qf_layers), self.max_images, self.max_queries,
                           self.max_pairs, dev.index, 1 if act_dtype == "fp16" else 0)
        h = L.c_void_p()
        with torch.cuda.device(dev):
            L.check(self._lib.sprc_create(L.ctypes.byref(cfg), L.ctypes.byref(h)))
        self._h = h
        self._gallery_cache = None
        self._ragged = os.environ.get("SPRC_RAGGED", "1") != "0"   # A/B switch, see encode_query

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_config(cls, cfg, **kw):
        get = cfg.get if hasattr(cfg, "get") else (lambda k, d=None: getattr(cfg, k, d))
        return cls(vit_model=get("vit_model", "eva_clip_g"), img_size=get("image_size", 224),
                   num_query_token=get("num_query_token", 32), cross_attention_freq=get("cross_attention_freq", 2),
                   max_txt_len=get("max_txt_len", 32), **kw)

    @classmethod
    def from_pretrained(cls, model_type, **kw):
        """The reference downloads BLIP-2 weights here (base_model.py:58-72); offline, weights arrive via
        `load_state_dict` (blip_validate.py:107-109) or an explicit checkpoint path."""
        if model_type not in MODEL_TYPES:
            raise KeyError(f"Unknown model type {model_type}")
        return cls(**MODEL_TYPES[model_type], **kw)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.sprc_destroy(h)
            except Exception:
                pass
            self._h = None

    # ------------------------------------------------------------------ nn.Module-like surface
    @property
    def device(self):
        return self._device

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        # Inference only.  The reference's CIRR scripts leave dropout on; this
        # implementation is deterministic by design.
        self.training = False
        return self

    def float(self):
        return self

    def half(self):
        return self

    def to(self, device=None, *a, **k):
        if device is None or isinstance(device, torch.dtype):
            return self
        d = torch.device(device)
        if d.type != "cuda" or (d.index is not None and d.index != self._device.index):
            raise L.SprcError(f"model lives on {self._device}; cannot move to {d} (no CPU path, one handle per GPU)")
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", device) if isinstance(device, int) else (device or self._device))

    def parameters(self):
        return iter(())

    def load_state_dict(self, state_dict, strict=False):
        """Accepts the reference checkpoint layout, fp32 / fp16, CPU or CUDA tensors."""
        descs, keep = [], []
        for name, t in state_dict.items():
            if not torch.is_tensor(t) or t.dtype not in _DTYPES:
                continue  # e.g. position_ids (int64): not used
            t = t.detach().contiguous()
            keep.append(t)
            d = L.SprcTensorDesc()
            d.name = name.encode()
            d.dtype = _DTYPES[t.dtype]
            d.ndim = t.dim()
            for i, s in enumerate(t.shape[:4]):
                d.shape[i] = s
            if t.dim() > 4:
                continue
            d.data = t.data_ptr()
            descs.append(d)
        arr = (L.SprcTensorDesc * len(descs))(*descs)
        n_missing = L.c_int(0)
        with torch.cuda.device(self._device):
            torch.cuda.synchronize()
            L.check(self._lib.sprc_load_weights(self._h, arr, len(descs), L.ctypes.byref(n_missing)))
        missing = []
        for i in range(n_missing.value):
            m = self._lib.sprc_missing_weight(self._h, i)
            if m:
                missing.append(m.decode())
        if strict and missing:
            raise RuntimeError(f"Missing key(s) in state_dict: {missing[:8]}...")
        return _IncompatibleKeys(missing, [])

    # ------------------------------------------------------------------ low-level calls
    def _stream(self):
        return L.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)

    def encode_gallery(self, images: torch.Tensor, want_f32=True, want_bf16=False, want_raws_f32=True,
                       want_raws_bf16=False):
        """images fp32 [B,3,224,224] on the device -> dict of requested outputs."""
        images = images.to(self._device, torch.float32).contiguous()
        B = images.shape[0]
        if tuple(images.shape[1:]) != (3, 224, 224):
            raise ValueError(f"expected images [B,3,224,224], got {tuple(images.shape)}")
        Dv = self.vit_width
        out = {}
        with torch.cuda.device(self._device):
            if want_f32:
                out["feats"] = torch.empty(B, 32, 256, device=self._device)
            if want_bf16:
                out["feats_bf16"] = torch.empty(B, 32, 256, device=self._device, dtype=self.act_torch_dtype)
            if want_raws_f32:
                out["raws"] = torch.empty(B, 257, Dv, device=self._device)
            if want_raws_bf16:
                out["raws_bf16"] = torch.empty(B, 257, Dv, device=self._device, dtype=self.act_torch_dtype)
            for s in range(0, B, self.max_images):
                e = min(B, s + self.max_images)
                sl = lambda t: L.ptr(t[s:e]) if t is not None else L.c_void_p(0)  # noqa: E731
                L.check(self._lib.sprc_encode_gallery(self._h, L.ptr(images[s:e]), e - s, sl(out.get("feats")),
                                                      sl(out.get("feats_bf16")), sl(out.get("raws")),
                                                      sl(out.get("raws_bf16")), self._stream()))
        return out

    def encode_query(self, reference_embeds: torch.Tensor, input_ids: torch.Tensor, attention_mask: torch.Tensor,
                     ref_rows: Optional[torch.Tensor] = None, out_dtype=None) -> torch.Tensor:
        """fusion_feats [Bq,256].  `reference_embeds` is [Bq,257,Dv] (fp32/bf16) or, with `ref_rows`
        (int32 [Bq]), a resident table [*,257,Dv] whose rows are gathered on the device."""
        out_dtype = out_dtype or self.act_torch_dtype
        ref = reference_embeds
        if ref.dtype not in (torch.float32, self.act_torch_dtype):
            ref = ref.float()
        ref = ref.to(self._device).contiguous()
        ids = input_ids.to(self._device, torch.int64).contiguous()
        # Caption lengths are known on the host (the tokenizer runs there): with a prefix mask on the CPU the query
        # passes run over the live text rows only (sprc_encode_query_lens); otherwise over all 64 padded rows.
        lens = None
        if attention_mask.device.type == "cpu" and self._ragged:
            m = attention_mask.to(torch.int64)
            if bool((m[:, :-1] >= m[:, 1:]).all()) and bool(((m == 0) | (m == 1)).all()):
                lens = m.sum(dim=1).to(torch.int32).contiguous()
        am = attention_mask.to(self._device, torch.int64).contiguous() if lens is None else None
        Bq = ids.shape[0]
        rows = None if ref_rows is None else ref_rows.to(self._device, torch.int32).contiguous()
        out = torch.empty(Bq, 256, device=self._device, dtype=out_dtype)
        with torch.cuda.device(self._device):
            for s in range(0, Bq, self.max_queries):
                e = min(Bq, s + self.max_queries)
                of = L.ptr(out[s:e]) if out_dtype == torch.float32 else L.c_void_p(0)
                ob = L.ptr(out[s:e]) if out_dtype != torch.float32 else L.c_void_p(0)
                rp = L.ptr(ref) if rows is not None else L.ptr(ref[s:e])
                rdt = L.F32 if ref.dtype == torch.float32 else L.BF16
                rw = L.ptr(rows[s:e]) if rows is not None else L.c_void_p(0)
                if lens is not None:
                    L.check(self._lib.sprc_encode_query_lens(self._h, rp, rdt, rw, L.ptr(ids[s:e]), L.ptr(lens[s:e]),
                                                             e - s, of, ob, self._stream()))
                else:
                    L.check(self._lib.sprc_encode_query(self._h, rp, rdt, rw, L.ptr(ids[s:e]), L.ptr(am[s:e]), e - s,
                                                        of, ob, self._stream()))
        return out

    def sim_topk(self, queries_bf16: torch.Tensor, gallery_bf16: torch.Tensor, k: int = 0, row_offset: int = 0,
                 want_full: bool = False):
        """-> (scores [Q,k] fp32, idx [Q,k] int32, full [Q,N] fp32 or None)."""
        q = queries_bf16.to(self._device, self.act_torch_dtype).contiguous()
        g = gallery_bf16
        assert g.dtype == self.act_torch_dtype and g.is_contiguous() and g.device == self._device
        Q, N = q.shape[0], g.shape[0]
        sc = ix = full = None
        with torch.cuda.device(self._device):
            if k > 0:
                sc = torch.empty(Q, k, device=self._device)
                ix = torch.empty(Q, k, device=self._device, dtype=torch.int32)
            if want_full:
                full = torch.empty(Q, N, device=self._device)
            L.check(self._lib.sprc_sim_topk(self._h, L.ptr(q), Q, L.ptr(g), N, row_offset, k, L.ptr(sc), L.ptr(ix),
                                            L.ptr(full), self._stream()))
        return sc, ix, full

    def sim_topk_grouped(self, queries_bf16: torch.Tensor, gallery_bf16: torch.Tensor, k: int, row_offset: int,
                         exchange: torch.Tensor):
        """The scan of this rank's shard for ALL ranks' queries in one launch, written straight into the packed exchange
        buffer `exchange` int32 [P, 2, c, k] (sprc_sim_topk_grouped): query q goes to slot q // c, row q % c; rows the
        scan does not write (P * c > Q) keep what the caller put there."""
        q = queries_bf16.to(self._device, self.act_torch_dtype).contiguous()
        g = gallery_bf16
        assert g.dtype == self.act_torch_dtype and g.is_contiguous() and g.device == self._device
        P, two, c, kk = exchange.shape
        assert two == 2 and kk == k and exchange.dtype == torch.int32 and exchange.is_contiguous()
        assert q.shape[0] <= P * c
        with torch.cuda.device(self._device):
            L.check(self._lib.sprc_sim_topk_grouped(self._h, L.ptr(q), q.shape[0], L.ptr(g), g.shape[0], row_offset, k,
                                                    L.ptr(exchange[0, 0]), L.ptr(exchange[0, 1]), c, 2 * c * k,
                                                    self._stream()))
        return exchange

    def gather_scores(self, queries_bf16, gallery_bf16, rows: torch.Tensor) -> torch.Tensor:
        q = queries_bf16.to(self._device, self.act_torch_dtype).contiguous()
        rows = rows.to(self._device, torch.int32).contiguous()
        out = torch.empty(rows.shape, device=self._device)
        with torch.cuda.device(self._device):
            L.check(self._lib.sprc_gather_scores(self._h, L.ptr(q), q.shape[0], L.ptr(gallery_bf16),
                                                 gallery_bf16.shape[0], L.ptr(rows), rows.shape[1], L.ptr(out),
                                                 self._stream()))
        return out

    def topk_merge(self, cand_score: torch.Tensor, cand_idx: torch.Tensor):
        """[P,Q,k] candidates (e.g. all-gathered per-shard top-k) -> merged ([Q,k], [Q,k])."""
        P, Q, k = cand_score.shape
        cs = cand_score.to(self._device, torch.float32).contiguous()
        ci = cand_idx.to(self._device, torch.int32).contiguous()
        sc = torch.empty(Q, k, device=self._device)
        ix = torch.empty(Q, k, device=self._device, dtype=torch.int32)
        with torch.cuda.device(self._device):
            L.check(self._lib.sprc_topk_merge(self._h, L.ptr(cs), L.ptr(ci), P, Q, k, L.ptr(sc), L.ptr(ix),
                                              self._stream()))
        return sc, ix

    def topk_merge_packed(self, cand: torch.Tensor):
        """cand int32 [P,2,Q,k] (scores as bit patterns, then global rows: the all-to-all exchange buffer) ->
        merged ([Q,k] fp32, [Q,k] int32), no repacking copies."""
        P, two, Q, k = cand.shape
        assert two == 2 and cand.dtype == torch.int32 and cand.is_contiguous() and cand.device == self._device
        sc = torch.empty(Q, k, device=self._device)
        ix = torch.empty(Q, k, device=self._device, dtype=torch.int32)
        with torch.cuda.device(self._device):
            L.check(self._lib.sprc_topk_merge_packed(self._h, L.ptr(cand), P, Q, k, L.ptr(sc), L.ptr(ix),
                                                     self._stream()))
        return sc, ix

    def query_topk_host(self, raws_bf16, gallery_bf16, ref_rows_host, ids_host, mask_host, k, out_score_host,
                        out_idx_host):
        """End-to-end step on HOST buffers (pinned CPU tensors): H2D, fusion, scan, top-k, D2H."""
        Bq = ids_host.shape[0]
        with torch.cuda.device(self._device):
            L.check(self._lib.sprc_query_topk_host(self._h, L.ptr(raws_bf16), L.ptr(gallery_bf16),
                                                   gallery_bf16.shape[0], L.ptr(ref_rows_host), L.ptr(ids_host),
                                                   L.ptr(mask_host), Bq, k, L.ptr(out_score_host),
                                                   L.ptr(out_idx_host), self._stream()))

    def query_topk_host_submit(self, raws_bf16, gallery_bf16, ref_rows_host, ids_host, mask_host, k, out_score_host,
                               out_idx_host):
        """Enqueue one end-to-end step (H2D, fusion, scan, top-k, D2H) and return; `query_topk_host_wait` blocks until
        the oldest submitted step has its results in the host buffers it was given (pinned, untouched until then)."""
        with torch.cuda.device(self._device):
            L.check(self._lib.sprc_query_topk_host_submit(self._h, L.ptr(raws_bf16), L.ptr(gallery_bf16),
                                                          gallery_bf16.shape[0], L.ptr(ref_rows_host), L.ptr(ids_host),
                                                          L.ptr(mask_host), ids_host.shape[0], k,
                                                          L.ptr(out_score_host), L.ptr(out_idx_host), self._stream()))

    def query_topk_host_wait(self):
        L.check(self._lib.sprc_query_topk_host_wait(self._h))

    def query_topk_strings_submit(self, raws_bf16, gallery_bf16, ref_rows_host, captions, k, out_score_host,
                                  out_idx_host):
        """One end-to-end step from caption STRINGS (what `inference` receives, align_prompt.py:312-329): the C++
        tokenizer fills pinned staging inside the library while earlier batches run on the GPU, then the batch is
        enqueued (H2D, fusion, scan, top-k, D2H).  Pair with `query_topk_host_wait`.  A batch holding a caption the C++
        tokenizer hands back (combining marks, final sigma) is tokenised here and submitted from ids."""
        import numpy as np

        tok = self.tokenizer
        if not hasattr(tok, "_native_handle"):
            raise L.SprcError("query_topk_strings needs the library tokenizer (sprc_b200.tokenizer.OfflineBertTokenizer)")
        tok._require_vocab()
        _, th = tok._native_handle()
        n = len(captions)
        try:
            enc = [c.encode("utf-8") for c in captions]
        except UnicodeEncodeError:
            enc = None
        rc = -84
        if enc is not None:
            offs = np.zeros(n + 1, dtype=np.int64)
            np.cumsum([len(e) for e in enc], out=offs[1:])
            with torch.cuda.device(self._device):
                rc = self._lib.sprc_query_topk_strings_submit(
                    self._h, th, L.ptr(raws_bf16), L.ptr(gallery_bf16), gallery_bf16.shape[0], L.ptr(ref_rows_host),
                    b"".join(enc), offs.ctypes.data, n, k, tok.threads, L.ptr(out_score_host), L.ptr(out_idx_host),
                    self._stream())
        if rc == -84:
            b = tok(list(captions), max_length=self.max_txt_len)
            keep = getattr(self, "_strings_keepalive", [])
            ids, mask = b.input_ids.pin_memory(), b.attention_mask.pin_memory()
            self._strings_keepalive = (keep + [(ids, mask)])[-8:]     # host buffers must outlive the async copies
            return self.query_topk_host_submit(raws_bf16, gallery_bf16, ref_rows_host, ids, mask, k, out_score_host,
                                               out_idx_host)
        L.check(rc)

    def query_topk_strings(self, raws_bf16, gallery_bf16, ref_rows_host, captions, k, out_score_host, out_idx_host):
        self.query_topk_strings_submit(raws_bf16, gallery_bf16, ref_rows_host, captions, k, out_score_host,
                                       out_idx_host)
        self.query_topk_host_wait()

    # ------------------------------------------------------------------ the reference's method surface
    def _tokenize(self, text):
        if isinstance(text, TokenBatch):
            return text
        if hasattr(text, "input_ids"):
            return TokenBatch(text.input_ids, text.attention_mask)
        return self.tokenizer(list(text) if not isinstance(text, str) else text, padding="max_length",
                              truncation=True, max_length=self.max_txt_len, return_tensors="pt")

    @torch.no_grad()
    def extract_target_features(self, image, mode="mean"):
        """(image_features [B,32,256] fp32 unit rows, image_embeds_frozen [B,257,Dv] fp32) — align_prompt.py:364-386."""
        o = self.encode_gallery(image, want_f32=True, want_raws_f32=True)
        return o["feats"], o["raws"]

    def _gallery_bf16(self, target_feats: torch.Tensor) -> torch.Tensor:
        if target_feats.dtype == self.act_torch_dtype and target_feats.device == self._device:
            return target_feats.contiguous()
        key = (target_feats.data_ptr(), tuple(target_feats.shape), target_feats._version, target_feats.device)
        if self._gallery_cache is None or self._gallery_cache[0] != key:
            self._gallery_cache = (key, target_feats.to(self._device, self.act_torch_dtype).contiguous())
        return self._gallery_cache[1]

    @torch.no_grad()
    def inference(self, reference_embeds, target_feats, text):
        """sim_i2t [Bq,N] fp32 ([N] when Bq == 1) — align_prompt.py:312-361."""
        tok = self._tokenize(text)
        fusion = self.encode_query(reference_embeds, tok.input_ids, tok.attention_mask)
        _, _, full = self.sim_topk(fusion, self._gallery_bf16(target_feats), k=0, want_full=True)
        return full.squeeze()

    @torch.no_grad()
    def inference_rerank(self, refereence_embeds, target_embeds, text):
        """p [R*T] — blip2_qformer_cir_rerank.py:399-445 (R references, T = len(target)/R candidates each)."""
        tok = self._tokenize(text)
        ref = refereence_embeds.to(self._device)
        tgt = target_embeds.to(self._device)
        R, n = ref.shape[0], tgt.shape[0]
        T = n // R if R > 1 else n
        table = torch.cat([ref, tgt], dim=0).to(self.act_torch_dtype).contiguous()
        ref_rows = torch.arange(R, device=self._device, dtype=torch.int32)
        cand_rows = torch.arange(R, R + R * T, device=self._device, dtype=torch.int32)
        return self.rerank_rows(table, ref_rows, cand_rows, tok.input_ids, tok.attention_mask, T)

    def rerank_rows(self, raws_bf16, ref_rows, cand_rows, input_ids, attention_mask, T):
        R = ref_rows.shape[0]
        ids = input_ids.to(self._device, torch.int64).contiguous()
        lens = None
        if attention_mask.device.type == "cpu" and self._ragged:   # caption lengths known on the host: ragged rows
            m = attention_mask.to(torch.int64)
            if bool((m[:, :-1] >= m[:, 1:]).all()) and bool(((m == 0) | (m == 1)).all()):
                lens = m.sum(dim=1).to(torch.int32).contiguous()
        p = torch.empty(R * T, device=self._device)
        rr = ref_rows.to(self._device, torch.int32).contiguous()
        cr = cand_rows.to(self._device, torch.int32).contiguous()
        with torch.cuda.device(self._device):
            if lens is not None:
                L.check(self._lib.sprc_rerank_lens(self._h, L.ptr(raws_bf16), L.ptr(rr), L.ptr(cr), L.ptr(ids),
                                                   L.ptr(lens), R, T, L.ptr(p), self._stream()))
            else:
                am = attention_mask.to(self._device, torch.int64).contiguous()
                L.check(self._lib.sprc_rerank(self._h, L.ptr(raws_bf16), L.ptr(rr), L.ptr(cr), L.ptr(ids), L.ptr(am), R,
                                              T, L.ptr(p), self._stream()))
        return p


class Blip2QformerCirRerank(Blip2QformerCirAlignPrompt):
    """Checkpoint key `Blip2QformerCirRerank`: same kernels, rerank head enabled."""

    def __init__(self, *a, max_pairs=512, **kw):
        super().__init__(*a, max_pairs=max_pairs, **kw)


class Blip2QformerCirCat(Blip2QformerCirAlignPrompt):
    """`blip2_cir_cat`, the DEFAULT --blip-model-name of cirr_test_submission.py:206 (checkpoint key
    `Blip2QformerCirCat`) — blip2_qformer_cir_cat.py:282-336, 401-428.
    Same ViT, Q-Former passes, ITC heads and kernels as align_prompt; differences kept here on the host:
      * `inference` returns sim / temp (:331) — the ranking is unchanged, the full matrix is scaled;
      * `extract_target_features` returns CPU tensors and takes target_only / ref_only (:401-428);
      * the model has no `prompt_tokens` parameter."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._temp = 0.07  # nn.Parameter(0.07 * torch.ones([])), blip2_qformer_cir_cat.py:84

    def load_state_dict(self, state_dict, strict=False):
        if "temp" in state_dict:
            self._temp = float(state_dict["temp"])
        return super().load_state_dict(state_dict, strict=strict)

    @torch.no_grad()
    def inference(self, reference_embeds, target_feats, text, return_attns=False):
        if return_attns:
            raise NotImplementedError("return_attns (cross-attention maps for visualisation) is outside the hot path")
        return super().inference(reference_embeds, target_feats, text) / self._temp

    @torch.no_grad()
    def extract_target_features(self, image, mode="mean", target_only=False, ref_only=False):
        feats, raws = super().extract_target_features(image, mode)
        if target_only:
            return feats.cpu()
        if ref_only:
            return raws
        return feats.cpu(), raws.cpu()

    @torch.no_grad()
    def inference_rerank(self, refereence_embeds, target_embeds, text):
        """sim [R*T] — blip2_qformer_cir_cat.py:337-398.  `target_embeds` are candidate FEATURE blocks
        [R*T,32,256] (the reference's docstring says raw embeds; the matmul at :392 needs 256-d rows): each reference's
        composed query (the T repeated rows of :349-361 are identical, computed once here) is scored against its own T
        candidates, max over the 32 tokens; no division by temp."""
        tok = self._tokenize(text)
        R, n = refereence_embeds.shape[0], target_embeds.shape[0]
        T = n // R if R > 1 else n
        fusion = self.encode_query(refereence_embeds, tok.input_ids, tok.attention_mask)
        rows = torch.arange(R * T, device=self._device, dtype=torch.int32).view(R, T)
        return self.gather_scores(fusion, self._gallery_bf16(target_embeds), rows).reshape(-1)


MODEL_REGISTRY = {"blip2_cir_align_prompt": Blip2QformerCirAlignPrompt, "blip2_cir_rerank": Blip2QformerCirRerank,
                  "blip2_cir_cat": Blip2QformerCirCat}


def load_model_and_preprocess(name, model_type, is_eval=False, device="cpu", **kw):
    """lavis/models/__init__.py:204-249 -> (model, vis_processors, txt_processors).  Only the caption
    processor is used by the retrieval scripts (`txt_processors["eval"]`, validate_blip.py:183,389)."""
    if name not in MODEL_REGISTRY:
        raise KeyError(f"model {name!r} is outside the scope of this library (have {sorted(MODEL_REGISTRY)})")
    model = MODEL_REGISTRY[name].from_pretrained(model_type, device=device, **kw)
    if is_eval:
        model.eval()
    txt = {"train": BlipCaptionProcessor(), "eval": BlipCaptionProcessor()}
    vis = {"train": None, "eval": None}
    return model, vis, txt


def sequence_to_list(x: Sequence) -> List:
    return list(x)
