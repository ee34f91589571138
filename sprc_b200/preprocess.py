"""GPU image preprocessing for gallery indexing.

Replaces the reference's per-image PIL pipeline `targetpad_transform(target_ratio, dim)` (src/data_utils.py:52-72
TargetPad, :91-105 Compose[TargetPad, Resize(dim, BICUBIC), CenterCrop(dim), _convert_image_to_rgb, ToTensor,
Normalize]) for decoded RGB uint8 images with two CUDA kernels (csrc/preprocess.cu), bit-exact with the
PIL/torchvision result.  The host part below computes, per distinct image size, the geometry (pad, resized size,
crop window) exactly as TargetPad / torchvision.transforms.functional.resize / center_crop do, and Pillow's fixed-point
bicubic coefficient tables (libImaging/Resample.c: precompute_coeffs + normalize_coeffs_8bpc) in double precision
with the same operation order; only the taps of the 224 columns / rows that survive the centre crop are emitted.

Decoding stays on the host.  `PngBatchDecoder` (C++ workers behind `sprc_png_decode_files`, csrc/png.cpp) decodes a
batch of PNG files — the format both datasets are stored in (data_utils.py:167-186,253-270) — into a pinned arena that
the resize kernels consume after one copy; `PngIndexFeeder` chains decoder and preprocessor.  The reference resizes
BEFORE converting to RGB, which equals convert-then-resize only for images Pillow opens as "RGB" or "L"; every other
file (palette, alpha, 1-bit, 16-bit, interlaced, not a PNG) goes through the reference's own PIL chain on the host.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L

PRECISION_BITS = 32 - 8 - 2
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)   # data_utils.py:104
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
_FIELDS = 16


def targetpad_geometry(w: int, h: int, target_ratio: float, dim: int):
    """(hp, vp, padded_w, padded_h, out_w, out_h, crop_left, crop_top) of data_utils.py:63-72 + torchvision's
    Resize(int) (`_compute_resized_output_size`) + CenterCrop(dim)."""
    actual_ratio = max(w, h) / min(w, h)
    hp = vp = 0
    if not actual_ratio < target_ratio:
        scaled_max_wh = max(w, h) / target_ratio
        hp = max(int((scaled_max_wh - w) / 2), 0)
        vp = max(int((scaled_max_wh - h) / 2), 0)
    pw, ph = w + 2 * hp, h + 2 * vp
    short, long_ = (pw, ph) if pw <= ph else (ph, pw)
    new_short, new_long = dim, int(dim * long_ / short)
    ow, oh = (new_short, new_long) if pw <= ph else (new_long, new_short)
    crop_top = int(round((oh - dim) / 2.0))
    crop_left = int(round((ow - dim) / 2.0))
    return hp, vp, pw, ph, ow, oh, crop_left, crop_top


def _bicubic(x: np.ndarray) -> np.ndarray:
    """Resample.c bicubic_filter with a = -0.5 (same expression order)."""
    a = -0.5
    x = np.abs(x)
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


def resample_coeffs(in_size: int, out_size: int, first: int, count: int) -> Tuple[np.ndarray, np.ndarray]:
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for output positions [first, first + count) of an
    in_size -> out_size bicubic resample.  Returns (bounds int32 [count,2] = (xmin, n), coeffs int32 [count, ksize]).
    in_size == out_size (Pillow skips the pass) yields the identity tap."""
    if in_size == out_size:
        bounds = np.stack([np.arange(first, first + count), np.ones(count, dtype=np.int64)], axis=1)
        return bounds.astype(np.int32), np.full((count, 1), 1 << PRECISION_BITS, dtype=np.int32)
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xx = np.arange(first, first + count, dtype=np.float64)
    center = (xx + 0.5) * scale
    ss = 1.0 / filterscale
    xmin = np.trunc(center - support + 0.5).astype(np.int64)
    xmin = np.maximum(xmin, 0)
    xmax = np.trunc(center + support + 0.5).astype(np.int64)
    xmax = np.minimum(xmax, in_size) - xmin
    j = np.arange(ksize, dtype=np.int64)[None, :]
    valid = j < xmax[:, None]
    w = _bicubic((j + xmin[:, None] - center[:, None] + 0.5) * ss)
    w = np.where(valid, w, 0.0)
    ww = np.cumsum(w, axis=1)[:, -1:]            # sequential accumulation, as the C loop does
    w = np.where(ww != 0.0, w / np.where(ww != 0.0, ww, 1.0), w)
    k = np.where(w < 0, np.trunc(-0.5 + w * (1 << PRECISION_BITS)), np.trunc(0.5 + w * (1 << PRECISION_BITS)))
    k = np.where(valid, k, 0.0).astype(np.int32)
    return np.stack([xmin, xmax], axis=1).astype(np.int32), k


class TargetPadPreprocessor:
    """`pre(images) -> float32 [n,3,dim,dim]` on `device` for a list of RGB uint8 arrays [H,W,3] (or PIL RGB images)."""

    def __init__(self, target_ratio: float = 1.25, dim: int = 224, device="cuda:0", mean=CLIP_MEAN, std=CLIP_STD):
        self.target_ratio, self.dim = float(target_ratio), int(dim)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("TargetPadPreprocessor runs on a CUDA device (there is no CPU path)")
        self._lib = L.load()
        self._mean = (L.c_float * 3)(*mean)
        self._std = (L.c_float * 3)(*std)
        self._plans: Dict[Tuple[int, int], dict] = {}

    def plan(self, w: int, h: int) -> dict:
        """Geometry + coefficient tables of one image size (cached)."""
        key = (w, h)
        p = self._plans.get(key)
        if p is None:
            dim = self.dim
            hp, vp, pw, ph, ow, oh, cl, ct = targetpad_geometry(w, h, self.target_ratio, dim)
            hb, hk = resample_coeffs(pw, ow, cl, dim)
            vb, vk = resample_coeffs(ph, oh, ct, dim)
            row0 = int(vb[:, 0].min())
            nrows = int((vb[:, 0] + vb[:, 1]).max()) - row0
            vb = vb.copy()
            vb[:, 0] -= row0
            p = dict(hp=hp, vp=vp, row0=row0, nrows=nrows, hb=hb, hk=hk, vb=vb, vk=vk)
            self._plans[key] = p
        return p

    @torch.no_grad()
    def __call__(self, images: Sequence) -> torch.Tensor:
        arrs: List[np.ndarray] = []
        for im in images:
            if not isinstance(im, np.ndarray):
                if getattr(im, "mode", "RGB") != "RGB":
                    raise NotImplementedError(f"image mode {im.mode!r}: the GPU preprocessor takes RGB images")
                im = np.asarray(im)
            if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3:
                raise ValueError(f"expected uint8 [H,W,3], got {im.dtype} {im.shape}")
            arrs.append(np.ascontiguousarray(im))
        if not arrs:
            return torch.empty(0, 3, self.dim, self.dim, device=self.device)
        pix = torch.from_numpy(np.concatenate([a.reshape(-1) for a in arrs])).pin_memory()
        offs, o = [], 0
        for a in arrs:
            offs.append(o)
            o += a.size
        return self.run_packed(pix, [(a.shape[1], a.shape[0]) for a in arrs], offs)

    @torch.no_grad()
    def run_packed(self, pix: torch.Tensor, sizes: Sequence[Tuple[int, int]], offsets: Sequence[int]) -> torch.Tensor:
        """`pix`: uint8 buffer (pinned host or device) holding image i as packed RGB at byte `offsets[i]`, `sizes[i]` =
        (width, height).  One H2D copy of the used prefix, then the two kernels.  Returns fp32 [n,3,dim,dim]."""
        n, dim = len(sizes), self.dim
        if n == 0:
            return torch.empty(0, 3, dim, dim, device=self.device)
        desc = np.zeros((n, _FIELDS), dtype=np.int64)
        tables: List[np.ndarray] = []
        table_off: Dict[Tuple[int, int], Tuple[int, int, int, int]] = {}
        t_len = 0
        tmp_off = 0
        max_rows = 0
        end = 0
        for i, (w, h) in enumerate(sizes):
            src_off = int(offsets[i])
            end = max(end, src_off + 3 * w * h)
            p = self.plan(w, h)
            if (w, h) not in table_off:   # images of one size share one set of tables
                offs = []
                for t in (p["hb"], p["hk"], p["vb"], p["vk"]):
                    offs.append(t_len)
                    tables.append(t.reshape(-1))
                    t_len += t.size
                table_off[(w, h)] = tuple(offs)
            o = table_off[(w, h)]
            desc[i, :14] = (src_off, w, h, p["hp"], p["vp"], p["row0"], p["nrows"], p["hk"].shape[1], p["vk"].shape[1],
                            o[0], o[1], o[2], o[3], tmp_off)
            tmp_off += p["nrows"] * dim * 3
            max_rows = max(max_rows, p["nrows"])
        dev = self.device
        pix_d = pix[:end].to(dev, non_blocking=True)
        desc_d = torch.from_numpy(desc).to(dev, non_blocking=True)
        tab_d = torch.from_numpy(np.concatenate(tables).astype(np.int32)).to(dev, non_blocking=True)
        tmp = torch.empty(tmp_off, dtype=torch.uint8, device=dev)
        out = torch.empty(n, 3, dim, dim, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            L.check(self._lib.sprc_preprocess_targetpad(L.ptr(pix_d), L.ptr(desc_d), L.ptr(tab_d), n, dim, max_rows,
                                                        L.ptr(tmp), self._mean, self._std, L.ptr(out),
                                                        L.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        return out


PIL_MODES = ("RGB", "L", "1", "P", "LA", "RGBA")


class PngBatch:
    """Result of `PngBatchDecoder.decode`: `arena` (pinned uint8) holds image i as packed RGB at `offsets[i]`;
    `wh[i]` = (width, height, Pillow mode code, index into PIL_MODES); `status[i]`: 0 decoded, 1 not taken (decode with
    Pillow), 2 corrupt, 3 unreadable."""

    def __init__(self, paths, arena, offsets, wh, status, slot=0):
        self.paths, self.arena, self.offsets, self.wh, self.status, self.slot = paths, arena, offsets, wh, status, slot

    def image(self, i: int) -> np.ndarray:
        """uint8 [H,W,3] view of image i (status 0 only)."""
        w, h = int(self.wh[i, 0]), int(self.wh[i, 1])
        o = int(self.offsets[i])
        return self.arena[o:o + 3 * w * h].numpy().reshape(h, w, 3)


class PngBatchDecoder:
    """Threaded PNG -> RGB8 decode into pinned arenas (sprc_png_decode_files).  Two arenas alternate so that batch i+1
    can be decoded while batch i is still being copied to the device; `decode` grows an arena when a batch needs more."""

    def __init__(self, threads: int = 0, arena_bytes: int = 64 << 20, pin: bool = True):
        self._lib = L.load()
        self.threads = int(threads)
        self._pin = bool(pin) and torch.cuda.is_available()
        self._arenas = [self._alloc(arena_bytes), self._alloc(arena_bytes)]
        self._copied = [None, None]   # CUDA event per arena: the last H2D copy that read it
        self._copy_stream = None
        self._turn = 0

    def to_device(self, batch: "PngBatch", nbytes: int, device) -> torch.Tensor:
        """The used prefix of the batch's arena on `device`, copied on a side stream (so the copy is not queued behind
        the encoder work of earlier batches) and ordered before whatever the current stream does next.  The arena is
        not decoded into again before that copy has run."""
        dev = torch.device(device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream(dev)
        with torch.cuda.stream(self._copy_stream):
            pix_d = batch.arena[:nbytes].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        cur.wait_event(ev)
        pix_d.record_stream(cur)
        self._copied[batch.slot] = ev
        return pix_d

    def _alloc(self, nbytes: int) -> torch.Tensor:
        t = torch.empty(int(nbytes), dtype=torch.uint8)
        return t.pin_memory() if self._pin else t

    def decode(self, paths: Sequence[str]) -> PngBatch:
        n = len(paths)
        enc = [str(p).encode("utf-8") for p in paths]
        poffs = np.zeros(n + 1, dtype=np.int64)
        np.cumsum([len(e) for e in enc], out=poffs[1:])
        blob = b"".join(enc)
        offs = np.zeros(n + 1, dtype=np.int64)
        wh = np.zeros((n, 3), dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        slot = self._turn
        self._turn ^= 1
        if self._copied[slot] is not None:
            self._copied[slot].synchronize()
            self._copied[slot] = None
        for _ in range(2):
            arena = self._arenas[slot]
            rc = self._lib.sprc_png_decode_files(blob, poffs.ctypes.data, n, self.threads, L.ptr(arena), arena.numel(),
                                                 offs.ctypes.data, wh.ctypes.data, status.ctypes.data)
            if rc != -34:
                break
            self._arenas[slot] = self._alloc(max(int(offs[n]), 2 * arena.numel()))   # grow once, decode again
        L.check(rc)
        return PngBatch(list(paths), self._arenas[slot], offs, wh, status, slot)


class PngIndexFeeder:
    """paths -> fp32 [m,3,dim,dim] on the device, the tensor `targetpad_transform` + default_collate would have produced
    for the readable images of the batch, plus the indices that were kept.  Files Pillow opens as "RGB" / "L" take the
    native decode + GPU resize; the rest run `fallback(PIL.Image.open(path))` — the reference's own transform — on the
    host; files that raise there are dropped, as the reference's datasets + collate_fn do (data_utils.py:191-192,
    utils.py:141-148)."""

    def __init__(self, pre: TargetPadPreprocessor, fallback=None, threads: int = 0):
        self.pre = pre
        self.decoder = PngBatchDecoder(threads=threads)
        self.fallback = fallback
        self.n_native = self.n_fallback = self.n_dropped = 0

    def decode(self, paths: Sequence[str]) -> PngBatch:
        """Host half (safe to run on a helper thread while the GPU works on the previous batch)."""
        return self.decoder.decode(paths)

    @torch.no_grad()
    def finish(self, batch: PngBatch):
        import PIL.Image

        n = len(batch.paths)
        native = [i for i in range(n) if batch.status[i] == 0 and batch.wh[i, 2] in (0, 1)]
        sizes = [(int(batch.wh[i, 0]), int(batch.wh[i, 1])) for i in native]
        offs = [int(batch.offsets[i]) for i in native]
        end = max([o + 3 * w * h for o, (w, h) in zip(offs, sizes)], default=0)
        pix = self.decoder.to_device(batch, end, self.pre.device) if end else batch.arena
        out_native = self.pre.run_packed(pix, sizes, offs)
        host = {}
        for i in range(n):
            if i in native:
                continue
            try:
                if self.fallback is None:
                    raise NotImplementedError(f"{batch.paths[i]}: not an RGB / L PNG and no PIL fallback transform given")
                host[i] = self.fallback(PIL.Image.open(batch.paths[i]))
            except NotImplementedError:
                raise
            except Exception as e:  # noqa: BLE001  the reference prints and drops (data_utils.py:191-192)
                print(f"Exception: {e}")
        self.n_native += len(native)
        self.n_fallback += len(host)
        self.n_dropped += n - len(native) - len(host)
        if not host:
            return out_native, native
        keep = sorted(native + list(host))
        pos = {i: j for j, i in enumerate(keep)}
        out = torch.empty(len(keep), 3, self.pre.dim, self.pre.dim, device=self.pre.device)
        if native:
            out[torch.tensor([pos[i] for i in native], device=out.device)] = out_native
        for i, t in host.items():
            out[pos[i]] = t.to(out.device, torch.float32)
        return out, keep

    def __call__(self, paths: Sequence[str]):
        return self.finish(self.decode(paths))
