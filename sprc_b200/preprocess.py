"""GPU image preprocessing for gallery indexing.

Replaces the reference's per-image PIL pipeline `targetpad_transform(target_ratio, dim)` (src/data_utils.py:52-72
TargetPad, :91-105 Compose[TargetPad, Resize(dim, BICUBIC), CenterCrop(dim), _convert_image_to_rgb, ToTensor,
Normalize]) for decoded RGB uint8 images with two CUDA kernels (csrc/preprocess.cu), bit-exact with the
PIL/torchvision result.  The host part below computes, per distinct image size, the geometry (pad, resized size,
crop window) exactly as TargetPad / torchvision.transforms.functional.resize / center_crop do, and Pillow's fixed-point
bicubic coefficient tables (libImaging/Resample.c: precompute_coeffs + normalize_coeffs_8bpc) in double precision
with the same operation order; only the taps of the 224 columns / rows that survive the centre crop are emitted.

Decoding (PNG/JPEG -> RGB uint8) stays on the host (PIL, thread pool); images must be mode "RGB" — the reference
resizes before converting to RGB, which for other modes (palette, greyscale) is a different operation.
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
        n, dim = len(arrs), self.dim
        if n == 0:
            return torch.empty(0, 3, dim, dim, device=self.device)
        desc = np.zeros((n, _FIELDS), dtype=np.int64)
        tables: List[np.ndarray] = []
        table_off: Dict[Tuple[int, int], Tuple[int, int, int, int]] = {}
        t_len = 0
        src_off = tmp_off = 0
        max_rows = 0
        for i, a in enumerate(arrs):
            h, w = a.shape[:2]
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
            src_off += a.size
            tmp_off += p["nrows"] * dim * 3
            max_rows = max(max_rows, p["nrows"])
        pix = torch.from_numpy(np.concatenate([a.reshape(-1) for a in arrs])).pin_memory()
        dev = self.device
        pix_d = pix.to(dev, non_blocking=True)
        desc_d = torch.from_numpy(desc).to(dev, non_blocking=True)
        tab_d = torch.from_numpy(np.concatenate(tables).astype(np.int32)).to(dev, non_blocking=True)
        tmp = torch.empty(tmp_off, dtype=torch.uint8, device=dev)
        out = torch.empty(n, 3, dim, dim, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            L.check(self._lib.sprc_preprocess_targetpad(L.ptr(pix_d), L.ptr(desc_d), L.ptr(tab_d), n, dim, max_rows,
                                                        L.ptr(tmp), self._mean, self._std, L.ptr(out),
                                                        L.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        return out
