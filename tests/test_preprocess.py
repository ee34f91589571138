"""GPU image preprocessing == the reference's PIL pipeline `targetpad_transform(1.25, 224)`
(src/data_utils.py:52-72, 91-105), bit for bit.

The checker is the real thing: PIL + torchvision run the reference's Compose (restated below from data_utils.py — the
reference module itself is not importable on the GPU box).  CPU tests validate the host logic (geometry and Pillow's
fixed-point coefficient tables) by applying the tables with numpy; GPU tests run the CUDA kernels through the C ABI.
"""
import numpy as np
import PIL.Image
import pytest
import torch
import torchvision.transforms.functional as F
from torchvision.transforms import CenterCrop, Compose, Normalize, Resize, ToTensor

from sprc_b200 import preprocess as P


class TargetPad:  # data_utils.py:52-72
    def __init__(self, target_ratio, size):
        self.size, self.target_ratio = size, target_ratio

    def __call__(self, image):
        w, h = image.size
        actual_ratio = max(w, h) / min(w, h)
        if actual_ratio < self.target_ratio:
            return image
        scaled_max_wh = max(w, h) / self.target_ratio
        hp = max(int((scaled_max_wh - w) / 2), 0)
        vp = max(int((scaled_max_wh - h) / 2), 0)
        return F.pad(image, [hp, vp, hp, vp], 0, "constant")


def reference_transform(target_ratio=1.25, dim=224):  # data_utils.py:91-105
    return Compose([TargetPad(target_ratio, dim), Resize(dim, interpolation=PIL.Image.BICUBIC), CenterCrop(dim),
                    lambda im: im.convert("RGB"), ToTensor(),
                    Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))])


SIZES = [(240, 200), (200, 240), (224, 224), (640, 480), (375, 500), (130, 100), (100, 331), (1000, 300), (224, 300),
         (300, 224), (257, 256), (59, 47), (512, 512), (1280, 720), (333, 1000)]  # (w, h)


def _images(seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for w, h in SIZES:
        a = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        a[: h // 3] = (a[: h // 3].astype(np.int32) * 3 // 4 + 64).astype(np.uint8)   # some structure, full range elsewhere
        out.append(a)
    return out


def _apply_tables_numpy(a, plan, dim):
    """The two integer passes of csrc/preprocess.cu in numpy (test helper)."""
    h, w = a.shape[:2]
    pad = np.zeros((h + 2 * plan["vp"], w + 2 * plan["hp"], 3), dtype=np.int64)
    pad[plan["vp"]:plan["vp"] + h, plan["hp"]:plan["hp"] + w] = a
    rows = pad[plan["row0"]:plan["row0"] + plan["nrows"]]
    tmp = np.empty((plan["nrows"], dim, 3), dtype=np.int64)
    for x in range(dim):
        x0, n = plan["hb"][x]
        acc = (rows[:, x0:x0 + n] * plan["hk"][x, :n].astype(np.int64)[None, :, None]).sum(1) + (1 << 21)
        tmp[:, x] = np.clip(acc >> 22, 0, 255)
    out = np.empty((dim, dim, 3), dtype=np.uint8)
    for y in range(dim):
        y0, n = plan["vb"][y]
        acc = (tmp[y0:y0 + n] * plan["vk"][y, :n].astype(np.int64)[:, None, None]).sum(0) + (1 << 21)
        out[y] = np.clip(acc >> 22, 0, 255)
    return out


class _HostOnly(P.TargetPadPreprocessor):
    def __init__(self, target_ratio=1.25, dim=224):  # planner without a device / library
        self.target_ratio, self.dim, self._plans = float(target_ratio), int(dim), {}


def test_host_tables_reproduce_pil_pipeline_bit_exact():
    pre = _HostOnly()
    ref = reference_transform()
    for a in _images():
        h, w = a.shape[:2]
        got_u8 = _apply_tables_numpy(a, pre.plan(w, h), 224)
        got = F.normalize(torch.from_numpy(got_u8).permute(2, 0, 1).float().div(255), P.CLIP_MEAN, P.CLIP_STD)
        want = ref(PIL.Image.fromarray(a))
        assert torch.equal(got, want), (w, h, (got - want).abs().max().item())


def test_host_tables_bit_exact_on_random_sizes():
    """120 random image sizes (1 x 1 up to 900 x 900, every aspect ratio on both sides of the TargetPad threshold):
    same bit-exact bar as above, so Pillow's support-window / coefficient rounding is pinned beyond the fixed list."""
    rng = np.random.default_rng(123)
    pre, ref = _HostOnly(), reference_transform()
    for i in range(120):
        w, h = int(rng.integers(20, 900)), int(rng.integers(20, 900))
        if i % 10 == 0:
            w, h = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        a = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        got_u8 = _apply_tables_numpy(a, pre.plan(w, h), 224)
        got = F.normalize(torch.from_numpy(got_u8).permute(2, 0, 1).float().div(255), P.CLIP_MEAN, P.CLIP_STD)
        assert torch.equal(got, ref(PIL.Image.fromarray(a))), (w, h)


def test_geometry_matches_torchvision():
    for w, h in SIZES + [(1, 1), (2000, 31), (31, 2000)]:
        im = PIL.Image.new("RGB", (w, h))
        padded = TargetPad(1.25, 224)(im)
        resized = Resize(224, interpolation=PIL.Image.BICUBIC)(padded)
        hp, vp, pw, ph, ow, oh, cl, ct = P.targetpad_geometry(w, h, 1.25, 224)
        assert (pw, ph) == padded.size and (ow, oh) == resized.size, (w, h)
        assert 0 <= cl <= ow - 224 and 0 <= ct <= oh - 224


@pytest.mark.gpu
def test_gpu_preprocess_bit_exact_vs_pil():
    pre = P.TargetPadPreprocessor(1.25, 224, device="cuda:0")
    ref = reference_transform()
    imgs = _images(1)
    out = pre(imgs + [PIL.Image.fromarray(imgs[3])])       # arrays and PIL images, mixed sizes, one batch
    torch.cuda.synchronize()
    assert out.shape == (len(imgs) + 1, 3, 224, 224)
    for i, a in enumerate(imgs):
        want = ref(PIL.Image.fromarray(a))
        assert torch.equal(out[i].cpu(), want), (a.shape, (out[i].cpu() - want).abs().max().item())
    assert torch.equal(out[-1], out[3])
    with pytest.raises(NotImplementedError):
        pre([PIL.Image.new("L", (50, 60))])


@pytest.mark.gpu
def test_gpu_preprocess_feeds_the_encoder_like_the_pil_path():
    """Same image through the PIL pipeline and through the GPU preprocessor gives identical gallery features."""
    from oracle import synth
    from sprc_b200.model import Blip2QformerCirAlignPrompt

    m = Blip2QformerCirAlignPrompt(vit_model="clip_L", device="cuda:0", max_images=8, max_queries=8, vit_depth=2,
                                   qf_layers=2)
    m.load_state_dict(synth.make_state_dict("clip_L", 2, 2, seed=0))
    imgs = _images(2)[:6]
    a = torch.stack([reference_transform()(PIL.Image.fromarray(x)) for x in imgs]).cuda()
    b = P.TargetPadPreprocessor(device="cuda:0")(imgs)
    fa, _ = m.extract_target_features(a)
    fb, _ = m.extract_target_features(b)
    assert torch.equal(fa, fb)
