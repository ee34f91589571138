"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference model classes from /root/reference.

Only `tests/`, `oracle/make_golden.py`, `__graft_entry__.smoke()` and `bench.py`'s reference /
cpu_baseline legs may import this.  It never ships on the product path and it only works where
`/root/reference` (or a staged copy, `SPRC_REFERENCE_SRC`) exists — i.e. in the build container;
the GPU box uses the committed golden vectors and `oracle/restatement.py` instead.

The reference is pure Python on top of LAVIS; its own `lavis/__init__.py` imports the whole LAVIS
zoo and several packages that are absent here (timm, omegaconf, iopath, fairscale) and it pins
transformers 4.36 while the container has 5.x.  The shims below stub
exactly those imports and the three network touch points (tokenizer, BERT config/weights download,
ViT weight download); no reference arithmetic is replaced.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from functools import partial

import torch
import torch.nn as nn

REFERENCE_SRC = os.environ.get("SPRC_REFERENCE_SRC", "/root/reference/src")

_loaded = {}


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "lavis", "models", "blip2_models"))


def _mod(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


class _FakeTokenizer:
    """Stands in for BertTokenizer('bert-base-uncased') + [DEC] (blip2.py:30-34): only `len()` is
    needed at construction time (resize_token_embeddings, align_prompt.py:73).  Callers that need
    token ids pass them through `TokenBatch` (see `call_with_ids`)."""

    def __len__(self):
        return 30523

    def __call__(self, text, **kw):
        if isinstance(text, TokenBatch):
            return text
        raise RuntimeError("offline oracle: pass a TokenBatch (ids, mask) instead of strings")


class TokenBatch:
    """Duck-types the BatchEncoding the reference reads (`.input_ids`, `.attention_mask`, `.to`)."""

    def __init__(self, input_ids, attention_mask):
        self.input_ids = input_ids
        self.attention_mask = attention_mask

    def to(self, device):
        return TokenBatch(self.input_ids.to(device), self.attention_mask.to(device))


def _install_shims():
    if "done" in _loaded:
        return
    import transformers  # noqa: F401  (must be imported BEFORE timm is stubbed)
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu

    if "timm" not in sys.modules:
        _mod("timm")
        _mod("timm.models")
        _mod(
            "timm.models.layers",
            drop_path=lambda x, p=0.0, training=False: x,
            to_2tuple=lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x),
            trunc_normal_=lambda t, mean=0.0, std=1.0, a=-2.0, b=2.0: torch.nn.init.trunc_normal_(
                t, mean=mean, std=std, a=a, b=b
            ),
        )
        _mod("timm.models.registry", register_model=lambda f: f)

        def _no_net(*a, **k):
            raise RuntimeError("offline oracle: network download requested")

        _mod("timm.models.hub", get_cache_dir=lambda *a, **k: "/tmp/sprc_cache", download_cached_file=_no_net)
    if "fairscale" not in sys.modules:
        _mod("fairscale")
        _mod("fairscale.nn")
        _mod("fairscale.nn.checkpoint")
        _mod("fairscale.nn.checkpoint.checkpoint_activations", checkpoint_wrapper=lambda m, *a, **k: m)
    if "omegaconf" not in sys.modules:
        import yaml

        class OmegaConf:
            @staticmethod
            def load(path):
                with open(path) as f:
                    return yaml.safe_load(f)

            @staticmethod
            def create(*a, **k):
                return {}

        _mod("omegaconf", OmegaConf=OmegaConf)
    if "iopath" not in sys.modules:
        _mod("iopath")
        _mod("iopath.common")
        _mod("iopath.common.download", download=None)
        _mod("iopath.common.file_io", file_lock=None, g_pathmgr=None)

    # transformers 5.x compatibility for Qformer.py:39-44,703,943
    if not hasattr(mu, "apply_chunking_to_forward"):
        mu.apply_chunking_to_forward = pu.apply_chunking_to_forward
    if not hasattr(mu, "prune_linear_layer"):
        mu.prune_linear_layer = pu.prune_linear_layer
    if not hasattr(mu, "find_pruneable_heads_and_indices"):
        def _no_prune(*a, **k):
            raise NotImplementedError

        mu.find_pruneable_heads_and_indices = _no_prune
    _orig_init = mu.PreTrainedModel.init_weights

    def _init_weights(self):
        if not hasattr(self, "all_tied_weights_keys"):
            return self.post_init()
        return _orig_init(self)

    mu.PreTrainedModel.init_weights = _init_weights
    if not hasattr(mu.PreTrainedModel, "get_head_mask"):
        mu.PreTrainedModel.get_head_mask = lambda self, head_mask, n, is_attention_chunked=False: [None] * n

    for pkg in ["lavis", "lavis.common", "lavis.models", "lavis.models.blip2_models", "lavis.models.blip_models",
                "lavis.processors"]:
        _mod(pkg).__path__ = [os.path.join(REFERENCE_SRC, *pkg.split("."))]
    from lavis.common.registry import registry

    registry.register_path("library_root", os.path.join(REFERENCE_SRC, "lavis"))
    sys.modules["lavis.models"].BaseModel = importlib.import_module("lavis.models.base_model").BaseModel
    _loaded["done"] = True


def _patch_factories(B2, eva_vit, clip_vit, Qformer_mod, vit_depth=None, qf_layers=None):
    from transformers.models.bert.configuration_bert import BertConfig

    B2.Blip2Base.init_tokenizer = classmethod(lambda cls, truncation_side="right": _FakeTokenizer())

    def init_qformer(cls, num_query_token, vision_width, cross_attention_freq=2):
        cfg = BertConfig()  # identical to bert-base-uncased's config
        if qf_layers:
            cfg.num_hidden_layers = qf_layers
        cfg.encoder_width = vision_width
        cfg.add_cross_attention = True
        cfg.cross_attention_freq = cross_attention_freq
        cfg.query_length = num_query_token
        q = Qformer_mod.BertLMHeadModel(cfg)
        qt = nn.Parameter(torch.zeros(1, num_query_token, cfg.hidden_size))
        qt.data.normal_(mean=0.0, std=cfg.initializer_range)
        return q, qt

    B2.Blip2Base.init_Qformer = classmethod(init_qformer)

    def create_g(img_size=224, drop_path_rate=0.4, use_checkpoint=False, precision="fp16"):
        return eva_vit.VisionTransformer(
            img_size=img_size, patch_size=14, use_mean_pooling=False, embed_dim=1408, depth=vit_depth or 39,
            num_heads=1408 // 88, mlp_ratio=4.3637, qkv_bias=True, drop_path_rate=drop_path_rate,
            norm_layer=partial(nn.LayerNorm, eps=1e-6), use_checkpoint=use_checkpoint)

    def create_l(img_size=224, use_checkpoint=False, precision="fp16"):
        return clip_vit.VisionTransformer(input_resolution=img_size, patch_size=14, width=1024,
                                          layers=vit_depth or 23, heads=16, use_grad_checkpointing=use_checkpoint)

    B2.create_eva_vit_g = create_g
    B2.create_clip_vit_L = create_l


def load_reference_class(kind="align_prompt"):
    """Return the reference's model class (unmodified source, executed from REFERENCE_SRC)."""
    if not available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_SRC}")
    _install_shims()
    if kind == "align_prompt":
        m = importlib.import_module("lavis.models.blip2_models.blip2_qformer_cir_align_prompt")
        return m.Blip2QformerCirAlignPrompt
    if kind == "rerank":
        m = importlib.import_module("lavis.models.blip2_models.blip2_qformer_cir_rerank")
        return m.Blip2QformerCirRerank
    if kind == "cat":
        if "skimage" not in sys.modules:   # imported at blip2_qformer_cir_cat.py:23, never used
            _mod("skimage", transform=_mod("skimage.transform"))
        m = importlib.import_module("lavis.models.blip2_models.blip2_qformer_cir_cat")
        return m.Blip2QformerCirCat
    raise ValueError(kind)


def build_reference_model(vit="clip_L", seed=0, vit_depth=None, qf_layers=None, kind="align_prompt"):
    """Construct the reference model with its own initialisers under a fixed seed, fp32, eval mode
    (the oracle must run with dropout off)."""
    cls = load_reference_class(kind)
    B2 = importlib.import_module("lavis.models.blip2_models.blip2")
    eva_vit = importlib.import_module("lavis.models.eva_vit")
    clip_vit = importlib.import_module("lavis.models.clip_vit")
    Qf = importlib.import_module("lavis.models.blip2_models.Qformer")
    _patch_factories(B2, eva_vit, clip_vit, Qf, vit_depth, qf_layers)
    torch.manual_seed(seed)
    model = cls(vit_model=vit, vit_precision="fp32")
    return model.float().eval()


def call_inference(model, reference_embeds, target_feats, input_ids, attention_mask):
    """model.inference(...) with pre-tokenised text (blip2_qformer_cir_align_prompt.py:312-361)."""
    return model.inference(reference_embeds, target_feats, TokenBatch(input_ids, attention_mask))


class no_cuda_moves:
    """`Blip2QformerCirCat.inference` moves its inputs with `.cuda()` (blip2_qformer_cir_cat.py:283-284); the oracle
    runs the reference on the CPU, so inside this context `Tensor.cuda()` returns the tensor unchanged."""

    def __enter__(self):
        self._orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda t, *a, **k: t
        return self

    def __exit__(self, *exc):
        torch.Tensor.cuda = self._orig
        return False


def load_caption_processor():
    """The reference's BlipCaptionProcessor (processors/blip_processors.py:28-68)."""
    _install_shims()
    _mod("lavis.processors.randaugment", RandomAugment=object)
    base = importlib.import_module("lavis.processors.base_processor")
    sys.modules["lavis.processors"].BaseProcessor = base.BaseProcessor  # registry.py:124 imports it from here
    bp = importlib.import_module("lavis.processors.blip_processors")
    return bp.BlipCaptionProcessor
