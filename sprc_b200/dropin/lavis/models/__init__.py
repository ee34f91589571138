"""`from lavis.models import load_model_and_preprocess` (lavis/models/__init__.py:204-249 in the reference),
served by the B200-native implementation.  SPRC_VIT_DEPTH / SPRC_QF_LAYERS / SPRC_MAX_IMAGES /
SPRC_MAX_QUERIES / SPRC_MAX_PAIRS (optional environment overrides) size the handle; depth overrides build
truncated models for tests."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from sprc_b200.model import (Blip2QformerCirAlignPrompt, Blip2QformerCirRerank,  # noqa: E402,F401
                             load_model_and_preprocess as _load)


def _env_int(name, default):
    v = os.environ.get(name)
    return int(v) if v else default


def load_model_and_preprocess(name, model_type, is_eval=False, device="cpu"):
    kw = dict(max_images=_env_int("SPRC_MAX_IMAGES", 64), max_queries=_env_int("SPRC_MAX_QUERIES", 64),
              vit_depth=_env_int("SPRC_VIT_DEPTH", 0), qf_layers=_env_int("SPRC_QF_LAYERS", 0))
    if name == "blip2_cir_rerank" or os.environ.get("SPRC_MAX_PAIRS"):
        kw["max_pairs"] = _env_int("SPRC_MAX_PAIRS", 2500)
    return _load(name, model_type, is_eval=is_eval, device=device, **kw)
