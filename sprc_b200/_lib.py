"""ctypes binding of libsprc_b200.so (the C ABI declared in include/sprc_b200.h).

The product path has no fallback: if the shared library is missing or a call fails, a
`SprcError` is raised.  Only plain pointers and integers cross this boundary; torch is
used on the Python side purely to own device memory and streams.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsprc_b200.so")


class SprcError(RuntimeError):
    pass


class SprcConfig(ctypes.Structure):
    _fields_ = [
        ("vit_kind", c_int),
        ("vit_depth", c_int),
        ("qf_layers", c_int),
        ("max_images", c_int),
        ("max_queries", c_int),
        ("max_pairs", c_int),
        ("device", c_int),
        ("act_dtype", c_int),
    ]


class SprcTensorDesc(ctypes.Structure):
    _fields_ = [
        ("name", c_char_p),
        ("dtype", c_int),
        ("ndim", c_int),
        ("shape", c_int64 * 4),
        ("data", c_void_p),
    ]


F32, F16, BF16, I64, I32 = 0, 1, 2, 3, 4
VIT_EVA_G, VIT_CLIP_L = 0, 1
ACT_NONE, ACT_GELU, ACT_QUICKGELU = 0, 1, 2

# name -> (restype, argtypes); every symbol include/sprc_b200.h declares
SIGNATURES = {
    "sprc_abi_version": (c_int, []),
    "sprc_last_error": (c_char_p, []),
    "sprc_create": (c_int, [POINTER(SprcConfig), POINTER(c_void_p)]),
    "sprc_destroy": (None, [c_void_p]),
    "sprc_load_weights": (c_int, [c_void_p, POINTER(SprcTensorDesc), c_int, POINTER(c_int)]),
    "sprc_missing_weight": (c_char_p, [c_void_p, c_int]),
    "sprc_encode_gallery": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sprc_encode_query": (
        c_int,
        [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p],
    ),
    "sprc_encode_query_lens": (
        c_int,
        [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p],
    ),
    "sprc_sim_topk": (
        c_int,
        [c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    ),
    "sprc_sim_topk_grouped": (
        c_int,
        [c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int, c_int64, c_void_p],
    ),
    "sprc_topk_merge": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "sprc_topk_merge_packed": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "sprc_gather_scores": (
        c_int,
        [c_void_p, c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p],
    ),
    "sprc_rerank": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p],
    ),
    "sprc_rerank_lens": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p],
    ),
    "sprc_query_topk_host": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
         c_void_p],
    ),
    "sprc_query_topk_host_submit": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
         c_void_p],
    ),
    "sprc_query_topk_host_wait": (c_int, [c_void_p]),
    "sprc_query_topk_strings_submit": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_char_p, c_void_p, c_int, c_int, c_int, c_void_p,
         c_void_p, c_void_p],
    ),
    "sprc_launch_count": (c_int64, []),
    "sprc_set_act_dtype": (c_int, [c_int]),
    "sprc_profile": (c_int, [c_int]),
    "sprc_profile_read": (c_int, [POINTER(ctypes.c_double), c_int]),
    "sprc_profile_dump": (c_int, [c_char_p]),
    "sprc_op_gemm": (
        c_int,
        [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
         c_void_p, c_int, c_int, c_int, c_void_p],
    ),
    "sprc_preprocess_targetpad": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    ),
    "sprc_op_gemm2w": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
         c_int, c_void_p],
    ),
    "sprc_op_attention_pairs": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
         c_int64, c_int64, c_float, c_void_p],
    ),
    "sprc_png_decode_files": (c_int, [c_char_p, c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "sprc_op_inflate_zlib": (c_int, [c_char_p, c_int64, c_void_p, c_int64]),
    "sprc_tokenizer_create": (c_int, [c_char_p, c_int64, POINTER(c_void_p)]),
    "sprc_tokenizer_destroy": (None, [c_void_p]),
    "sprc_tokenize_host": (c_int, [c_void_p, c_char_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                   c_void_p]),
    "sprc_op_attention_ragged": (
        c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_float, c_void_p]),
    "sprc_op_layernorm": (
        c_int,
        [c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p],
    ),
    "sprc_op_attention": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
         c_int, c_int, c_void_p, c_float, c_void_p],
    ),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the library (once) and bind every declared symbol; raise SprcError if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SprcError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for this path)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.sprc_abi_version() != 1:
        raise SprcError(f"ABI version mismatch: library reports {lib.sprc_abi_version()}, binding expects 1")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().sprc_last_error()
        raise SprcError(f"libsprc_b200 call failed (code {rc}): {msg.decode() if msg else '?'}")


def ptr(t) -> c_void_p:
    """Device/host pointer of a torch tensor (or None)."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def cur_stream() -> c_void_p:
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)
