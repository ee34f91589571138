"""Seeded synthetic checkpoints and inputs (SURVEY.md §8d "Synthetic inputs") shared by bench.py, the parity tests,
the golden-vector generator and the oracle (oracle/synth.py re-exports this module).  Data generation only: no part
of the retrieval path is computed here.

No dataset, checkpoint or tokenizer vocabulary exists offline, so both the reference (in the build
container) and the CUDA path (on the GPU box) are driven from the same generated state dict
(reference key names and shapes) and the same generated images / token ids.
Everything here is deterministic CPU torch RNG, so the GPU box regenerates bit-identical tensors.
"""
from __future__ import annotations

import torch

VIT_DIMS = {
    # name: (width, full depth, heads, mlp)
    "eva_clip_g": (1408, 39, 16, 6144),
    "clip_L": (1024, 23, 16, 4096),
}


def state_dict_spec(vit: str, vit_depth: int | None = None, qf_layers: int = 12, with_lm_head: bool = False):
    """[(key, shape, kind)] in checkpoint order; kind in {w, b, ln_w, ln_b, emb}."""
    Dv, full_depth, _, mlp = VIT_DIMS[vit]
    depth = vit_depth or full_depth
    spec = [("query_tokens", (1, 32, 768), "emb"), ("temp", (), "temp"), ("prompt_tokens", (1, 32, 768), "emb")]
    ve = "visual_encoder."
    if vit == "eva_clip_g":
        spec += [(ve + "cls_token", (1, 1, Dv), "emb"), (ve + "pos_embed", (1, 257, Dv), "emb"),
                 (ve + "patch_embed.proj.weight", (Dv, 3, 14, 14), "w"), (ve + "patch_embed.proj.bias", (Dv,), "b")]
        for i in range(depth):
            p = f"{ve}blocks.{i}."
            spec += [(p + "norm1.weight", (Dv,), "ln_w"), (p + "norm1.bias", (Dv,), "ln_b"),
                     (p + "attn.q_bias", (Dv,), "b"), (p + "attn.v_bias", (Dv,), "b"),
                     (p + "attn.qkv.weight", (3 * Dv, Dv), "w"),
                     (p + "attn.proj.weight", (Dv, Dv), "w"), (p + "attn.proj.bias", (Dv,), "b"),
                     (p + "norm2.weight", (Dv,), "ln_w"), (p + "norm2.bias", (Dv,), "ln_b"),
                     (p + "mlp.fc1.weight", (mlp, Dv), "w"), (p + "mlp.fc1.bias", (mlp,), "b"),
                     (p + "mlp.fc2.weight", (Dv, mlp), "w"), (p + "mlp.fc2.bias", (Dv,), "b")]
    else:
        spec += [(ve + "class_embedding", (Dv,), "emb"), (ve + "positional_embedding", (257, Dv), "emb"),
                 (ve + "conv1.weight", (Dv, 3, 14, 14), "w"),
                 (ve + "ln_pre.weight", (Dv,), "ln_w"), (ve + "ln_pre.bias", (Dv,), "ln_b")]
        for i in range(depth):
            p = f"{ve}transformer.resblocks.{i}."
            spec += [(p + "attn.in_proj_weight", (3 * Dv, Dv), "w"), (p + "attn.in_proj_bias", (3 * Dv,), "b"),
                     (p + "attn.out_proj.weight", (Dv, Dv), "w"), (p + "attn.out_proj.bias", (Dv,), "b"),
                     (p + "ln_1.weight", (Dv,), "ln_w"), (p + "ln_1.bias", (Dv,), "ln_b"),
                     (p + "mlp.c_fc.weight", (mlp, Dv), "w"), (p + "mlp.c_fc.bias", (mlp,), "b"),
                     (p + "mlp.c_proj.weight", (Dv, mlp), "w"), (p + "mlp.c_proj.bias", (Dv,), "b"),
                     (p + "ln_2.weight", (Dv,), "ln_w"), (p + "ln_2.bias", (Dv,), "ln_b")]
    spec += [("ln_vision.weight", (Dv,), "ln_w"), ("ln_vision.bias", (Dv,), "ln_b")]
    qb = "Qformer.bert."
    spec += [(qb + "embeddings.word_embeddings.weight", (30523, 768), "emb"),
             (qb + "embeddings.position_embeddings.weight", (512, 768), "emb"),
             (qb + "embeddings.LayerNorm.weight", (768,), "ln_w"), (qb + "embeddings.LayerNorm.bias", (768,), "ln_b")]
    for l in range(qf_layers):
        p = f"{qb}encoder.layer.{l}."

        def attn(prefix, kv_width):
            return [(prefix + "self.query.weight", (768, 768), "w"), (prefix + "self.query.bias", (768,), "b"),
                    (prefix + "self.key.weight", (768, kv_width), "w"), (prefix + "self.key.bias", (768,), "b"),
                    (prefix + "self.value.weight", (768, kv_width), "w"), (prefix + "self.value.bias", (768,), "b"),
                    (prefix + "output.dense.weight", (768, 768), "w"), (prefix + "output.dense.bias", (768,), "b"),
                    (prefix + "output.LayerNorm.weight", (768,), "ln_w"),
                    (prefix + "output.LayerNorm.bias", (768,), "ln_b")]

        spec += attn(p + "attention.", 768)
        if l % 2 == 0:
            spec += attn(p + "crossattention.", Dv)
        for nm in ("", "_query"):
            spec += [(p + f"intermediate{nm}.dense.weight", (3072, 768), "w"),
                     (p + f"intermediate{nm}.dense.bias", (3072,), "b"),
                     (p + f"output{nm}.dense.weight", (768, 3072), "w"), (p + f"output{nm}.dense.bias", (768,), "b"),
                     (p + f"output{nm}.LayerNorm.weight", (768,), "ln_w"),
                     (p + f"output{nm}.LayerNorm.bias", (768,), "ln_b")]
    spec += [("vision_proj.weight", (256, 768), "w"), ("vision_proj.bias", (256,), "b"),
             ("text_proj.weight", (256, 768), "w"), ("text_proj.bias", (256,), "b"),
             ("itm_head.weight", (2, 768), "w_itm"), ("itm_head.bias", (2,), "b")]
    return spec


def make_state_dict(vit: str, vit_depth: int | None = None, qf_layers: int = 12, seed: int = 0, gain: float = 1.0):
    """Seeded fp32 state dict under the reference's key names.  Scales follow the reference's own
    initialisers (std 0.02, eva_vit.py:300-315 / Qformer.py:670-680) but LayerNorm affine parameters and
    biases are perturbed away from (1, 0) so that every parameter participates in the parity check.
    `gain` multiplies every Linear / conv weight matrix: at the initialisers' scale (gain 1) the features of all
    images nearly coincide (similarities within 0.157..0.173); gain 2.5 spreads them over [-0.07, 0.13], closer to a
    trained model's, which is what a Recall@K comparison needs (tests/test_parity_gpu.py recall parity)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape, kind in state_dict_spec(vit, vit_depth, qf_layers):
        if kind == "temp":
            sd[key] = torch.tensor(0.07)
        elif kind == "w":
            sd[key] = torch.randn(shape, generator=g) * 0.02 * gain   # gain 1.0: bit-identical to the earlier data
        elif kind == "w_itm":
            sd[key] = torch.randn(shape, generator=g) * 0.2
        elif kind == "b":
            sd[key] = torch.randn(shape, generator=g) * 0.02
        elif kind == "ln_w":
            sd[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "ln_b":
            sd[key] = 0.05 * torch.randn(shape, generator=g)
        elif kind == "emb":
            sd[key] = torch.randn(shape, generator=g) * 0.02
        else:
            raise ValueError(kind)
    return sd


def make_images(n: int, seed: int = 1234):
    """randn images clamped to the post-Normalize range of data_utils.py:104."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 3, 224, 224, generator=g).clamp_(-2.2, 2.2)


def make_token_ids(n: int, seed: int = 4321):
    """[CLS] w_1..w_L [SEP] pad...  with L ~ U{3..20}, w ~ U{1000..29999}; returns (ids, mask) int64 [n,32]."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.zeros(n, 32, dtype=torch.long)
    for i in range(n):
        L = int(torch.randint(3, 21, (1,), generator=g))
        ids[i, 0] = 101
        ids[i, 1:1 + L] = torch.randint(1000, 30000, (L,), generator=g)
        ids[i, 1 + L] = 102
    return ids, (ids != 0).long()


def make_gallery_features(n: int, seed: int = 99, device="cpu", dtype=torch.float32):
    """Unit-norm randn gallery features [n,32,256] for scan-only measurements."""
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn(n, 32, 256, generator=g, device=device, dtype=torch.float32)
    return torch.nn.functional.normalize(x, dim=-1).to(dtype)


def make_dyadic(shape, seed: int, device="cpu"):
    """Entries k/16 with |k| <= 4: every 256-term dot product is exact in fp32 and in bf16 storage, so
    rankings are independent of summation order (SURVEY.md §7 "Bit-exact top-k")."""
    g = torch.Generator().manual_seed(seed)
    return (torch.randint(-4, 5, shape, generator=g).float() / 16.0).to(device)


# ---------------------------------------------------------------------------------------------------
# synthetic vocabulary + caption STRINGS (the bert-base-uncased vocabulary is not available offline)
# ---------------------------------------------------------------------------------------------------
def _word_of(i: int) -> str:
    """Distinct all-letter pseudo-word for vocabulary id i (consonant/vowel alternation, so BasicTokenizer never
    splits it and lower-casing / accent stripping leave it alone)."""
    cons, vow = "bcdfghjklmnprstvwz", "aeiou"
    w, k = [], i
    for pos in range(6):
        if pos % 2 == 0:
            w.append(cons[k % len(cons)])
            k //= len(cons)
        else:
            w.append(vow[k % len(vow)])
            k //= len(vow)
    return "".join(w)


def make_vocab():
    """30 522 tokens laid out like bert-base-uncased: [PAD]=0, [unused*], [UNK]=100, [CLS]=101, [SEP]=102, [MASK]=103,
    more [unused*] up to id 998, punctuation at 999.., then one whole-word token per id (`_word_of`) and a tail of
    `##` continuation pieces.  Every id in [1000, 30000) is a whole word, so `make_captions` round-trips to
    `make_token_ids` through any correct WordPiece tokenizer."""
    toks = ["[PAD]"] + [f"[unused{i}]" for i in range(99)] + ["[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    toks += [f"[unused{i}]" for i in range(99, 99 + 1000 - len(toks))]
    assert len(toks) == 1000
    toks += [_word_of(i) for i in range(1000, 30000)]
    punct = list("!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~")
    toks += punct
    toks += ["##" + _word_of(i)[:3] for i in range(30000 + len(punct), 30522)]
    assert len(toks) == 30522 and len(set(toks)) == 30522
    return toks


def write_vocab(path: str) -> str:
    with open(path, "w", encoding="utf-8") as f:
        f.write("\n".join(make_vocab()) + "\n")
    return path


def make_captions(n: int, seed: int = 4321, vocab=None):
    """Caption strings whose WordPiece ids over `make_vocab()` are exactly `make_token_ids(n, seed)`."""
    vocab = vocab or make_vocab()
    ids, mask = make_token_ids(n, seed)
    return [" ".join(vocab[int(t)] for t in row[1:int(m.sum()) - 1]) for row, m in zip(ids, mask)]


def make_structured_images(n: int, seed: int = 2468):
    """Images with low-frequency content (per image a random 1x1 / 2x2 / 4x4 colour-block field, nearest-upsampled,
    plus 30 % pixel noise), clamped to the post-Normalize range.  White-noise images (`make_images`) all look alike
    to a randomly initialised ViT (similarities of a query to the whole gallery spread by 6e-3); these spread them by
    3e-2 at the same numerical noise, which is what a Recall@K comparison needs (tests/golden/recall_full_L.pt)."""
    g = torch.Generator().manual_seed(seed)
    out = torch.empty(n, 3, 224, 224)
    for i in range(n):
        cells = (1, 2, 4)[int(torch.randint(0, 3, (1,), generator=g))]
        low = torch.randn(1, 3, cells, cells, generator=g)
        up = torch.nn.functional.interpolate(low, size=(224, 224), mode="nearest")
        out[i] = (up[0] * 1.54 + 0.3 * torch.randn(3, 224, 224, generator=g)).clamp_(-2.2, 2.2)
    return out
