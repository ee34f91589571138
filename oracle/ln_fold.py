"""TEST INFRASTRUCTURE — CPU emulation of the Q-Former DEVICE schedules of sprc_b200/csrc, with their 16-bit rounding
points: the default schedule (GEMM + LayerNorm kernel) and the LayerNorm-FOLDED schedule of csrc/ln_fold.cu
(SPRC_LN_FOLD=1).  Only tests/ may import it.

What it pins down (tests/test_ln_fold.py):
* the algebra of the fold — with rounding switched off the folded schedule reproduces oracle/restatement.qformer
  (i.e. /root/reference/src/lavis/models/blip2_models/Qformer.py:408-480 BertLayer.forward with the post-LN sublayers
  :291-295, :373-381) to fp32 round-off: which LayerNorm's (gamma, beta) is folded into which weight, which statistics
  buffer each row range reads, the 12-partial Chan merge, the materialising LayerNorms in front of the last layer;
* its numerical cost — with bf16 rounding on, the folded schedule is as close to the fp32 oracle as the default one.

Schedule restated (csrc/ln_fold.cu Model::qformer_layers_ragged_fold, csrc/gemm2.cu FOLD = 1 epilogues):
  producer   s' = a16 W16^T + b + LN(s)            -> s' fp32, s16 = round16(s'), 12 x (mean, M2) per row of 768
  consumer   round16(act(rstd (s16 Wf16^T - mean c) + d)),  Wf16 = round16(W16 diag(gamma)), c = rowsum(Wf16),
             d = W16 beta + b
Rows are kept per sample ([B, 64, 768]: 32 query rows + 32 text rows with the padding mask); the device's ragged layout
only drops the dead text rows, every live row sees the same arithmetic.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import restatement as R

PARTS = 12   # csrc/common.h kFoldParts


def _r16(x, on):
    return x.bfloat16().float() if on else x


def row_stats_partials(s):
    """12 (mean, M2) partials per row of 768, each over the 64 columns one epilogue thread owns, built chunk by chunk
    (16 columns, Chan merge) exactly as csrc/gemm2.cu's producer epilogue does."""
    rows = s.reshape(-1, PARTS, 4, 16)
    cm = rows.mean(-1)                                    # [R, 12, 4]
    cm2 = ((rows - cm[..., None]) ** 2).sum(-1)
    mean, m2 = cm[..., 0], cm2[..., 0]
    for cc in range(1, 4):
        na, nt = 16.0 * cc, 16.0 * cc + 16.0
        dl = cm[..., cc] - mean
        mean = mean + dl * (16.0 / nt)
        m2 = m2 + cm2[..., cc] + dl * dl * (na * 16.0 / nt)
    return torch.stack([mean, m2], dim=-1).reshape(*s.shape[:-1], PARTS, 2)


def merge_stats(st, eps):
    """fold_row_stats (csrc/gemm2.cu): equal-count Chan merge of the 12 partials -> (mean, rstd)."""
    m = st[..., 0].mean(-1)
    m2 = (st[..., 1] + 64.0 * (st[..., 0] - m[..., None]) ** 2).sum(-1)
    return m, torch.rsqrt(m2 / 768.0 + eps)


def fold_weight(w, bias, gamma, beta, rnd):
    """fold_weight_kernel (csrc/ln_fold.cu): Wf = round16(W16 diag(gamma)); c = rowsum(Wf); d = W16 beta + bias."""
    w16 = _r16(w.float(), rnd)
    wf = _r16(w16 * gamma[None, :], rnd)
    return wf, wf.sum(-1), w16 @ beta + bias


class _Stream:
    """Residual stream of a row range: materialised (x = LN output) or raw (pre-LN sums + statistics + owed LN)."""

    def __init__(self, x):
        self.x, self.raw, self.st, self.g, self.b, self.ln = x, False, None, None, None, None


def static_fold_table(l, with_enc):
    """Which LayerNorm csrc/ln_fold.cu Model::prepare_fold folds into which weight of layer l (the device derives these
    tables once per weight load; the emulation below folds whatever LayerNorm the stream owes and asserts both agree).
    Keys: (weight, row range)."""
    p, pp = f"Qformer.bert.encoder.layer.{l}.", f"Qformer.bert.encoder.layer.{l - 1}."
    cross = l % 2 == 0
    t = {
        ("qkv", "query"): pp + ("output_query.LayerNorm" if with_enc else "output.LayerNorm"),   # QfFold::qkv_q / qkv_t
        ("qkv", "text"): pp + "output.LayerNorm",                                               # QfFold::qkv_t
        ("crossattention.self.query", "query"): p + "attention.output.LayerNorm",               # QfFold::cq
        ("intermediate_query.dense", "query"): p + ("crossattention.output.LayerNorm" if cross
                                                    else "attention.output.LayerNorm"),         # QfFold::qi
        ("intermediate.dense", "query"): p + "attention.output.LayerNorm",                      # QfFold::ti (text pass)
        ("intermediate.dense", "text"): p + "attention.output.LayerNorm",                       # QfFold::ti
    }
    return t


def _lin_params(sd, name):
    return sd[name + ".weight"].float(), sd[name + ".bias"].float()


def qformer_device(sd, query_embeds, input_ids, attention_mask, enc, fold, rnd=True, eps=1e-12):
    """One Q-Former pass (fusion: enc given; text: enc None) in the device schedule.  fold=False: GEMM + LayerNorm
    kernel per sublayer (csrc/model.cu qformer_layers_ragged); fold=True: layers 0 .. L-2 folded, LayerNorms
    materialised in front of the last layer (csrc/ln_fold.cu).  Returns [B, 64, 768] like restatement.qformer."""
    _, _, n_layers = R._infer_dims(sd)
    e = "Qformer.bert.embeddings."
    B = query_embeds.shape[0]
    t = sd[e + "word_embeddings.weight"][input_ids] + sd[e + "position_embeddings.weight"][: input_ids.shape[1]]
    x = torch.cat([query_embeds, t], dim=1)
    full = torch.cat([torch.ones(B, 32, dtype=attention_mask.dtype), attention_mask], dim=1)
    mask = (1.0 - full.float()) * -10000.0
    x = R._ln(x, sd[e + "LayerNorm.weight"], sd[e + "LayerNorm.bias"], eps)
    enc16 = _r16(enc, rnd) if enc is not None else None
    # per-range streams: query rows [:, :32] and text rows [:, 32:]
    sq, stx = _Stream(x[:, :32].contiguous()), _Stream(x[:, 32:].contiguous())

    def ln_of(s):       # what LN the raw stream owes, applied with its own 12-partial statistics
        if not s.raw:
            return s.x
        m, rs = merge_stats(s.st, eps)
        return (s.x - m[..., None]) * rs[..., None] * s.g + s.b

    def operand(s):     # 16-bit A operand copy of the stream (normalised when materialised, raw otherwise)
        return _r16(s.x, rnd)

    def consume(s, name, act=None, key=None):
        """GEMM reading LN(stream) -> 16-bit output."""
        w, b = _lin_params(sd, name)
        if not s.raw:
            y = operand(s) @ _r16(w, rnd).t() + b
        else:
            assert s.ln == table[key], (name, key, s.ln, table[key])   # the device's static table folds the same LN
            wf, c, d = fold_weight(w, b, s.g, s.b, rnd)
            m, rs = merge_stats(s.st, eps)
            y = rs[..., None] * (operand(s) @ wf.t() - m[..., None] * c) + d
        if act is not None:
            y = act(y)
        return _r16(y, rnd)

    def produce(s, a16, name, ln_name, folded):
        """Post-LN sublayer: stream <- LN(a16 W^T + b + LN(stream)); folded: keep the raw sums + statistics."""
        w, b = _lin_params(sd, name)
        snew = a16 @ _r16(w, rnd).t() + b + ln_of(s)
        g, be = sd[ln_name + ".weight"].float(), sd[ln_name + ".bias"].float()
        if folded:
            s.x, s.raw, s.st, s.g, s.b, s.ln = snew, True, row_stats_partials(snew), g, be, ln_name
        else:
            s.x, s.raw = R._ln(snew, g, be, eps), False

    def materialise(s):
        if s.raw:
            # csrc/ln_fold.cu: the plain LayerNorm kernel (its own two-pass statistics) over the raw sums
            s.x, s.raw = R._ln(s.x, s.g, s.b, eps), False

    for l in range(n_layers):
        p = f"Qformer.bert.encoder.layer.{l}."
        folded = fold and l < n_layers - 1
        table = static_fold_table(l, enc is not None)
        if fold and l == n_layers - 1:
            materialise(sq), materialise(stx)
        # self-attention over all 64 rows
        kq, kt = ("qkv", "query"), ("qkv", "text")
        q = torch.cat([consume(sq, p + "attention.self.query", key=kq),
                       consume(stx, p + "attention.self.query", key=kt)], 1)
        k = torch.cat([consume(sq, p + "attention.self.key", key=kq), consume(stx, p + "attention.self.key", key=kt)], 1)
        v = torch.cat([consume(sq, p + "attention.self.value", key=kq),
                       consume(stx, p + "attention.self.value", key=kt)], 1)
        ctx = _r16(R._mha(q, k, v, 12, 0.125, mask), rnd)
        produce(sq, ctx[:, :32], p + "attention.output.dense", p + "attention.output.LayerNorm", folded)
        produce(stx, ctx[:, 32:], p + "attention.output.dense", p + "attention.output.LayerNorm", folded)
        if enc is not None:
            if l % 2 == 0:
                cq = consume(sq, p + "crossattention.self.query", key=("crossattention.self.query", "query"))
                wk, bk = _lin_params(sd, p + "crossattention.self.key")
                wv, bv = _lin_params(sd, p + "crossattention.self.value")
                ck = _r16(enc16 @ _r16(wk, rnd).t() + bk, rnd)
                cv = _r16(enc16 @ _r16(wv, rnd).t() + bv, rnd)
                cctx = _r16(R._mha(cq, ck, cv, 12, 0.125), rnd)
                produce(sq, cctx, p + "crossattention.output.dense", p + "crossattention.output.LayerNorm", folded)
            hq = consume(sq, p + "intermediate_query.dense", F.gelu, key=("intermediate_query.dense", "query"))
            produce(sq, hq, p + "output_query.dense", p + "output_query.LayerNorm", folded)
            ht = consume(stx, p + "intermediate.dense", F.gelu, key=("intermediate.dense", "text"))
            produce(stx, ht, p + "output.dense", p + "output.LayerNorm", folded)
        else:
            for s, rng in ((sq, "query"), (stx, "text")):
                h = consume(s, p + "intermediate.dense", F.gelu, key=("intermediate.dense", rng))
                produce(s, h, p + "output.dense", p + "output.LayerNorm", folded)
    materialise(sq), materialise(stx)
    return torch.cat([sq.x, stx.x], dim=1)


def fusion_features_device(sd, reference_embeds, input_ids, attention_mask, fold, rnd=True):
    """restatement.fusion_features in the device schedule: fusion pass, text pass, text_proj + normalise."""
    B = reference_embeds.shape[0]
    q = sd["query_tokens"].float().expand(B, -1, -1)
    fusion = qformer_device(sd, q, input_ids, attention_mask, reference_embeds, fold, rnd)
    text = qformer_device(sd, fusion[:, :32], input_ids, attention_mask, None, fold, rnd)
    w, b = _lin_params(sd, "text_proj")
    return F.normalize(_r16(text[:, 32], rnd) @ _r16(w, rnd).t() + b, dim=-1)


# ------------------------------------------------------------------------------------------------
# ViT blocks (pre-LN) in the device schedules: csrc/model.cu Model::vit_forward / csrc/ln_fold.cu vit_blocks_fold
# ------------------------------------------------------------------------------------------------
def _stats_any(s):
    """Row statistics partials for rows of any width that is a multiple of 64 (16 for ViT-L, 22 for ViT-g)."""
    parts = s.shape[-1] // 64
    rows = s.reshape(-1, parts, 4, 16)
    cm = rows.mean(-1)
    cm2 = ((rows - cm[..., None]) ** 2).sum(-1)
    mean, m2 = cm[..., 0], cm2[..., 0]
    for cc in range(1, 4):
        na, nt = 16.0 * cc, 16.0 * cc + 16.0
        dl = cm[..., cc] - mean
        mean = mean + dl * (16.0 / nt)
        m2 = m2 + cm2[..., cc] + dl * dl * (na * 16.0 / nt)
    return torch.stack([mean, m2], dim=-1).reshape(*s.shape[:-1], parts, 2)


def _merge_any(st, eps):
    parts = st.shape[-2]
    m = st[..., 0].mean(-1)
    m2 = (st[..., 1] + 64.0 * (st[..., 0] - m[..., None]) ** 2).sum(-1)
    return m, torch.rsqrt(m2 / (64.0 * parts) + eps)


def image_embeds_device(sd, images, fold, rnd=True):
    """restatement.image_embeds (eva_vit.py:324-340 / clip_vit.py:171-185 + ln_vision) in the device schedule.
    fold=True: norm1 of blocks >= 1 and every norm2 are folded into qkv / fc1 (raw residual stream, statistics written
    by the proj / fc2 producers); block 0's norm1 and ln_vision stay LayerNorm kernels."""
    vit, depth, _ = R._infer_dims(sd)
    p = "visual_encoder."
    eva = vit == "eva_clip_g"
    if eva:
        Dv = sd[p + "cls_token"].shape[-1]
        x = _r16(R.patchify(images), rnd) @ _r16(sd[p + "patch_embed.proj.weight"].float().view(Dv, 588), rnd).t() \
            + sd[p + "patch_embed.proj.bias"].float()
        x = torch.cat([sd[p + "cls_token"].float().expand(x.shape[0], -1, -1), x], 1) + sd[p + "pos_embed"].float()
        eps, act = 1e-6, F.gelu
    else:
        Dv = sd[p + "class_embedding"].shape[0]
        x = _r16(R.patchify(images), rnd) @ _r16(sd[p + "conv1.weight"].float().view(Dv, 588), rnd).t()
        cls = sd[p + "class_embedding"].float().view(1, 1, Dv).expand(x.shape[0], -1, -1)
        x = torch.cat([cls, x], 1) + sd[p + "positional_embedding"].float()
        x = R._ln(x, sd[p + "ln_pre.weight"], sd[p + "ln_pre.bias"], 1e-5)
        eps, act = 1e-5, (lambda h: h * torch.sigmoid(1.702 * h))
    scale = (Dv // 16) ** -0.5
    st = None   # statistics of the raw stream (fold only)

    def names(i):
        if eva:
            b = f"{p}blocks.{i}."
            bias = torch.cat([sd[b + "attn.q_bias"], torch.zeros_like(sd[b + "attn.v_bias"]), sd[b + "attn.v_bias"]])
            return (b + "norm1", b + "norm2", sd[b + "attn.qkv.weight"].float(), bias.float(), b + "attn.proj",
                    b + "mlp.fc1", b + "mlp.fc2")
        b = f"{p}transformer.resblocks.{i}."
        return (b + "ln_1", b + "ln_2", sd[b + "attn.in_proj_weight"].float(), sd[b + "attn.in_proj_bias"].float(),
                b + "attn.out_proj", b + "mlp.c_fc", b + "mlp.c_proj")

    def read_ln(x, st, ln, w, b, a=None):
        g, be = sd[ln + ".weight"].float(), sd[ln + ".bias"].float()
        if st is None:
            y = _r16(R._ln(x, g, be, eps), rnd) @ _r16(w, rnd).t() + b
        else:
            wf, c, d = fold_weight(w, b, g, be, rnd)
            m, rs = _merge_any(st, eps)
            y = rs[..., None] * (_r16(x, rnd) @ wf.t() - m[..., None] * c) + d
        return _r16(a(y) if a is not None else y, rnd)

    for i in range(depth):
        n1, n2, wqkv, bqkv, proj, fc1, fc2 = names(i)
        qkv = read_ln(x, st if fold else None, n1, wqkv, bqkv)
        q, k, v = qkv.split(Dv, dim=-1)
        a = _r16(R._mha(q, k, v, 16, scale), rnd)
        wp, bp = _lin_params(sd, proj)
        x = x + a @ _r16(wp, rnd).t() + bp
        st = _stats_any(x) if fold else None
        w1, b1 = _lin_params(sd, fc1)
        h = read_ln(x, st, n2, w1, b1, act)
        w2, b2 = _lin_params(sd, fc2)
        x = x + h @ _r16(w2, rnd).t() + b2
        st = _stats_any(x) if fold else None
    return R._ln(x, sd["ln_vision.weight"], sd["ln_vision.bias"], 1e-5)
