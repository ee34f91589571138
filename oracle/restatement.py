"""TEST INFRASTRUCTURE — CPU/torch fp32 restatement of the reference's composed-image-retrieval
inference arithmetic.  NOT part of the product: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / reference legs may import it; sprc_b200/ never does.

Pinning: the reference ships no tests or golden vectors, so this restatement is pinned
against outputs of the reference ITSELF, produced in the build container by oracle/make_golden.py
(unmodified reference classes imported from /root/reference through oracle/ref_loader.py) and
committed under tests/golden/.  tests/test_oracle.py checks restatement == golden (fp32, 1e-5) and,
where /root/reference is present, restatement == live reference.

Every function cites the reference lines it restates (paths relative to /root/reference/src/lavis/models).
All arithmetic is floating point (fp32 here; the reference's GPU path autocasts the ViT to fp16).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _ln(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def _mha(q, k, v, heads, scale, mask=None):
    """softmax(q k^T * scale + mask) v with head-major column blocks (eva_vit.py:126-145, Qformer.py:167-274)."""
    B, Lq, D = q.shape
    Lk = k.shape[1]
    dh = D // heads
    qh = q.view(B, Lq, heads, dh).transpose(1, 2)
    kh = k.view(B, Lk, heads, dh).transpose(1, 2)
    vh = v.view(B, Lk, heads, dh).transpose(1, 2)
    s = torch.matmul(qh, kh.transpose(-1, -2)) * scale
    if mask is not None:
        s = s + mask[:, None, None, :]
    p = s.softmax(dim=-1)
    return torch.matmul(p, vh).transpose(1, 2).reshape(B, Lq, D)


def patchify(images):
    """Stride-14 14x14 conv as a matmul operand: [B,3,224,224] -> [B,256,588] in (c,ky,kx) order
    (eva_vit.py:196,203; clip_vit.py:160,173-175)."""
    B = images.shape[0]
    x = images.view(B, 3, 16, 14, 16, 14).permute(0, 2, 4, 1, 3, 5)
    return x.reshape(B, 256, 588)


def vit_eva_g(sd, images, depth):
    """eva_vit.VisionTransformer.forward_features (eva_vit.py:324-340), Block :173-176, Attention
    :118-148 (bias = cat(q_bias, 0, v_bias), q scaled by dh^-0.5), Mlp :54-61 (erf GELU); no final norm."""
    p = "visual_encoder."
    Dv = sd[p + "cls_token"].shape[-1]
    x = patchify(images) @ sd[p + "patch_embed.proj.weight"].float().view(Dv, 588).t() + sd[
        p + "patch_embed.proj.bias"].float()
    x = torch.cat([sd[p + "cls_token"].float().expand(x.shape[0], -1, -1), x], dim=1) + sd[p + "pos_embed"].float()
    heads = 16
    scale = (Dv // heads) ** -0.5
    for i in range(depth):
        b = f"{p}blocks.{i}."
        h = _ln(x, sd[b + "norm1.weight"], sd[b + "norm1.bias"], 1e-6)
        bias = torch.cat([sd[b + "attn.q_bias"], torch.zeros_like(sd[b + "attn.v_bias"]), sd[b + "attn.v_bias"]])
        qkv = h @ sd[b + "attn.qkv.weight"].float().t() + bias.float()
        q, k, v = qkv.split(Dv, dim=-1)
        a = _mha(q, k, v, heads, scale)
        x = x + a @ sd[b + "attn.proj.weight"].float().t() + sd[b + "attn.proj.bias"].float()
        h = _ln(x, sd[b + "norm2.weight"], sd[b + "norm2.bias"], 1e-6)
        h = F.gelu(h @ sd[b + "mlp.fc1.weight"].float().t() + sd[b + "mlp.fc1.bias"].float())
        x = x + h @ sd[b + "mlp.fc2.weight"].float().t() + sd[b + "mlp.fc2.bias"].float()
    return x


def vit_clip_l(sd, images, depth):
    """clip_vit.VisionTransformer.forward (clip_vit.py:171-185): ln_pre, ResidualAttentionBlock :132-139
    (nn.MultiheadAttention, QuickGELU :109-111, LN eps 1e-5 in fp32); ln_final is not applied."""
    p = "visual_encoder."
    Dv = sd[p + "class_embedding"].shape[0]
    x = patchify(images) @ sd[p + "conv1.weight"].float().view(Dv, 588).t()
    cls = sd[p + "class_embedding"].float().view(1, 1, Dv).expand(x.shape[0], -1, -1)
    x = torch.cat([cls, x], dim=1) + sd[p + "positional_embedding"].float()
    x = _ln(x, sd[p + "ln_pre.weight"], sd[p + "ln_pre.bias"], 1e-5)
    heads = 16
    scale = (Dv // heads) ** -0.5
    for i in range(depth):
        b = f"{p}transformer.resblocks.{i}."
        h = _ln(x, sd[b + "ln_1.weight"], sd[b + "ln_1.bias"], 1e-5)
        qkv = h @ sd[b + "attn.in_proj_weight"].float().t() + sd[b + "attn.in_proj_bias"].float()
        q, k, v = qkv.split(Dv, dim=-1)
        a = _mha(q, k, v, heads, scale)
        x = x + a @ sd[b + "attn.out_proj.weight"].float().t() + sd[b + "attn.out_proj.bias"].float()
        h = _ln(x, sd[b + "ln_2.weight"], sd[b + "ln_2.bias"], 1e-5)
        h = h @ sd[b + "mlp.c_fc.weight"].float().t() + sd[b + "mlp.c_fc.bias"].float()
        h = h * torch.sigmoid(1.702 * h)
        x = x + h @ sd[b + "mlp.c_proj.weight"].float().t() + sd[b + "mlp.c_proj.bias"].float()
    return x


def _infer_dims(sd):
    vit = "eva_clip_g" if "visual_encoder.cls_token" in sd else "clip_L"
    pre = "visual_encoder.blocks." if vit == "eva_clip_g" else "visual_encoder.transformer.resblocks."
    depth = 1 + max(int(k[len(pre):].split(".")[0]) for k in sd if k.startswith(pre))
    ql = 1 + max(int(k.split(".")[4]) for k in sd if k.startswith("Qformer.bert.encoder.layer."))
    return vit, depth, ql


def image_embeds(sd, images):
    """ln_vision(visual_encoder(image)).float()  (blip2_qformer_cir_align_prompt.py:366-368, blip2.py:193-199)."""
    vit, depth, _ = _infer_dims(sd)
    x = vit_eva_g(sd, images, depth) if vit == "eva_clip_g" else vit_clip_l(sd, images, depth)
    return _ln(x, sd["ln_vision.weight"], sd["ln_vision.bias"], 1e-5)


def _lin(sd, name, x):
    return x @ sd[name + ".weight"].float().t() + sd[name + ".bias"].float()


def qformer(sd, query_embeds, input_ids=None, attention_mask=None, enc=None):
    """Qformer.BertModel.forward (Qformer.py:810-973) in the three modes the path uses.
    embeddings :98-113 (cat(query_embeds, word+pos) then ONE LayerNorm over all rows, eps 1e-12);
    mask :799-808 ((1-m) * -10000, query rows always visible); BertLayer.forward :408-480: self-attention
    on all rows; if `enc` is given, rows[:32] take cross-attention (even layers) and the *_query FFN while
    rows[32:] take the text FFN; if `enc` is None every row takes the text FFN (:434-435,469-475)."""
    _, _, n_layers = _infer_dims(sd)
    e = "Qformer.bert.embeddings."
    B = query_embeds.shape[0]
    x = query_embeds
    mask = None
    if input_ids is not None:
        t = sd[e + "word_embeddings.weight"][input_ids] + sd[e + "position_embeddings.weight"][: input_ids.shape[1]]
        x = torch.cat([query_embeds, t], dim=1)
        full = torch.cat([torch.ones(B, 32, dtype=attention_mask.dtype), attention_mask], dim=1)
        mask = (1.0 - full.float()) * -10000.0
    x = _ln(x, sd[e + "LayerNorm.weight"], sd[e + "LayerNorm.bias"], 1e-12)
    for l in range(n_layers):
        p = f"Qformer.bert.encoder.layer.{l}."
        a = _mha(_lin(sd, p + "attention.self.query", x), _lin(sd, p + "attention.self.key", x),
                 _lin(sd, p + "attention.self.value", x), 12, 0.125, mask)
        x = _ln(_lin(sd, p + "attention.output.dense", a) + x, sd[p + "attention.output.LayerNorm.weight"],
                sd[p + "attention.output.LayerNorm.bias"], 1e-12)

        def ffn(h, nm):
            i = F.gelu(_lin(sd, p + f"intermediate{nm}.dense", h))
            return _ln(_lin(sd, p + f"output{nm}.dense", i) + h, sd[p + f"output{nm}.LayerNorm.weight"],
                       sd[p + f"output{nm}.LayerNorm.bias"], 1e-12)

        if enc is not None:
            q = x[:, :32]
            if l % 2 == 0:
                c = _mha(_lin(sd, p + "crossattention.self.query", q), _lin(sd, p + "crossattention.self.key", enc),
                         _lin(sd, p + "crossattention.self.value", enc), 12, 0.125)
                q = _ln(_lin(sd, p + "crossattention.output.dense", c) + q,
                        sd[p + "crossattention.output.LayerNorm.weight"],
                        sd[p + "crossattention.output.LayerNorm.bias"], 1e-12)
            out = ffn(q, "_query")
            if x.shape[1] > 32:
                out = torch.cat([out, ffn(x[:, 32:], "")], dim=1)
            x = out
        else:
            x = ffn(x, "")
    return x


def extract_target_features(sd, images):
    """blip2_qformer_cir_align_prompt.py:364-386 -> (image_features [B,32,256], image_embeds_frozen [B,257,Dv])."""
    raws = image_embeds(sd, images)
    q = sd["query_tokens"].float().expand(raws.shape[0], -1, -1)
    h = qformer(sd, q, enc=raws)
    feats = F.normalize(_lin(sd, "vision_proj", h), dim=-1)
    return feats, raws


def fusion_features(sd, reference_embeds, input_ids, attention_mask):
    """The fusion half of `inference` (blip2_qformer_cir_align_prompt.py:312-350) -> [Bq,256]."""
    B = reference_embeds.shape[0]
    q = sd["query_tokens"].float().expand(B, -1, -1)
    fusion = qformer(sd, q, input_ids, attention_mask, enc=reference_embeds)
    text = qformer(sd, fusion[:, :32], input_ids, attention_mask, enc=None)
    return F.normalize(_lin(sd, "text_proj", text[:, 32]), dim=-1)


def similarity(fusion_feats, target_feats):
    """sim[b,n] = max_t <f_b, g_{n,t}>  (blip2_qformer_cir_align_prompt.py:353-358) as one matmul."""
    N = target_feats.shape[0]
    s = fusion_feats @ target_feats.reshape(N * 32, -1).t()
    return s.view(fusion_feats.shape[0], N, 32).max(dim=-1).values


def inference(sd, reference_embeds, target_feats, input_ids, attention_mask):
    return similarity(fusion_features(sd, reference_embeds, input_ids, attention_mask), target_feats)


def ranking(sim, k=None):
    """argsort(1 - sim) (validate_blip.py:44-46,253-255) made deterministic: ties -> lower gallery row."""
    order = torch.argsort(-sim, dim=-1, stable=True)
    return order if k is None else order[:, :k]


def inference_rerank(sd, ref_embeds, tgt_embeds, input_ids, attention_mask):
    """blip2_qformer_cir_rerank.py:399-445: each of R references is paired with T = len(tgt)/R targets;
    enc = cat(ref, tgt) (514 tokens); p = softmax(mean_q itm_head(h[:, :32]))[:, 1]."""
    R = ref_embeds.shape[0]
    T = tgt_embeds.shape[0] // R if R > 1 else tgt_embeds.shape[0]
    ref = ref_embeds.repeat_interleave(T, dim=0)
    ids = input_ids.repeat_interleave(T, dim=0)
    am = attention_mask.repeat_interleave(T, dim=0)
    q = sd["query_tokens"].float().expand(ref.shape[0], -1, -1)
    h = qformer(sd, q, ids, am, enc=torch.cat([ref, tgt_embeds], dim=1))
    logits = _lin(sd, "itm_head", h[:, :32]).mean(dim=1)
    return logits.softmax(dim=-1)[:, -1]


def cat_inference_rerank(sd, ref_embeds, tgt_feats, input_ids, attention_mask):
    """blip2_qformer_cir_cat.py:337-398: the composed query of each of R references against its own T candidate
    feature blocks `tgt_feats` [R*T,32,256]; sim = max over the 32 tokens (:392-394), no temperature."""
    R = ref_embeds.shape[0]
    T = tgt_feats.shape[0] // R if R > 1 else tgt_feats.shape[0]
    f = fusion_features(sd, ref_embeds, input_ids, attention_mask).repeat_interleave(T, dim=0)   # [R*T,256]
    return torch.einsum("ntd,nd->nt", tgt_feats, f).max(dim=-1).values


# ------------------------------------------------------------------------------------------------
# metrics tail of validate_blip.py on integer ids; used to check Recall@K parity
# ------------------------------------------------------------------------------------------------
def cirr_recalls(order, reference_idx, target_idx, group_members):
    """validate_blip.py:253-285 with gallery rows instead of name strings.
    order: [Q,N] ranking; reference_idx/target_idx: [Q]; group_members: [Q,6] (includes ref and target)."""
    Q, N = order.shape
    keep = order != reference_idx[:, None]
    ranked = order[keep].view(Q, N - 1)
    labels = ranked == target_idx[:, None]
    gm = (ranked[:, :, None] == group_members[:, None, :]).any(-1)
    glabels = labels[gm].view(Q, -1)
    assert bool((labels.sum(-1) == 1).all()) and bool((glabels.sum(-1) == 1).all())
    rec = lambda l, k: (l[:, :k].sum().item() / Q) * 100.0  # noqa: E731
    return (rec(glabels, 1), rec(glabels, 2), rec(glabels, 3), rec(labels, 1), rec(labels, 5), rec(labels, 10),
            rec(labels, 50))


def fiq_recalls(order, target_idx):
    """validate_blip.py:44-55."""
    labels = order == target_idx[:, None]
    Q = order.shape[0]
    return (labels[:, :10].sum().item() / Q) * 100.0, (labels[:, :50].sum().item() / Q) * 100.0


def plant_targets(order, reference_idx, seed=11):
    """Labels for a synthetic split, planted from an ORACLE ranking: after the reference image is
    removed from each query's ranking (validate_blip.py:259-262) the target is the item at rank r, r drawn with
    P(r<=1)=.2, P(r<=5)=.5, P(r<=10)=.7, P(r<=50)=.95, rest <=100 — so the oracle's recalls are ~20/50/70/95 %.
    Returns (target_idx [Q], ranks [Q] 1-based, group_members [Q,6] holding the reference, the target and 4 others)."""
    g = torch.Generator().manual_seed(seed)
    Q, N = order.shape
    hi = min(100, N - 1)
    bands = [(1, 1, .2), (2, 5, .3), (6, 10, .2), (11, min(50, hi), .25), (min(51, hi), hi, .05)]
    u = torch.rand(Q, generator=g)
    ranks = torch.empty(Q, dtype=torch.long)
    for j in range(Q):
        acc = 0.0
        for lo, up, p in bands:
            acc += p
            if u[j] < acc or (lo, up, p) == bands[-1]:
                ranks[j] = int(torch.randint(lo, up + 1, (1,), generator=g))
                break
    keep = order != reference_idx[:, None]
    ranked = order[keep].view(Q, N - 1)
    target = ranked[torch.arange(Q), ranks - 1]
    members = torch.empty(Q, 6, dtype=torch.long)
    for j in range(Q):
        others = [int(x) for x in torch.randperm(N, generator=g) if int(x) not in (int(reference_idx[j]), int(target[j]))][:4]
        members[j] = torch.tensor([int(reference_idx[j]), int(target[j])] + others)
    return target, ranks, members


def plant_targets_with_margin(sim, reference_idx, margin, seed=11, cutoffs=(1, 5, 10, 50)):
    """`plant_targets` on an oracle SIMILARITY matrix, with every label planted where the oracle's own ranking is
    robust: after the reference is removed, the target's score differs by >= `margin` from the score of the item on
    the other side of every cut-off K (the item at rank K+1 if the target is inside the top K, the item at rank K if
    it is outside), and from the scores of the 4 other group members.  With margin = 2 x the embedding tolerance two
    implementations whose similarities agree to that tolerance MUST produce identical hit/miss decisions, so
    "Recall@K within +-0.05" is testable as written; ranks keep the 20/50/70/95 % bands of `plant_targets`.
    Returns (target_idx [Q], ranks [Q], group_members [Q,6])."""
    g = torch.Generator().manual_seed(seed)
    order = ranking(sim)
    Q, N = order.shape
    hi = min(100, N - 1)
    bands = [(1, 1, .2), (2, 5, .3), (6, 10, .2), (11, min(50, hi), .25), (min(51, hi), hi, .05)]
    keep = order != reference_idx[:, None]
    ranked = order[keep].view(Q, N - 1)
    s_ranked = torch.gather(sim, 1, ranked)

    def robust(j, r):   # r 1-based
        st = float(s_ranked[j, r - 1])
        for K in cutoffs:
            if K >= N - 1:
                continue
            if r <= K and st - float(s_ranked[j, K]) < margin:
                return False
            if r > K and float(s_ranked[j, K - 1]) - st < margin:
                return False
        return True

    target = torch.empty(Q, dtype=torch.long)
    ranks = torch.empty(Q, dtype=torch.long)
    members = torch.empty(Q, 6, dtype=torch.long)
    for j in range(Q):
        chosen = None
        for _ in range(64):
            u, acc = float(torch.rand(1, generator=g)), 0.0
            for lo, up, p in bands:
                acc += p
                if u < acc or (lo, up, p) == bands[-1]:
                    cand = [r for r in torch.randperm(up - lo + 1, generator=g).add(lo).tolist() if robust(j, r)]
                    if cand:
                        chosen = cand[0]
                    break
            if chosen is not None:
                break
        if chosen is None:
            raise RuntimeError(f"query {j}: no rank with margin {margin} at every cut-off")
        ranks[j] = chosen
        target[j] = ranked[j, chosen - 1]
        st = float(s_ranked[j, chosen - 1])
        others = [int(x) for x in torch.randperm(N, generator=g)
                  if int(x) not in (int(reference_idx[j]), int(target[j])) and abs(float(sim[j, int(x)]) - st) >= margin][:4]
        if len(others) < 4:
            raise RuntimeError(f"query {j}: fewer than 4 group members with margin {margin}")
        members[j] = torch.tensor([int(reference_idx[j]), int(target[j])] + others)
    return target, ranks, members


def target_ranks(sim, reference_idx, target_idx):
    """1-based rank of each query's target in argsort(1 - sim) after the reference row is removed."""
    order = ranking(sim)
    Q, N = order.shape
    ranked = order[order != reference_idx[:, None]].view(Q, N - 1)
    return (ranked == target_idx[:, None]).float().argmax(dim=1) + 1
