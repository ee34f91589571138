"""CPU checks of the LayerNorm-folded Q-Former schedule (csrc/ln_fold.cu, opt-in SPRC_LN_FOLD=1) through its
emulation oracle/ln_fold.py: the fold's algebra against the fp32 restatement of the reference
(/root/reference/src/lavis/models/blip2_models/Qformer.py:291-295, 373-381, 408-480) and its bf16 cost against the
default device schedule.  The CUDA kernels themselves are compared in tests/test_ln_fold_gpu.py."""
import pytest
import torch

from oracle import ln_fold as LF
from oracle import restatement as R
from oracle import synth


def _case(qf_layers, seed=3, B=3):
    sd = synth.make_state_dict("clip_L", vit_depth=1, qf_layers=qf_layers, seed=0, gain=2.5)
    g = torch.Generator().manual_seed(seed)
    # non-trivial LayerNorm parameters: with the initialisers' gamma = 1, beta = 0 a mis-routed fold would go unnoticed
    for k in sd:
        if "LayerNorm.weight" in k:
            sd[k] = 1.0 + 0.3 * torch.randn(sd[k].shape, generator=g)
        elif "LayerNorm.bias" in k:
            sd[k] = 0.2 * torch.randn(sd[k].shape, generator=g)
    enc = torch.randn(B, 257, 1024, generator=g)
    ids, mask = synth.make_token_ids(B)
    return sd, enc, ids, mask


def test_row_statistics_partials_merge_to_the_row_moments():
    g = torch.Generator().manual_seed(0)
    s = torch.randn(37, 768, generator=g) * 3.0 + 1.5
    m, rs = LF.merge_stats(LF.row_stats_partials(s), 1e-12)
    assert torch.allclose(m, s.mean(-1), atol=1e-5)
    assert torch.allclose(rs, torch.rsqrt(s.var(-1, unbiased=False) + 1e-12), rtol=1e-5)


def test_folded_weight_identity():
    """LN(s) W^T + b == rstd (s Wf^T - mean c) + d  with the fold_weight_kernel quantities (no rounding)."""
    g = torch.Generator().manual_seed(1)
    s = torch.randn(16, 768, generator=g) * 2.0 + 0.7
    w, b = torch.randn(96, 768, generator=g) * 0.05, torch.randn(96, generator=g)
    gamma, beta = 1.0 + 0.3 * torch.randn(768, generator=g), 0.2 * torch.randn(768, generator=g)
    wf, c, d = LF.fold_weight(w, b, gamma, beta, rnd=False)
    m, rs = LF.merge_stats(LF.row_stats_partials(s), 1e-12)
    got = rs[:, None] * (s @ wf.t() - m[:, None] * c) + d
    want = R._ln(s, gamma, beta, 1e-12) @ w.t() + b
    assert (got - want).abs().max() < 2e-4


@pytest.mark.parametrize("qf_layers", [3, 4])
def test_folded_schedule_equals_reference_without_rounding(qf_layers):
    """Pure algebra: every (gamma, beta) routed to the right weight, statistics of the right row range, materialising
    LayerNorms in front of the last layer - fusion pass (cross-attention, dual FFN) and text pass."""
    sd, enc, ids, mask = _case(qf_layers)
    q = sd["query_tokens"].float().expand(enc.shape[0], -1, -1)
    want = R.qformer(sd, q, ids, mask, enc=enc)
    got = LF.qformer_device(sd, q, ids, mask, enc, fold=True, rnd=False)
    live = torch.cat([torch.ones_like(mask), mask], 1).bool()
    assert (got - want)[live].abs().max() < 5e-4
    want_t = R.qformer(sd, want[:, :32], ids, mask, enc=None)
    got_t = LF.qformer_device(sd, want[:, :32], ids, mask, None, fold=True, rnd=False)
    assert (got_t - want_t)[live].abs().max() < 5e-4
    f_want = R.fusion_features(sd, enc, ids, mask)
    f_got = LF.fusion_features_device(sd, enc, ids, mask, fold=True, rnd=False)
    assert (f_got - f_want).abs().max() < 1e-4


def test_folded_schedule_costs_no_accuracy_in_bf16():
    """With the device's bf16 rounding points the folded schedule is as close to the fp32 reference arithmetic as the
    default (GEMM + LayerNorm kernel) schedule: feeding raw pre-LN sums to the tensor cores instead of normalised rows
    does not amplify the operand rounding (|row mean| stays below the row's spread in a post-LN residual stream)."""
    sd, enc, ids, mask = _case(4)
    want = R.fusion_features(sd, enc, ids, mask)
    e_def = (LF.fusion_features_device(sd, enc, ids, mask, fold=False) - want).norm() / want.norm()
    e_fold = (LF.fusion_features_device(sd, enc, ids, mask, fold=True) - want).norm() / want.norm()
    print(f"rel-Frobenius vs fp32: default schedule {e_def:.3e}, folded {e_fold:.3e}")
    assert e_def < 2e-2
    assert e_fold < 1.5 * e_def + 1e-3


@pytest.mark.parametrize("vit", ["clip_L", "eva_clip_g"])
def test_vit_folded_schedule(vit):
    """ViT blocks with norm1 / norm2 folded into qkv / fc1 (csrc/ln_fold.cu vit_blocks_fold): algebra against the fp32
    restatement (eva_vit.py:173-176, clip_vit.py:132-139) with non-trivial LayerNorm parameters, 16 (ViT-L) and 22
    (ViT-g) statistics partials per token, and the bf16 cost against the default schedule."""
    sd = synth.make_state_dict(vit, vit_depth=3, qf_layers=1, seed=0, gain=2.5)
    g = torch.Generator().manual_seed(9)
    for k in sd:
        if k.startswith("visual_encoder.") and ("norm" in k or "ln_" in k):
            if k.endswith("weight"):
                sd[k] = 1.0 + 0.3 * torch.randn(sd[k].shape, generator=g)
            elif k.endswith("bias"):
                sd[k] = 0.2 * torch.randn(sd[k].shape, generator=g)
    images = synth.make_images(2)
    want = R.image_embeds(sd, images)
    got = LF.image_embeds_device(sd, images, fold=True, rnd=False)
    assert (got - want).abs().max() < 1e-3, (got - want).abs().max()
    e_def = (LF.image_embeds_device(sd, images, fold=False) - want).norm() / want.norm()
    e_fold = (LF.image_embeds_device(sd, images, fold=True) - want).norm() / want.norm()
    print(f"[{vit}] rel-Frobenius vs fp32: default schedule {e_def:.3e}, folded {e_fold:.3e}")
    assert e_fold < 1.5 * e_def + 1e-3
