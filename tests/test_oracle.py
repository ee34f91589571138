"""CPU tests pinning the oracle (oracle/restatement.py) to the REFERENCE's own outputs:
  * against the committed golden vectors (tests/golden/*.pt, produced by oracle/make_golden.py from the
    unmodified reference classes) — runs everywhere;
  * against the live reference imported from /root/reference — runs only in the build container.
The reference ships no tests / golden vectors of its own, so these are the pins.
"""
import os

import pytest
import torch

from oracle import ref_loader
from oracle import restatement as R
from oracle import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _case(name):
    g = torch.load(os.path.join(GOLDEN, f"{name}.pt"))
    c = g["case"]
    sd = synth.make_state_dict(c["vit"], c["vit_depth"], c["qf_layers"], seed=c["seed"])
    return g, c, sd


@pytest.mark.parametrize("name", ["tiny_L", "tiny_g", "full_L", "full_g"])
def test_restatement_matches_reference_golden(name):
    g, c, sd = _case(name)
    images = synth.make_images(c["n_images"])
    with torch.no_grad():
        feats, raws = R.extract_target_features(sd, images)
        fusion = R.fusion_features(sd, raws[g["ref_rows"]], g["input_ids"], g["attention_mask"])
        sim = R.similarity(fusion, feats)
    assert (raws[:, g["raw_rows"]] - g["raws_rows"]).abs().max().item() < 5e-5   # ViT-g, 39 blocks: 7e-6 measured
    assert (feats - g["feats"]).abs().max().item() < 2e-6
    assert (fusion - g["fusion"]).abs().max().item() < 2e-6
    assert (sim - g["sim"]).abs().max().item() < 2e-6
    # the ranking the reference derives from its own sim (validate_blip.py:253-254) is reproduced
    assert torch.equal(R.ranking(sim), torch.argsort(-g["sim"], dim=-1, stable=True))


@pytest.mark.parametrize("name", ["tiny_L", "tiny_g"])
def test_restatement_rerank_matches_reference_golden(name):
    g, c, sd = _case(name)
    rr = g["rerank_p"]
    with torch.no_grad():
        _, raws = R.extract_target_features(sd, synth.make_images(c["n_images"]))
        p = R.inference_rerank(sd, raws[rr["ref_rows"]], raws[rr["cand_rows"]], g["input_ids"][: rr["R"]],
                               g["attention_mask"][: rr["R"]])
    assert (p - rr["p"]).abs().max().item() < 2e-6


def test_golden_full_g_is_self_consistent():
    """Invariants of the stored reference outputs (unit-norm rows, sim == max-over-tokens of fusion . feats); the
    restatement itself is re-derived against this golden above (3 images through all 39 ViT-g blocks: ~10 s)."""
    g = torch.load(os.path.join(GOLDEN, "full_g.pt"))
    assert (g["feats"].norm(dim=-1) - 1).abs().max().item() < 1e-5
    assert (g["fusion"].norm(dim=-1) - 1).abs().max().item() < 1e-5
    assert (R.similarity(g["fusion"], g["feats"]) - g["sim"]).abs().max().item() < 2e-6


def test_similarity_equals_reference_broadcast_form():
    """The reference's broadcast matmul + max (align_prompt.py:353-358), literally, vs the one-GEMM form."""
    torch.manual_seed(0)
    f = torch.nn.functional.normalize(torch.randn(5, 256), dim=-1)
    t = torch.nn.functional.normalize(torch.randn(40, 32, 256), dim=-1)
    ref = torch.matmul(f.unsqueeze(1).unsqueeze(1), t.permute(0, 2, 1)).squeeze().max(-1).values
    assert (R.similarity(f, t) - ref).abs().max().item() < 1e-6


def test_recall_tail_matches_reference_string_version():
    """cirr_recalls / fiq_recalls on integer rows == the reference's name-matching code
    (validate_blip.py:253-285, 44-55), restated literally with numpy string arrays."""
    import numpy as np

    g = torch.Generator().manual_seed(3)
    Q, N = 12, 70
    sim = torch.rand(Q, N, generator=g)
    names = np.array([f"img{i:04d}" for i in range(N)])
    ref = torch.randint(0, N, (Q,), generator=g)
    order = R.ranking(sim)
    tgt = torch.stack([order[q][order[q] != ref[q]][[0, 2, 4, 9, 30, 55][q % 6]] for q in range(Q)])
    members = []
    for q in range(Q):
        others = [int(x) for x in order[q] if int(x) not in (int(ref[q]), int(tgt[q]))][:4]
        members.append([int(ref[q]), int(tgt[q])] + others)
    members = torch.tensor(members)
    got = R.cirr_recalls(order, ref, tgt, members)
    # literal restatement of validate_blip.py:253-285
    sorted_names = names[torch.argsort(1 - sim, dim=-1, stable=True).numpy()]
    ref_names, tgt_names = names[ref.numpy()], names[tgt.numpy()]
    mask = torch.tensor(sorted_names != np.repeat(ref_names, N).reshape(Q, -1))
    sorted_names = sorted_names[mask].reshape(Q, N - 1)
    labels = torch.tensor(sorted_names == np.repeat(tgt_names, N - 1).reshape(Q, -1))
    gm = names[members.numpy()]
    gmask = (sorted_names[..., None] == gm[:, None, :]).sum(-1).astype(bool)
    glabels = labels[gmask].reshape(Q, -1)
    want = tuple((torch.sum(l[:, :k]) / Q).item() * 100 for l, ks in ((glabels, (1, 2, 3)), (labels, (1, 5, 10, 50)))
                 for k in ks)
    assert got == pytest.approx(want, abs=1e-4)
    f10, f50 = R.fiq_recalls(order, tgt)
    lab = torch.tensor(names[order.numpy()] == np.repeat(tgt_names, N).reshape(Q, -1))
    assert (f10, f50) == pytest.approx(((torch.sum(lab[:, :10]) / Q).item() * 100,
                                        (torch.sum(lab[:, :50]) / Q).item() * 100), abs=1e-4)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_restatement_matches_live_reference_and_key_layout():
    """Build-container only: synthetic checkpoint keys/shapes == the reference's state_dict, and the
    restatement equals the live reference on a fresh seed (not the golden one)."""
    model = ref_loader.build_reference_model("clip_L", seed=0, vit_depth=1, qf_layers=2)
    sd = synth.make_state_dict("clip_L", 1, 2, seed=5)
    ref_keys = {k for k in model.state_dict() if not k.startswith("Qformer.cls.") and "position_ids" not in k}
    assert ref_keys == set(sd)
    for k, v in model.state_dict().items():
        if k in sd:
            assert tuple(v.shape) == tuple(sd[k].shape), k
    model.load_state_dict(sd, strict=False)
    images = synth.make_images(3, seed=9)
    ids, mask = synth.make_token_ids(2, seed=8)
    with torch.no_grad():
        f_ref, r_ref = model.extract_target_features(images)
        s_ref = ref_loader.call_inference(model, r_ref[:2], f_ref, ids, mask)
        f, r = R.extract_target_features(sd, images)
        s = R.inference(sd, r[:2], f, ids, mask)
    assert (f - f_ref).abs().max().item() < 2e-6
    assert (r - r_ref).abs().max().item() < 5e-5
    assert (s - s_ref).abs().max().item() < 2e-6


def test_cir_cat_golden_is_align_prompt_similarity_over_temp():
    """blip2_cir_cat differs from align_prompt in `inference` only by `/ self.temp` (blip2_qformer_cir_cat.py:331):
    the two reference runs on the same checkpoint must agree to fp32 noise."""
    import os

    import torch

    gd = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    a = torch.load(os.path.join(gd, "tiny_L.pt"))
    c = torch.load(os.path.join(gd, "tiny_L_cat.pt"))
    assert torch.allclose(c["feats"], a["feats"], atol=2e-6)
    assert torch.allclose(c["sim"] * c["temp"], a["sim"], atol=2e-6)


def test_cir_cat_rerank_restatement_matches_reference_golden():
    """blip2_cir_cat.inference_rerank (blip2_qformer_cir_cat.py:337-398) restated vs the reference's own output."""
    import os

    import torch

    from oracle import restatement as R
    from oracle import synth

    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_L_cat.pt"))
    case, rr = g["case"], g["rerank"]
    sd = synth.make_state_dict(case["vit"], case["vit_depth"], case["qf_layers"], seed=0)
    with torch.no_grad():
        feats, raws = R.extract_target_features(sd, synth.make_images(case["n_images"]))
        s = R.cat_inference_rerank(sd, raws[rr["ref_rows"]], feats[rr["cand_rows"]], g["input_ids"][:rr["R"]],
                                   g["attention_mask"][:rr["R"]])
    assert s.shape == rr["sim"].shape
    assert (s - rr["sim"]).abs().max().item() < 2e-6


def test_spread_golden_restatement_and_planted_recalls():
    """Recall-parity case (tests/golden/spread_L.pt, the reference's `inference` on a gain-2.5 checkpoint): the
    restatement reproduces the reference's similarity, and labels planted from that ranking give the oracle recalls
    the planting distribution promises (every recall > 0, exactly one positive per query)."""
    import os

    import torch

    from oracle import restatement as R
    from oracle import synth

    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "spread_L.pt"))
    c = g["case"]
    sd = synth.make_state_dict(c["vit"], c["vit_depth"], c["qf_layers"], seed=0, gain=c["gain"])
    with torch.no_grad():
        feats, raws = R.extract_target_features(sd, synth.make_images(c["n_images"]))
        sim = R.inference(sd, raws[g["ref_rows"]], feats, g["input_ids"], g["attention_mask"])
    # fp32 summation-order noise (batch 64 here vs 16 in the reference run) through the amplified weights
    assert (sim - g["sim"]).abs().max().item() < 5e-5
    assert g["sim"].max() - g["sim"].min() > 0.15         # spread over ~0.2, not the 0.016 of the gain-1 checkpoints
    order = R.ranking(g["sim"])
    tgt, ranks, members = R.plant_targets(order, g["ref_rows"])
    assert torch.equal(R.target_ranks(g["sim"], g["ref_rows"], tgt), ranks)
    rec = R.cirr_recalls(order, g["ref_rows"], tgt, members)
    Q = ranks.numel()
    assert rec[3:] == tuple(100.0 * int((ranks <= k).sum()) / Q for k in (1, 5, 10, 50))
    assert 5 < rec[3] < 40 and 30 < rec[4] < 70 and 50 < rec[5] < 90 and rec[6] > 85 and all(r > 0 for r in rec)


def test_restatement_rederives_full_g_golden():
    """the full-depth ViT-g golden re-derived by the restatement (39 EVA blocks, 3 images, 3 queries;
    about 20 s of CPU work), not only checked for self-consistency."""
    g, c, sd = _case("full_g")
    images = synth.make_images(c["n_images"])
    with torch.no_grad():
        feats, raws = R.extract_target_features(sd, images)
        fusion = R.fusion_features(sd, raws[g["ref_rows"]], g["input_ids"], g["attention_mask"])
        sim = R.similarity(fusion, feats)
    assert (raws[:, g["raw_rows"]] - g["raws_rows"]).abs().max().item() < 1e-4
    assert (feats - g["feats"]).abs().max().item() < 4e-6
    assert (fusion - g["fusion"]).abs().max().item() < 4e-6
    assert (sim - g["sim"]).abs().max().item() < 4e-6


def test_recall_full_golden_restatement_labels_and_rerank():
    """tests/golden/recall_full_L.pt (full-depth ViT-L + 12-layer Q-Former, 256 structured images, 512 queries; the
    reference's own similarity, labels planted from it with margin 2e-3, full-depth inference_rerank for 2 x 16 pairs):
    the restatement reproduces a sample of the similarity matrix and the rerank probabilities at FULL depth, the
    stored labels are what `plant_targets_with_margin` gives, each label really has the margin at every cut-off, and
    the stored recalls are the recall tail's on those labels."""
    g = torch.load(os.path.join(GOLDEN, "recall_full_L.pt"))
    c = g["case"]
    sim_ref, ref = g["sim"], g["ref_rows"]
    assert sim_ref.shape == (512, 256) and sim_ref.max() - sim_ref.min() > 0.15
    tgt, ranks, members = R.plant_targets_with_margin(sim_ref, ref, g["margin"])
    assert torch.equal(tgt, g["target"]) and torch.equal(ranks, g["ranks"]) and torch.equal(members, g["members"])
    assert torch.equal(R.target_ranks(sim_ref, ref, tgt), ranks)
    rec = R.cirr_recalls(R.ranking(sim_ref), ref, tgt, members)
    assert rec == pytest.approx(tuple(float(x) for x in g["recalls_ref"]))
    assert all(r > 0 for r in rec) and 30 < rec[4] < 70 and rec[6] > 85
    masked = sim_ref.clone()
    masked[torch.arange(512), ref] = float("-inf")
    srt = masked.sort(dim=1, descending=True).values
    s_t = sim_ref[torch.arange(512), tgt]
    for K in (1, 5, 10, 50):
        other = torch.where(ranks <= K, srt[:, K], srt[:, K - 1])
        assert float((s_t - other).abs().min()) >= g["margin"] - 1e-7
    # full depth through the restatement: the images of the first 6 queries' references + the rerank pairs
    rr = g["rerank"]
    q = torch.arange(6)
    need = torch.cat([ref[q], rr["ref_rows"], rr["cand_rows"], torch.arange(4)])
    uniq, inv = torch.unique(need, return_inverse=True)
    images = synth.make_structured_images(c["n_images"], seed=c["image_seed"])[uniq]
    sd = synth.make_state_dict(c["vit"], None, 12, seed=0)
    with torch.no_grad():
        feats, raws = R.extract_target_features(sd, images)
        sim = R.inference(sd, raws[inv[:6]], feats, g["input_ids"][q], g["attention_mask"][q])
        sd_r = dict(sd)
        sd_r["itm_head.weight"] = sd["itm_head.weight"] * c["itm_scale"]
        R_, T_ = rr["R"], rr["T"]
        p = R.inference_rerank(sd_r, raws[inv[6:6 + R_]], raws[inv[6 + R_:6 + R_ + R_ * T_]], g["input_ids"][:R_],
                               g["attention_mask"][:R_])
    assert (sim - sim_ref[q][:, uniq]).abs().max().item() < 1e-5
    assert (feats[inv[-4:]] - g["feats_rows"]).abs().max().item() < 4e-6
    assert (p - rr["p"]).abs().max().item() < 1e-5 and float(rr["p"].std()) > 1e-2
