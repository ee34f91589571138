"""TEST INFRASTRUCTURE — generates tests/golden/*.pt by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_loader.py) on the seeded synthetic checkpoints / inputs
of oracle/synth.py.  Run in the build container only:  `python -m oracle.make_golden [case ...]`.

Each golden file holds the case description (so tests regenerate identical weights and inputs from
the seeds) and the reference's outputs at the stage boundaries of SURVEY.md §7 step 1:
  raws_rows  image_embeds_frozen[:, [0,1,128,256], :]   (ln_vision(ViT(x)), align_prompt.py:367-368)
  feats      image_features [B,32,256]                  (align_prompt.py:385)
  sim        inference(...) [Bq,N]                      (align_prompt.py:312-361)
  fusion     fusion_feats [Bq,256]                      (align_prompt.py:348-350; recomputed from the
                                                         reference's own sub-modules, same statements)
  rerank_p   inference_rerank(...) [R*T]                (rerank.py:399-445)
"""
from __future__ import annotations

import os
import sys
import time

import torch

from . import ref_loader, synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: vit, vit_depth (None = full), qf_layers, n_images, n_queries
    "tiny_L": dict(vit="clip_L", vit_depth=2, qf_layers=2, n_images=5, n_queries=3),
    "tiny_g": dict(vit="eva_clip_g", vit_depth=2, qf_layers=4, n_images=4, n_queries=3),
    "full_L": dict(vit="clip_L", vit_depth=None, qf_layers=12, n_images=4, n_queries=4),
    "full_g": dict(vit="eva_clip_g", vit_depth=None, qf_layers=12, n_images=3, n_queries=3),
}
RAW_ROWS = [0, 1, 128, 256]


def reference_fusion_feats(model, reference_embeds, ids, mask):
    """The first half of the reference's `inference`, statement for statement (align_prompt.py:314-350),
    executed on the reference's own sub-modules; `inference` itself only returns the similarity."""
    import torch.nn.functional as F

    image_atts = torch.ones(reference_embeds.size()[:-1], dtype=torch.long)
    query_tokens = model.query_tokens.expand(reference_embeds.shape[0], -1, -1)
    query_atts = torch.ones(query_tokens.size()[:-1], dtype=torch.long)
    attention_mask = torch.cat([query_atts, mask], dim=1)
    fusion_output = model.Qformer.bert(ids, query_embeds=query_tokens, attention_mask=attention_mask,
                                       encoder_hidden_states=reference_embeds, encoder_attention_mask=image_atts,
                                       return_dict=True)
    text_output = model.Qformer.bert(ids, query_embeds=fusion_output.last_hidden_state[:, :32, :],
                                     attention_mask=attention_mask, return_dict=True)
    return F.normalize(model.text_proj(text_output.last_hidden_state[:, 32, :]), dim=-1)


def run_case(name, cfg, check_keys=True):
    t0 = time.time()
    sd = synth.make_state_dict(cfg["vit"], cfg["vit_depth"], cfg["qf_layers"], seed=0)
    model = ref_loader.build_reference_model(cfg["vit"], seed=0, vit_depth=cfg["vit_depth"],
                                             qf_layers=cfg["qf_layers"])
    ref_keys = {k for k in model.state_dict() if not k.startswith("Qformer.cls.") and "position_ids" not in k}
    if check_keys:
        assert ref_keys == set(sd), (sorted(ref_keys - set(sd))[:5], sorted(set(sd) - ref_keys)[:5])
        for k, v in model.state_dict().items():
            if k in sd:
                assert tuple(v.shape) == tuple(sd[k].shape), (k, v.shape, sd[k].shape)
    msg = model.load_state_dict(sd, strict=False)
    assert not msg.unexpected_keys, msg.unexpected_keys
    images = synth.make_images(cfg["n_images"])
    ids, mask = synth.make_token_ids(cfg["n_queries"])
    ref_rows = torch.arange(cfg["n_queries"]) % cfg["n_images"]
    with torch.no_grad():
        feats, raws = model.extract_target_features(images)
        ref_embeds = raws[ref_rows]
        sim = ref_loader.call_inference(model, ref_embeds, feats, ids, mask)
        fusion = reference_fusion_feats(model, ref_embeds, ids, mask)
        # rerank: reference class shares every statement of the fusion pass; use the rerank model class
        # only for tiny cases (it instantiates a second Q-Former)
    out = dict(case=dict(cfg, name=name, seed=0), raw_rows=RAW_ROWS, raws_rows=raws[:, RAW_ROWS].clone(),
               feats=feats.clone(), sim=sim.reshape(cfg["n_queries"], -1).clone(), fusion=fusion.clone(),
               ref_rows=ref_rows, input_ids=ids, attention_mask=mask)
    if name.startswith("tiny"):
        out["rerank_p"] = run_rerank(cfg, sd, raws, ids, mask)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.save(out, os.path.join(GOLDEN_DIR, f"{name}.pt"))
    print(f"[golden] {name}: feats {tuple(feats.shape)} sim {tuple(sim.shape)} in {time.time()-t0:.1f}s", flush=True)
    return out


def run_rerank(cfg, sd, raws, ids, mask):
    """inference_rerank of the reference's rerank class with the same Q-Former / itm_head weights:
    R = 2 references x T = 2 candidates."""
    model = ref_loader.build_reference_model(cfg["vit"], seed=0, vit_depth=cfg["vit_depth"],
                                             qf_layers=cfg["qf_layers"], kind="rerank")
    model.load_state_dict(sd, strict=False)
    R, T = 2, 2
    ref = raws[[0, 1]]
    tgt = raws[[2, 3, 1, 2]]
    with torch.no_grad():
        p = model.inference_rerank(ref, tgt, ref_loader.TokenBatch(ids[:R], mask[:R]))
    return dict(R=R, T=T, ref_rows=torch.tensor([0, 1]), cand_rows=torch.tensor([2, 3, 1, 2]), p=p.clone())


def run_cat(name="tiny_L_cat", base="tiny_L"):
    """The scripts' default model `blip2_cir_cat` (blip2_qformer_cir_cat.py:282-336, 401-428) on the checkpoint and
    inputs of `base`: same fusion / text passes, similarity divided by the learned temperature."""
    cfg = CASES[base]
    sd = synth.make_state_dict(cfg["vit"], cfg["vit_depth"], cfg["qf_layers"], seed=0)
    sd = {k: v for k, v in sd.items() if k != "prompt_tokens"}   # Blip2QformerCirCat has no prompt tokens
    model = ref_loader.build_reference_model(cfg["vit"], seed=0, vit_depth=cfg["vit_depth"],
                                             qf_layers=cfg["qf_layers"], kind="cat")
    ref_keys = {k for k in model.state_dict() if not k.startswith("Qformer.cls.") and "position_ids" not in k}
    assert ref_keys == set(sd), (sorted(ref_keys - set(sd))[:5], sorted(set(sd) - ref_keys)[:5])
    msg = model.load_state_dict(sd, strict=False)
    assert not msg.unexpected_keys, msg.unexpected_keys
    images = synth.make_images(cfg["n_images"])
    ids, mask = synth.make_token_ids(cfg["n_queries"])
    ref_rows = torch.arange(cfg["n_queries"]) % cfg["n_images"]
    with torch.no_grad(), ref_loader.no_cuda_moves():
        feats, raws = model.extract_target_features(images)
        sim = ref_loader.call_inference(model, raws[ref_rows], feats, ids, mask)
        # its own inference_rerank (:337-398): R = 2 references x T = 2 candidate FEATURE blocks [R*T,32,256]
        rr = model.inference_rerank(raws[[0, 1]], feats[[2, 3, 1, 2]], ref_loader.TokenBatch(ids[:2], mask[:2]))
    out = dict(case=dict(cfg, name=name, seed=0, kind="cat"), feats=feats.clone(),
               sim=sim.reshape(cfg["n_queries"], -1).clone(), temp=float(sd["temp"]), ref_rows=ref_rows, input_ids=ids,
               attention_mask=mask,
               rerank=dict(R=2, T=2, ref_rows=torch.tensor([0, 1]), cand_rows=torch.tensor([2, 3, 1, 2]),
                           sim=rr.clone()))
    torch.save(out, os.path.join(GOLDEN_DIR, f"{name}.pt"))
    print(f"[golden] {name}: sim {tuple(out['sim'].shape)} temp {out['temp']}", flush=True)
    return out


def run_spread(name="spread_L"):
    """Recall@K parity case: the reference's `inference` similarity [Q,N] on a gain-2.5 checkpoint (synth.make_state_dict:
    similarities spread over ~0.2 instead of 0.016; larger gains make the network ill-conditioned — at gain 4 torch's
    own bf16 autocast of the reference moves similarities by 7e-2), N = 128 gallery images, Q = 64 composed queries.  The test plants the
    labels from THIS ranking (restatement.plant_targets) and compares the recalls of the CUDA path with it."""
    cfg = dict(vit="clip_L", vit_depth=2, qf_layers=2, n_images=128, n_queries=64, gain=2.5)
    sd = synth.make_state_dict(cfg["vit"], cfg["vit_depth"], cfg["qf_layers"], seed=0, gain=cfg["gain"])
    model = ref_loader.build_reference_model(cfg["vit"], seed=0, vit_depth=cfg["vit_depth"], qf_layers=cfg["qf_layers"])
    msg = model.load_state_dict(sd, strict=False)
    assert not msg.unexpected_keys, msg.unexpected_keys
    images = synth.make_images(cfg["n_images"])
    ids, mask = synth.make_token_ids(cfg["n_queries"])
    ref_rows = torch.randint(0, cfg["n_images"], (cfg["n_queries"],), generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        feats, raws = model.extract_target_features(images)
        sim = torch.cat([ref_loader.call_inference(model, raws[ref_rows[i:i + 16]], feats, ids[i:i + 16], mask[i:i + 16])
                         for i in range(0, cfg["n_queries"], 16)])
    out = dict(case=dict(cfg, name=name, seed=0), sim=sim.clone(), ref_rows=ref_rows, input_ids=ids, attention_mask=mask)
    torch.save(out, os.path.join(GOLDEN_DIR, f"{name}.pt"))
    print(f"[golden] {name}: sim {tuple(sim.shape)} range [{sim.min():.3f}, {sim.max():.3f}]", flush=True)
    return out


def run_recall_full(name="recall_full_L", n_images=256, n_queries=512, n_rerank=(2, 16)):
    """Full-depth pins: (i) the reference's own `inference` similarity [512, 256] with the FULL ViT-L (23
    blocks) and the full 12-layer Q-Former on the standard synthetic checkpoint, structured images
    (synth.make_structured_images: similarities spread over ~0.2), 512 composed queries; the GPU test plants labels
    from THIS matrix with margins (restatement.plant_targets_with_margin) and compares Recall@K; (ii) the reference's
    own `inference_rerank` probabilities for R = 2 queries x T = 16 candidates at the same full depth."""
    t0 = time.time()
    cfg = dict(vit="clip_L", vit_depth=None, qf_layers=12, n_images=n_images, n_queries=n_queries, image_seed=2468)
    sd = synth.make_state_dict(cfg["vit"], None, 12, seed=0)
    model = ref_loader.build_reference_model(cfg["vit"], seed=0)
    msg = model.load_state_dict(sd, strict=False)
    assert not msg.unexpected_keys, msg.unexpected_keys
    images = synth.make_structured_images(n_images, seed=cfg["image_seed"])
    ids, mask = synth.make_token_ids(n_queries, seed=4321)
    ref_rows = torch.randint(0, n_images, (n_queries,), generator=torch.Generator().manual_seed(7))
    feats, raws = [], []
    with torch.no_grad():
        for s in range(0, n_images, 64):                                      # utils.py:54 batch size
            f, r = model.extract_target_features(images[s:s + 64])
            feats.append(f)
            raws.append(r)
            print(f"[golden] {name}: indexed {s + 64}/{n_images} ({time.time() - t0:.0f}s)", flush=True)
        feats, raws = torch.vstack(feats), torch.vstack(raws)
        sim = torch.cat([ref_loader.call_inference(model, raws[ref_rows[i:i + 32]], feats, ids[i:i + 32],
                                                   mask[i:i + 32]).reshape(-1, n_images)
                         for i in range(0, n_queries, 32)])                    # validate_blip.py:373 batch size
    out = dict(case=dict(cfg, name=name, seed=0), sim=sim.clone(), ref_rows=ref_rows, input_ids=ids, attention_mask=mask,
               feats_rows=feats[:4].clone(), raws_rows=raws[:4][:, RAW_ROWS].clone(), raw_rows=RAW_ROWS)
    torch.save(out, os.path.join(GOLDEN_DIR, f"{name}.pt"))
    print(f"[golden] {name}: sim {tuple(sim.shape)} range [{sim.min():.3f}, {sim.max():.3f}] per-query std "
          f"{sim.std(dim=1).mean():.3e} in {time.time() - t0:.0f}s", flush=True)
    return finish_recall_full(name, n_rerank=n_rerank, raws=raws)


def finish_recall_full(name="recall_full_L", n_rerank=(2, 16), margin=2e-3, itm_scale=0.1, raws=None):
    """Second half of `run_recall_full` (can be re-run on an existing file): labels planted from the reference's
    similarity with margin 2e-3 = twice the north star's embedding tolerance (restatement.plant_targets_with_margin),
    the reference's recalls on them, and the full-depth `inference_rerank` probabilities.  The rerank part scales
    itm_head.weight by `itm_scale` (recorded in the case): at the checkpoint's 0.2-std head every probability saturates
    above 0.9 and a comparison of p would be blind to logit errors."""
    from . import restatement as R

    t0 = time.time()
    path = os.path.join(GOLDEN_DIR, f"{name}.pt")
    out = torch.load(path)
    cfg = out["case"]
    sim, ref_rows, ids, mask = out["sim"], out["ref_rows"], out["input_ids"], out["attention_mask"]
    target, ranks, members = R.plant_targets_with_margin(sim, ref_rows, margin)
    out.update(target=target, ranks=ranks, members=members, margin=margin,
               recalls_ref=torch.tensor(R.cirr_recalls(R.ranking(sim), ref_rows, target, members)))
    R_, T_ = n_rerank
    order = torch.argsort(1 - sim[:R_], dim=-1)
    cand_rows = order[:, :T_].reshape(-1)                                      # each query's own top-T, as the drivers do
    sd = synth.make_state_dict(cfg["vit"], None, 12, seed=0)
    sd["itm_head.weight"] = sd["itm_head.weight"] * itm_scale
    rr_model = ref_loader.build_reference_model(cfg["vit"], seed=0, kind="rerank")
    rr_model.load_state_dict(sd, strict=False)
    need = torch.cat([ref_rows[:R_], cand_rows])
    uniq, inv = torch.unique(need, return_inverse=True)
    with torch.no_grad():
        if raws is None:
            images = synth.make_structured_images(cfg["n_images"], seed=cfg["image_seed"])
            _, table = rr_model.extract_target_features(images[uniq])
        else:
            table = raws[uniq]
        p = rr_model.inference_rerank(table[inv[:R_]], table[inv[R_:]], ref_loader.TokenBatch(ids[:R_], mask[:R_]))
    out["case"] = dict(cfg, itm_scale=itm_scale)
    out["rerank"] = dict(R=R_, T=T_, ref_rows=ref_rows[:R_].clone(), cand_rows=cand_rows.clone(), p=p.clone())
    torch.save(out, path)
    print(f"[golden] {name}: planted ranks hist {torch.bincount(ranks.clamp_max(60))[:12].tolist()}..., recalls_ref "
          f"{[round(float(x), 2) for x in out['recalls_ref']]}; rerank p {p.min():.3f}..{p.max():.3f} "
          f"(spread {p.std():.3e}) in {time.time() - t0:.0f}s", flush=True)
    return out


if __name__ == "__main__":
    names = sys.argv[1:] or (list(CASES) + ["tiny_L_cat", "spread_L", "recall_full_L"])
    for n in names:
        if n == "tiny_L_cat":
            run_cat()
        elif n == "spread_L":
            run_spread()
        elif n == "recall_full_L":
            run_recall_full()
        elif n == "recall_full_L:finish":
            finish_recall_full()
        else:
            run_case(n, CASES[n])
