"""Drop-in proof on the GPU: the reference's UNCHANGED scripts (byte-identical copies staged by
tools/stage_reference_scripts.py into baseline/_ref/src) run against our `lavis.models` replacement on a
generated CIRR / FashionIQ tree.

Mode A: reference's own utils.py / validate_blip.py loops call our model's `extract_target_features` /
        `inference` (full [Bq,N] similarity returned, their argsort + string matching computes recalls).
Mode B: sprc_b200/dropin_fast shadows utils / validate_blip with the fused scan + top-k drivers.
Targets are planted from our own ranking at known ranks, so the recalls the REFERENCE code prints are
known in advance (and keep every recall > 0); Mode A and Mode B must print the same JSON.
"""
import json
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref", "src")
DROPIN = os.path.join(ROOT, "sprc_b200", "dropin")
DROPIN_FAST = os.path.join(ROOT, "sprc_b200", "dropin_fast")
# no bert-base-uncased vocabulary offline: the scripts' caption strings use the hashed stand-in vocabulary (explicit opt-in)
DEPTH_ENV = {"SPRC_VIT_DEPTH": "2", "SPRC_QF_LAYERS": "2", "SPRC_MAX_IMAGES": "64", "SPRC_MAX_QUERIES": "32",
             "SPRC_SYNTHETIC_VOCAB": "1"}

N_GALLERY, N_QUERIES = 56, 16
PLANT_RANKS = [1, 1, 1, 1, 3, 3, 4, 5, 7, 8, 9, 10, 20, 30, 40, 52]  # rank of the target AFTER reference removal


def _write_png(path, rng, w=240, h=200):
    from PIL import Image

    Image.fromarray(rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)).save(path)


def _words(rng, n):
    vocab = ["red", "blue", "longer", "sleeves", "dog", "two", "remove", "add", "darker", "stripes", "instead",
             "of", "the", "is", "more", "and", "with", "a", "shorter", "bright"]
    return " ".join(rng.choice(vocab, size=n).tolist())


@pytest.fixture(scope="module")
def staged_tree(tmp_path_factory):
    if not os.path.exists(os.path.join(STAGED, "blip_validate.py")):
        pytest.skip("reference driver scripts are not staged (baseline/_ref/src; build() stages them when "
                    "/root/reference exists)")
    from oracle import synth
    from sprc_b200.model import Blip2QformerCirAlignPrompt

    root = str(tmp_path_factory.mktemp("sprc_dropin"))
    shutil.copytree(STAGED, os.path.join(root, "src"), ignore=shutil.ignore_patterns("lavis"))  # driver scripts only
    os.environ["SPRC_SYNTHETIC_VOCAB"] = "1"
    rng = np.random.default_rng(0)
    names = [f"dev-{i:03d}-img{i % 3}" for i in range(N_GALLERY)]
    # ---- images: CIRR dev split + FashionIQ images (same pixels, two directory layouts) ----
    os.makedirs(os.path.join(root, "cirr_dataset", "dev"))
    os.makedirs(os.path.join(root, "cirr_dataset", "cirr", "captions"))
    os.makedirs(os.path.join(root, "cirr_dataset", "cirr", "image_splits"))
    os.makedirs(os.path.join(root, "fashionIQ_dataset", "images"))
    os.makedirs(os.path.join(root, "fashionIQ_dataset", "captions"))
    os.makedirs(os.path.join(root, "fashionIQ_dataset", "image_splits"))
    for n in names:
        p = os.path.join(root, "cirr_dataset", "dev", n + ".png")
        _write_png(p, rng)
        shutil.copyfile(p, os.path.join(root, "fashionIQ_dataset", "images", n + ".png"))
    split = {n: f"./dev/{n}.png" for n in names}
    for sp in ("val", "test1"):
        with open(os.path.join(root, "cirr_dataset", "cirr", "image_splits", f"split.rc2.{sp}.json"), "w") as f:
            json.dump(split, f)
        with open(os.path.join(root, "cirr_dataset", "cirr", "captions", f"cap.rc2.{sp}.json"), "w") as f:
            json.dump([], f)  # placeholder: CIRRDataset opens it even in 'classic' mode (data_utils.py:235)
    # ---- checkpoint (truncated ViT-L, synthetic weights) ----
    sd = synth.make_state_dict("clip_L", 2, 2, seed=0)
    ckpt = os.path.join(root, "ckpt.pt")
    torch.save({"epoch": 0, "Blip2QformerCirAlignPrompt": sd,
                "Blip2QformerCirCat": {k: v for k, v in sd.items() if k != "prompt_tokens"}}, ckpt)
    # ---- plant targets from OUR ranking over the images as the reference's data pipeline decodes them ----
    sys.path.insert(0, os.path.join(root, "src"))
    try:
        import importlib

        du = importlib.import_module("data_utils")
        importlib.reload(du)
        pre = du.targetpad_transform(1.25, 224)
        ds = du.CIRRDataset("val", "classic", pre)
        imgs = torch.stack([ds[i][1] for i in range(len(ds))])
        assert [ds[i][0] for i in range(len(ds))] == names
    finally:
        sys.path.remove(os.path.join(root, "src"))
        sys.modules.pop("data_utils", None)
    model = Blip2QformerCirAlignPrompt(vit_model="clip_L", device="cuda:0", max_images=64, max_queries=32,
                                       vit_depth=2, qf_layers=2)
    model.load_state_dict(sd)
    feats, raws = model.extract_target_features(imgs.cuda())
    from sprc_b200.tokenizer import BlipCaptionProcessor

    proc = BlipCaptionProcessor()
    ref_idx = rng.choice(N_GALLERY, size=N_QUERIES, replace=False)
    captions = [_words(rng, int(rng.integers(3, 9))).capitalize() + "." for _ in range(N_QUERIES)]
    sim = model.inference(raws[torch.as_tensor(ref_idx).cuda()], feats, [proc(c) for c in captions]).cpu()
    order = torch.argsort(1 - sim, dim=-1)
    cirr, expected = [], {"ranks": [], "granks": []}
    for j in range(N_QUERIES):
        ranked = [int(x) for x in order[j] if int(x) != int(ref_idx[j])]
        tgt = ranked[PLANT_RANKS[j] - 1]
        # 6 group members: reference + target + 4 others; the target's subset rank is known from `ranked`
        others = [x for x in ranked if x != tgt][j % 5:: 9][:4]
        members = [int(ref_idx[j]), tgt] + others
        sub = [x for x in ranked if x in members]
        expected["ranks"].append(PLANT_RANKS[j])
        expected["granks"].append(sub.index(tgt) + 1)
        rng.shuffle(members)
        cirr.append({"pairid": j, "reference": names[ref_idx[j]], "target_hard": names[tgt], "caption": captions[j],
                     "img_set": {"members": [names[m] for m in members]}})
    for sp in ("val", "test1"):
        with open(os.path.join(root, "cirr_dataset", "cirr", "captions", f"cap.rc2.{sp}.json"), "w") as f:
            json.dump(cirr, f)
    # FashionIQ: three categories share the gallery split; captions are pairs
    fiq_expected = {}
    for ci, cat in enumerate(("dress", "toptee", "shirt")):
        trip = []
        ranks = []
        caps2 = [(_words(rng, 3), _words(rng, 4)) for _ in range(N_QUERIES)]
        joined = [proc(f"{a.strip('.?, ').capitalize()} and {b.strip('.?, ')}") for a, b in caps2]
        s2 = model.inference(raws[torch.as_tensor(ref_idx).cuda()], feats, joined).cpu()
        o2 = torch.argsort(1 - s2, dim=-1)
        for j in range(N_QUERIES):
            r = [1, 5, 10, 11, 30, 50, 51, 56][(j + ci) % 8]
            tgt = int(o2[j][r - 1])
            ranks.append(r)
            trip.append({"candidate": names[ref_idx[j]], "target": names[tgt], "captions": list(caps2[j])})
        fiq_expected[cat] = ranks
        with open(os.path.join(root, "fashionIQ_dataset", "captions", f"cap.{cat}.val.json"), "w") as f:
            json.dump(trip, f)
        with open(os.path.join(root, "fashionIQ_dataset", "image_splits", f"split.{cat}.val.json"), "w") as f:
            json.dump(names, f)
    del model
    torch.cuda.empty_cache()
    return dict(root=root, ckpt=ckpt, expected=expected, fiq_expected=fiq_expected, names=names)


def _run(tree, script, args, fast):
    env = dict(os.environ)
    env.update(DEPTH_ENV)
    paths = ([DROPIN_FAST] if fast else []) + [DROPIN, ROOT]
    env["PYTHONPATH"] = os.pathsep.join(paths)
    cmd = [sys.executable, os.path.join(tree["root"], "src", script)] + args
    if fast:
        # python puts the script's own directory first on sys.path, which would pick the staged utils.py /
        # validate_blip.py; run the unchanged file through runpy with the fast modules ahead of it
        code = ("import sys, runpy; sys.path.insert(0, %r); sys.path.insert(1, %r); sys.argv = %r; "
                "runpy.run_path(%r, run_name='__main__')") % (
                    DROPIN_FAST, os.path.join(tree["root"], "src"), [script] + args,
                    os.path.join(tree["root"], "src", script))
        cmd = [sys.executable, "-c", code]
    p = subprocess.run(cmd, env=env, cwd=tree["root"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + "\n" + p.stderr[-3000:]
    return p.stdout


def _last_json(stdout):
    end = stdout.rfind("}")
    start = stdout.rfind("{", 0, end)
    return json.loads(stdout[start:end + 1])


def _pct(ranks, k):
    return 100.0 * sum(r <= k for r in ranks) / len(ranks)


@pytest.mark.parametrize("fast", [False, True], ids=["modeA_reference_loops", "modeB_fused_scan"])
def test_blip_validate_cirr_unchanged_script(staged_tree, fast):
    out = _run(staged_tree, "blip_validate.py", ["--dataset", "CIRR", "--blip-model-name", "blip2_cir_align_prompt",
                                                 "--backbone", "pretrain_vitL", "--model-path", staged_tree["ckpt"]],
               fast)
    assert "Missing keys []" in out
    res = _last_json(out)
    e = staged_tree["expected"]
    for k in (1, 5, 10, 50):
        assert res[f"recall_at{k}"] == pytest.approx(_pct(e["ranks"], k), abs=0.05), (k, res)
    for k in (1, 2, 3):
        assert res[f"group_recall_at{k}"] == pytest.approx(_pct(e["granks"], k), abs=0.05), (k, res)


def test_blip_validate_default_model_name_cir_cat(staged_tree):
    """`blip2_cir_cat` (the default --blip-model-name of cirr_test_submission.py:206) resolves to our
    Blip2QformerCirCat; dividing by temp does not change the ranking, so the recalls are the same."""
    out = _run(staged_tree, "blip_validate.py", ["--dataset", "CIRR", "--blip-model-name", "blip2_cir_cat",
                                                 "--backbone", "pretrain_vitL", "--model-path", staged_tree["ckpt"]],
               False)
    assert "Missing keys []" in out
    res = _last_json(out)
    e = staged_tree["expected"]
    for k in (1, 5, 10, 50):
        assert res[f"recall_at{k}"] == pytest.approx(_pct(e["ranks"], k), abs=0.05), (k, res)


@pytest.mark.parametrize("fast", [False, True], ids=["modeA_reference_loops", "modeB_fused_scan"])
def test_blip_validate_fashioniq_unchanged_script(staged_tree, fast):
    out = _run(staged_tree, "blip_validate.py", ["--dataset", "fashionIQ", "--blip-model-name",
                                                 "blip2_cir_align_prompt", "--backbone", "pretrain_vitL",
                                                 "--model-path", staged_tree["ckpt"]], fast)
    res = _last_json(out)
    fe = staged_tree["fiq_expected"]
    for cat in ("dress", "toptee", "shirt"):
        assert res[f"{cat}_recall_at10"] == pytest.approx(_pct(fe[cat], 10), abs=0.05), res
        assert res[f"{cat}_recall_at50"] == pytest.approx(_pct(fe[cat], 50), abs=0.05), res


def test_cirr_test_submission_unchanged_script(staged_tree):
    _run(staged_tree, "cirr_test_submission.py",
         ["--blip-model-name", "blip2_cir_align_prompt", "--backbone", "pretrain_vitL", "--model-path",
          staged_tree["ckpt"]], False)
    sub_dir = os.path.join(staged_tree["root"], "submission", "CIRR")
    files = sorted(os.listdir(sub_dir))
    assert len(files) == 2, files
    for f in files:
        with open(os.path.join(sub_dir, f)) as fh:
            d = json.load(fh)
        assert d["version"] == "rc2"
        rows = {k: v for k, v in d.items() if k not in ("version", "metric")}
        assert len(rows) == N_QUERIES
        want = 3 if d["metric"] == "recall_subset" else 50
        for v in rows.values():
            assert set(v) <= set(staged_tree["names"]) and len(set(v)) == len(v) == want


def test_cirr_submission_on_integer_rows_equals_reference_script(staged_tree):
    """`retrieval.generate_cirr_test_dicts` (top-51 rows + 6 subset scores, no [Q,N] matrix, no string
    compares) must write the same two dicts as the reference's unchanged cirr_test_submission.py (Mode A above)."""
    sub_dir = os.path.join(staged_tree["root"], "submission", "CIRR")
    if not os.path.isdir(sub_dir) or len(os.listdir(sub_dir)) != 2:
        _run(staged_tree, "cirr_test_submission.py",
             ["--blip-model-name", "blip2_cir_align_prompt", "--backbone", "pretrain_vitL", "--model-path",
              staged_tree["ckpt"]], False)
    want = {}
    for f in os.listdir(sub_dir):
        with open(os.path.join(sub_dir, f)) as fh:
            d = json.load(fh)
        want[d["metric"]] = {k: v for k, v in d.items() if k not in ("version", "metric")}
    import importlib

    from sprc_b200 import retrieval as R
    from sprc_b200.model import Blip2QformerCirAlignPrompt
    from sprc_b200.tokenizer import BlipCaptionProcessor

    src = os.path.join(staged_tree["root"], "src")
    sys.path.insert(0, src)
    try:
        du = importlib.import_module("data_utils")
        importlib.reload(du)
        pre = du.targetpad_transform(1.25, 224)
        classic = du.CIRRDataset("test1", "classic", pre)
        relative = du.CIRRDataset("test1", "relative", pre)
        model = Blip2QformerCirAlignPrompt(vit_model="clip_L", device="cuda:0", max_images=64, max_queries=32,
                                           vit_depth=2, qf_layers=2)
        model.load_state_dict(torch.load(staged_tree["ckpt"])["Blip2QformerCirAlignPrompt"])
        feats, names = R.extract_index_blip_features(classic, model)
        txt = {"eval": BlipCaptionProcessor()}
        got_g, got_s = R.generate_cirr_test_dicts(relative, model, feats, names, txt, rerank=False)
        # the reference ranks by argsort(1 - sim) in fp32 (cirr_test_submission.py:80-85): similarities a few ulps apart
        # at ~0.16 collapse to ONE distance at ~0.84, and its (unstable) argsort orders such ties arbitrarily.  Our rows
        # are ordered by the similarity itself, so the two lists must agree exactly as sequences of those distance keys.
        items = [relative[i] for i in range(len(relative))]
        index = feats.index
        tok = model._tokenize([txt["eval"](it[2]) for it in items])
        fusion = model.encode_query(index.raws, tok.input_ids, tok.attention_mask,
                                    ref_rows=index.rows_of([it[1] for it in items]))
        _, _, full = model.sim_topk(fusion, index.feats, k=0, want_full=True)
        dist = (1 - full.reshape(len(items), -1)).cpu()
        key = {str(it[0]): {n: float(dist[j, index.name_to_row[n]]) for n in names} for j, it in enumerate(items)}
    finally:
        sys.path.remove(src)
        sys.modules.pop("data_utils", None)
    n_tie_swaps = 0
    for got, ref in ((got_g, want["recall"]), (got_s, want["recall_subset"])):
        assert got.keys() == ref.keys()
        for pid in ref:
            if got[pid] != ref[pid]:
                n_tie_swaps += 1
                assert [key[pid][n] for n in got[pid]] == [key[pid][n] for n in ref[pid]], (pid, got[pid], ref[pid])
    print(f"[submission] lists differing only inside ties of the reference's fp32 distance: {n_tie_swaps}")


def test_index_build_with_gpu_preprocess_equals_pil_path(staged_tree):
    """the reference's CIRRDataset with `preprocess=retrieval.raw_rgb` + the GPU preprocessor builds the
    same index (bit-identical features) as with the reference's own `targetpad_transform` in the DataLoader workers."""
    import importlib

    from sprc_b200 import retrieval as R
    from sprc_b200.model import Blip2QformerCirAlignPrompt
    from sprc_b200.preprocess import TargetPadPreprocessor

    src = os.path.join(staged_tree["root"], "src")
    sys.path.insert(0, src)
    try:
        du = importlib.import_module("data_utils")
        importlib.reload(du)
        model = Blip2QformerCirAlignPrompt(vit_model="clip_L", device="cuda:0", max_images=64, max_queries=32,
                                           vit_depth=2, qf_layers=2)
        model.load_state_dict(torch.load(staged_tree["ckpt"])["Blip2QformerCirAlignPrompt"])
        a = R.build_index(du.CIRRDataset("val", "classic", du.targetpad_transform(1.25, 224)), model, num_workers=0)
        b = R.build_index(du.CIRRDataset("val", "classic", R.raw_rgb), model, num_workers=0,
                          gpu_preprocess=TargetPadPreprocessor(1.25, 224, device="cuda:0"))
    finally:
        sys.path.remove(src)
        sys.modules.pop("data_utils", None)
    assert a.names == b.names == staged_tree["names"]
    assert torch.equal(a.feats, b.feats) and torch.equal(a.raws, b.raws)


def test_index_build_with_native_png_feeder_equals_pil_path(staged_tree, tmp_path):
    """SURVEY §8f N2, input side: the index built from file NAMES (`preprocess=retrieval.image_path`: native threaded PNG
    decode into a pinned arena + GPU resize, no DataLoader workers) is bit-identical to the one the reference's chain
    builds (`PIL.Image.open` + `targetpad_transform` in DataLoader workers, utils.py:54-64, data_utils.py:91-105), on a
    folder that mixes what the datasets hold (RGB of many sizes) with what they might (gray, palette, alpha, 16-bit, a
    JPEG, a corrupt file: Pillow's own chain for those, dropped where it raises)."""
    import importlib

    import PIL.Image

    from sprc_b200 import retrieval as R
    from sprc_b200.model import Blip2QformerCirAlignPrompt
    from sprc_b200.preprocess import PngIndexFeeder, TargetPadPreprocessor

    rng = np.random.default_rng(11)
    folder = tmp_path / "imgs"
    folder.mkdir()
    files = []
    for i in range(70):
        h, w = int(rng.integers(60, 420)), int(rng.integers(60, 520))
        px = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
        kind = ("RGB", "RGB", "RGB", "RGB", "L", "RGB", "P", "RGBA", "I16", "JPG")[i % 10]
        p = str(folder / f"im{i:03d}.png")
        if kind == "RGB":
            PIL.Image.fromarray(px[..., :3], "RGB").save(p)
        elif kind == "L":
            PIL.Image.fromarray(px[..., 0], "L").save(p)
        elif kind == "P":
            PIL.Image.fromarray(px[..., :3], "RGB").quantize(64).save(p)
        elif kind == "RGBA":
            PIL.Image.fromarray(px, "RGBA").save(p)
        elif kind == "I16":
            PIL.Image.fromarray(rng.integers(0, 65535, (h, w)).astype(np.uint16)).save(p)
        else:
            PIL.Image.fromarray(px[..., :3], "RGB").save(p, format="JPEG")
        files.append(p)
    bad = str(folder / "im999.png")
    data = bytearray(open(files[0], "rb").read())
    data[data.index(b"IDAT") + 30] ^= 0x55
    open(bad, "wb").write(bytes(data))
    files.insert(17, bad)

    class Folder:   # the datasets' classic-mode protocol (data_utils.py:253-270: name, preprocess(Image.open(path)); None on error)
        def __init__(self, preprocess):
            self.preprocess = preprocess

        def __len__(self):
            return len(files)

        def __getitem__(self, i):
            try:
                return os.path.basename(files[i])[:-4], self.preprocess(PIL.Image.open(files[i]))
            except Exception as e:  # noqa: BLE001
                print(f"Exception: {e}")
                return None

    src = os.path.join(staged_tree["root"], "src")
    sys.path.insert(0, src)
    try:
        du = importlib.import_module("data_utils")
        importlib.reload(du)
        ref_tf = du.targetpad_transform(1.25, 224)
        model = Blip2QformerCirAlignPrompt(vit_model="clip_L", device="cuda:0", max_images=32, max_queries=8,
                                           vit_depth=2, qf_layers=2)
        model.load_state_dict(torch.load(staged_tree["ckpt"])["Blip2QformerCirAlignPrompt"])
        from torch.utils.data.dataloader import default_collate

        a = R.build_index(Folder(ref_tf), model, batch_size=16, num_workers=0,
                          collate_fn=lambda b: default_collate([x for x in b if x is not None]))
        feeder = PngIndexFeeder(TargetPadPreprocessor(1.25, 224, device="cuda:0"), fallback=ref_tf, threads=4)
        b = R.build_index(Folder(R.image_path), model, batch_size=16, png_feeder=feeder)
    finally:
        sys.path.remove(src)
        sys.modules.pop("data_utils", None)
    print(f"\n[png feeder] native {feeder.n_native}, Pillow chain {feeder.n_fallback}, dropped {feeder.n_dropped}")
    assert feeder.n_dropped == 1 and feeder.n_native == 42 and feeder.n_fallback == 28
    assert a.names == b.names and "im999" not in a.names and len(a.names) == 70
    assert torch.equal(a.feats, b.feats) and torch.equal(a.raws, b.raws)
