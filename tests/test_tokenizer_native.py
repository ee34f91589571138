"""The C++ tokenizer behind the C ABI (csrc/tokenizer.cpp, sprc_tokenize_host) against the third-party algorithm the
reference uses (transformers' slow BertTokenizer = `BertTokenizerLegacy` in transformers 5, the 4.36 source:
BasicTokenizer + WordPiece; blip2.py:30-34, align_prompt.py:323-329) and against the Python restatement in
sprc_b200/tokenizer.py.  Host code only: runs without a GPU."""
import random

import pytest
import torch

from sprc_b200 import synth
from sprc_b200.tokenizer import OfflineBertTokenizer


def _toy_vocab(tmp_path, rnd):
    alpha = "abcdefgh"
    pieces = {"".join(rnd.choice(alpha) for _ in range(rnd.randint(1, 4))) for _ in range(400)}
    pieces |= {"##" + "".join(rnd.choice(alpha) for _ in range(rnd.randint(1, 3))) for _ in range(300)}
    toks = ["[PAD]"] + [f"[unused{i}]" for i in range(99)] + ["[UNK]", "[CLS]", "[SEP]", "[MASK]"] + sorted(pieces)
    toks += list(",.!?;:'\"()-") + ["长", "é", "1", "2", "##1", "e", "u", "ss", "i", "豈", "σ", "ς", "a", "##a"]
    toks = list(dict.fromkeys(toks))
    vf = tmp_path / "vocab.txt"
    vf.write_text("\n".join(toks) + "\n", encoding="utf-8")
    return str(vf), toks


EXTRAS = [" ", "  ", "\t", "\n", ",", ".", "!", "-", "'", "(", ")", "长", "é", "É", "ü", " ", "​", "\x00",
          "�", "1", "12", "ß", "İ", "$", "^", "`", "~", "　", "x" * 120, "长a", "a长b", "́", "é",
          "\x7f", " ", "[", "]", "[pad]", "[SEP]", "[UNK]", "[MASK]", "[CLS]", "[PAD]", "a[SEP]b", "[DEC]",
          "", "͸", "Σ", "aΣ", "Å", "Å", "豈", "豈", "한", "ᅡ", "ཱི", "\U0001f600", "ः",
          "\x1c", "\x0b", "\x85", " ", " ", "¿", "«", "—", "Ǆ", "ǅ", "ﬁ", "Ω", "İ", "ẞ", "\U00010400",
          "­", "ـ", "־", "ª", "²", "㐀", "\U00020000", "[SE", "P]"]


def _random_texts(rnd, n):
    alpha = "abcdefgh"
    texts = []
    for _ in range(n):
        parts = []
        for _ in range(rnd.randint(0, 14)):
            if rnd.random() < 0.6:
                w = "".join(rnd.choice(alpha) for _ in range(rnd.randint(1, 9)))
                parts.append(w.capitalize() if rnd.random() < 0.3 else w)
            else:
                parts.append(rnd.choice(EXTRAS))
            if rnd.random() < 0.7:
                parts.append(" ")
        texts.append("".join(parts))
    return texts


def test_native_tokenizer_matches_the_library_and_the_python_path(tmp_path):
    tr = pytest.importorskip("transformers.models.bert.tokenization_bert_legacy")
    rnd = random.Random(0)
    vf, toks = _toy_vocab(tmp_path, rnd)
    ref = tr.BertTokenizerLegacy(vf)
    ref.add_special_tokens({"bos_token": "[DEC]"})
    nat = OfflineBertTokenizer(vf, native=True, threads=3)
    py = OfflineBertTokenizer(vf, native=False)
    texts = _random_texts(rnd, 4000) + ["", " ", "a" * 500, "é" * 40, "[SEP]" * 40]
    a = ref(texts, padding="max_length", truncation=True, max_length=32, return_tensors="pt")
    b = nat(texts, max_length=32)
    c = py(texts, max_length=32)
    fix = lambda ids: torch.where(ids == 30522, torch.full_like(ids, ref.bos_token_id), ids)  # noqa: E731
    bad = [(t, x.tolist(), y.tolist()) for t, x, y in zip(texts, a.input_ids, fix(b.input_ids)) if not torch.equal(x, y)]
    assert not bad, bad[:3]
    assert torch.equal(a.attention_mask, b.attention_mask)
    assert torch.equal(b.input_ids, c.input_ids) and torch.equal(b.attention_mask, c.attention_mask)
    assert torch.equal(b.lens.long(), b.attention_mask.sum(dim=1))


def test_native_tokenizer_decides_most_captions_itself(tmp_path):
    """The flag is for neighbour-dependent characters only: plain captions (ASCII, precomposed accents, CJK, Hangul,
    punctuation) never leave the C++ path, and a flagged caption still gets the library's ids via the Python path."""
    import ctypes

    import numpy as np

    from sprc_b200 import _lib as L

    rnd = random.Random(1)
    vf, _ = _toy_vocab(tmp_path, rnd)
    nat = OfflineBertTokenizer(vf)
    lib, h = nat._native_handle()
    texts = ["Is darker, and has a longer sleeve!", "café É ü ß", "长城 한국어", "tabs\tand\nnewlines", "x" * 120,
             "é decomposed", "final Σ sigma"]
    enc = [t.encode() for t in texts]
    offs = np.zeros(len(texts) + 1, dtype=np.int64)
    np.cumsum([len(e) for e in enc], out=offs[1:])
    ids = torch.empty(len(texts), 32, dtype=torch.long)
    mask = torch.empty_like(ids)
    lens = torch.empty(len(texts), dtype=torch.int32)
    flags = torch.empty(len(texts), dtype=torch.uint8)
    L.check(lib.sprc_tokenize_host(h, b"".join(enc), offs.ctypes.data, len(texts), 32, 1, L.ptr(ids), L.ptr(mask),
                                   L.ptr(lens), L.ptr(flags)))
    assert flags.tolist() == [0, 0, 0, 0, 0, 1, 1]
    assert lens[:5].min() >= 3 and lens[5:].tolist() == [-1, -1]
    # error contract
    z = ctypes.c_void_p(0)
    assert lib.sprc_tokenize_host(z, None, z, 1, 32, 1, z, z, z, z) == -22 and b"null" in lib.sprc_last_error()
    assert lib.sprc_tokenizer_create(None, 5, ctypes.byref(ctypes.c_void_p())) == -22


def test_synthetic_vocabulary_round_trip_and_hashed_mode():
    """Caption strings over the generated 30 522-token vocabulary give back exactly synth.make_token_ids (what the
    string-driven bench relies on); the hashed stand-in vocabulary is identical in C++ and Python."""
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        vf = synth.write_vocab(d + "/vocab.txt")
        caps = synth.make_captions(300, seed=11)
        ids, mask = synth.make_token_ids(300, seed=11)
        for native in (True, False):
            b = OfflineBertTokenizer(vf, native=native)(caps)
            assert torch.equal(b.input_ids, ids) and torch.equal(b.attention_mask, mask)
    texts = _random_texts(random.Random(2), 500)
    a = OfflineBertTokenizer(synthetic=True, native=True)(texts)
    b = OfflineBertTokenizer(synthetic=True, native=False)(texts)
    assert torch.equal(a.input_ids, b.input_ids) and torch.equal(a.attention_mask, b.attention_mask)


def test_strings_without_a_vocabulary_are_refused(monkeypatch):
    """no silent hashed ids for real checkpoints."""
    monkeypatch.delenv("SPRC_BERT_VOCAB", raising=False)
    monkeypatch.delenv("SPRC_SYNTHETIC_VOCAB", raising=False)
    monkeypatch.setenv("HF_HUB_CACHE", "/nonexistent")
    monkeypatch.setenv("HOME", "/nonexistent")
    t = OfflineBertTokenizer()
    with pytest.raises(RuntimeError, match="no BERT vocabulary"):
        t(["a red dress"])
    from sprc_b200.tokenizer import TokenBatch

    tb = TokenBatch(torch.zeros(1, 32, dtype=torch.long), torch.zeros(1, 32, dtype=torch.long))
    assert t(tb) is tb
    monkeypatch.setenv("SPRC_SYNTHETIC_VOCAB", "1")
    assert OfflineBertTokenizer()(["a red dress"]).input_ids[0, 0] == 101
