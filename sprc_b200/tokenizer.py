"""Host-side text front end of the composed query (stays on the CPU).

* `BlipCaptionProcessor` mirrors lavis/processors/blip_processors.py:28-68 (`txt_processors["eval"]`):
  lower-case, strip ``.!"()*#:;~``, collapse whitespace, cap at 50 words.
* `OfflineBertTokenizer` mirrors what the reference gets from
  ``BertTokenizer.from_pretrained("bert-base-uncased") + add_special_tokens({"bos_token": "[DEC]"})``
  (blip2_models/blip2.py:30-34) for the single call pattern the hot path uses
  (blip2_qformer_cir_align_prompt.py:323-329): pad to 32, truncate, return int64 ids + mask.
  The third-party algorithm is WordPiece from transformers==4.36.2 (requirements.txt:9): BERT basic
  tokenisation (lower-case, NFD accent stripping, punctuation splitting) then greedy longest-match-first
  sub-word lookup with the ``##`` continuation prefix, ``[UNK]`` for unmatched words (>100 chars too); CJK
  ideographs are split into single tokens and literal special tokens (``[SEP]`` ...) are kept whole, as the library
  does.  tests/test_host.py checks ids and masks against the library's own BertTokenizer on a generated vocabulary.
  The bert-base-uncased vocabulary is not available offline: pass ``vocab_file`` (or set
  ``SPRC_BERT_VOCAB``) to get real ids; without it a deterministic synthetic vocabulary maps every
  word to one id in [1000, 29999] by FNV-1a hash so that string-driven synthetic runs are reproducible.
"""
from __future__ import annotations

import os
import re
import unicodedata

import torch

PAD, UNK, CLS, SEP, MASK = 0, 100, 101, 102, 103
VOCAB_SIZE = 30522  # bert-base-uncased; +1 for [DEC] -> len(tokenizer) == 30523


class BlipCaptionProcessor:
    def __init__(self, prompt: str = "", max_words: int = 50):
        self.prompt = prompt
        self.max_words = max_words

    def __call__(self, caption: str) -> str:
        return self.prompt + self.pre_caption(caption)

    def pre_caption(self, caption: str) -> str:
        caption = re.sub(r"([.!\"()*#:;~])", " ", caption.lower())
        caption = re.sub(r"\s{2,}", " ", caption)
        caption = caption.rstrip("\n").strip(" ")
        words = caption.split(" ")
        if len(words) > self.max_words:
            caption = " ".join(words[: self.max_words])
        return caption


class TokenBatch:
    """What the model reads from a tokenizer call: `.input_ids`, `.attention_mask`, `.to(device)`."""

    def __init__(self, input_ids: torch.Tensor, attention_mask: torch.Tensor):
        self.input_ids = input_ids
        self.attention_mask = attention_mask

    def to(self, device):
        return TokenBatch(self.input_ids.to(device), self.attention_mask.to(device))


def _is_punct(ch: str) -> bool:
    cp = ord(ch)
    if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
        return True
    return unicodedata.category(ch).startswith("P")


def _is_cjk(ch: str) -> bool:
    """BasicTokenizer._is_chinese_char (transformers 4.36 tokenization_bert.py): CJK ideographs become single tokens."""
    cp = ord(ch)
    return (0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0x20000 <= cp <= 0x2A6DF or 0x2A700 <= cp <= 0x2B73F
            or 0x2B740 <= cp <= 0x2B81F or 0x2B820 <= cp <= 0x2CEAF or 0xF900 <= cp <= 0xFAFF
            or 0x2F800 <= cp <= 0x2FA1F)


# tokens the tokenizer never splits or lower-cases when they appear verbatim in the text (PreTrainedTokenizer.tokenize
# splits on all_special_tokens first); [DEC] is the bos token blip2.py:33 adds, id = vocabulary size
_SPECIAL = {"[PAD]": PAD, "[UNK]": UNK, "[CLS]": CLS, "[SEP]": SEP, "[MASK]": MASK, "[DEC]": VOCAB_SIZE}
_SPECIAL_RE = re.compile("(" + "|".join(re.escape(t) for t in _SPECIAL) + ")")


def _fnv1a(s: str) -> int:
    h = 0x811C9DC5
    for b in s.encode("utf-8"):
        h = ((h ^ b) * 0x01000193) & 0xFFFFFFFF
    return h


def _find_cached_vocab():
    """vocab.txt of bert-base-uncased in the local Hugging Face cache (what `BertTokenizer.from_pretrained` would read
    offline), or None."""
    roots = [os.environ.get("HF_HUB_CACHE"), os.path.join(os.environ.get("HF_HOME", ""), "hub") if os.environ.get(
        "HF_HOME") else None, os.path.expanduser("~/.cache/huggingface/hub")]
    for r in roots:
        if not r:
            continue
        base = os.path.join(r, "models--bert-base-uncased", "snapshots")
        if os.path.isdir(base):
            for snap in sorted(os.listdir(base)):
                f = os.path.join(base, snap, "vocab.txt")
                if os.path.isfile(f):
                    return f
    return None


class OfflineBertTokenizer:
    """vocab_file (or SPRC_BERT_VOCAB, or a bert-base-uncased vocab.txt found in the local Hugging Face cache): real
    WordPiece ids.  synthetic=True (or SPRC_SYNTHETIC_VOCAB=1): the hashed stand-in vocabulary, an explicit opt-in for
    tests and synthetic benchmarks.  With neither, tokenising a STRING raises (a real checkpoint must never be driven
    with made-up ids); pre-tokenised `TokenBatch` input always passes through.
    native=True: strings go through the threaded C++ tokenizer of libsprc_b200 (sprc_tokenize_host); captions it flags
    as neighbour-dependent (combining marks, final sigma) take the Python path below, which is the same algorithm on
    whole strings."""

    def __init__(self, vocab_file: str | None = None, max_word_chars: int = 100, synthetic: bool | None = None,
                 native: bool = True, threads: int = 0):
        vocab_file = vocab_file or os.environ.get("SPRC_BERT_VOCAB") or None
        if synthetic is None:
            synthetic = os.environ.get("SPRC_SYNTHETIC_VOCAB") == "1"
        if not vocab_file and not synthetic:
            vocab_file = _find_cached_vocab()
        self.vocab = None
        self.vocab_file = vocab_file
        if vocab_file:
            with open(vocab_file, encoding="utf-8") as f:
                self.vocab = {tok.rstrip("\n"): i for i, tok in enumerate(f)}
        self.synthetic = bool(synthetic) and self.vocab is None
        self.max_word_chars = max_word_chars
        self.bos_token_id = VOCAB_SIZE  # "[DEC]"
        self.native = native
        self.threads = threads
        self._nat = None

    def _require_vocab(self):
        if self.vocab is None and not self.synthetic:
            raise RuntimeError(
                "no BERT vocabulary: pass vocab_file=... (or set SPRC_BERT_VOCAB) with bert-base-uncased's vocab.txt; "
                "for synthetic runs opt in to the hashed stand-in vocabulary with synthetic=True / "
                "SPRC_SYNTHETIC_VOCAB=1 (the reference downloads the vocabulary, blip2.py:30-34; there is no network "
                "here)")

    def _native_handle(self):
        if self._nat is None:
            import ctypes

            from . import _lib as L

            lib = L.load()
            h = ctypes.c_void_p()
            if self.vocab is None:
                L.check(lib.sprc_tokenizer_create(None, 0, ctypes.byref(h)))
            else:
                with open(self.vocab_file, "rb") as f:
                    blob = f.read()
                if blob.endswith(b"\n"):
                    blob = blob[:-1]
                L.check(lib.sprc_tokenizer_create(blob, len(blob), ctypes.byref(h)))
            self._nat = (lib, h)
        return self._nat

    def __del__(self):
        nat = getattr(self, "_nat", None)
        if nat is not None:
            try:
                nat[0].sprc_tokenizer_destroy(nat[1])
            except Exception:
                pass
            self._nat = None

    def __len__(self) -> int:
        return VOCAB_SIZE + 1

    # ---- BERT BasicTokenizer -------------------------------------------------------------------
    def _basic(self, text: str):
        cleaned = []
        for ch in text:
            if ch in " \t\n\r":
                cleaned.append(" ")
            elif ord(ch) in (0, 0xFFFD) or unicodedata.category(ch).startswith("C"):
                continue   # _clean_text / _is_control: every C* category (Cc, Cf, Cn, Co, Cs) except \t \n \r
            elif ch.isspace():
                cleaned.append(" ")   # _is_whitespace (Zs) and the separators str.split() also splits on
            elif _is_cjk(ch):
                cleaned.append(" " + ch + " ")
            else:
                cleaned.append(ch)
        text = "".join(cleaned)
        out = []
        for tok in text.strip().split():
            tok = unicodedata.normalize("NFD", tok.lower())
            tok = "".join(ch for ch in tok if unicodedata.category(ch) != "Mn")
            cur = ""
            for ch in tok:
                if _is_punct(ch):
                    if cur:
                        out.append(cur)
                        cur = ""
                    out.append(ch)
                else:
                    cur += ch
            if cur:
                out.append(cur)
        return out

    # ---- WordPiece -----------------------------------------------------------------------------
    def _wordpiece(self, word: str):
        if self.vocab is None:
            return [1000 + _fnv1a(word) % 29000]
        if len(word) > self.max_word_chars:
            return [UNK]
        ids, start = [], 0
        while start < len(word):
            end = len(word)
            cur = None
            while start < end:
                sub = word[start:end]
                if start > 0:
                    sub = "##" + sub
                if sub in self.vocab:
                    cur = self.vocab[sub]
                    break
                end -= 1
            if cur is None:
                return [UNK]
            ids.append(cur)
            start = end
        return ids

    def encode(self, text: str, max_length: int = 32):
        ids = [CLS]
        for seg in _SPECIAL_RE.split(text):
            if seg in _SPECIAL:
                ids.append(_SPECIAL[seg] if (self.vocab is None or seg == "[DEC]") else self.vocab.get(seg, UNK))
                continue
            for w in self._basic(seg):
                ids.extend(self._wordpiece(w))
        ids = ids[: max_length - 1] + [SEP]  # truncation keeps [CLS] ... [SEP]
        return ids

    def __call__(self, text, padding="max_length", truncation=True, max_length=32, return_tensors="pt"):
        if isinstance(text, TokenBatch):
            return text
        if isinstance(text, str):
            text = [text]
        self._require_vocab()
        n = len(text)
        if self.native and n > 0:
            return self._call_native(text, max_length)
        ids = torch.zeros(n, max_length, dtype=torch.long)
        mask = torch.zeros(n, max_length, dtype=torch.long)
        for i, t in enumerate(text):
            self._fill_row(ids, mask, i, t, max_length)
        return TokenBatch(ids, mask)

    def _fill_row(self, ids, mask, i, t, max_length):
        e = self.encode(t, max_length)
        ids[i].zero_()
        mask[i].zero_()
        ids[i, : len(e)] = torch.tensor(e, dtype=torch.long)
        mask[i, : len(e)] = 1   # by length, not by id: a literal "[PAD]" in the text is a live token

    def tokenize_into(self, text, ids: torch.Tensor, mask: torch.Tensor, lens: torch.Tensor | None = None):
        """Tokenise `text` (list of n strings) straight into caller-owned HOST buffers: ids / mask int64 [n, max_len]
        (e.g. the pinned staging buffers of `query_topk_host_submit`), lens int32 [n] (optional).  C++ threads; no
        allocation of the big buffers per call."""
        self._require_vocab()
        n, max_length = ids.shape
        assert len(text) == n and mask.shape == ids.shape and ids.dtype == mask.dtype == torch.int64
        assert ids.is_contiguous() and mask.is_contiguous() and ids.device.type == "cpu"
        out = self._call_native(text, max_length, ids=ids, mask=mask)
        if lens is not None:
            lens.copy_(out.lens)
        return out

    def _call_native(self, text, max_length, pin=False, ids=None, mask=None):
        import numpy as np

        from . import _lib as L

        lib, h = self._native_handle()
        n = len(text)
        enc, py_rows = [], []
        for i, t in enumerate(text):
            try:
                enc.append(t.encode("utf-8"))
            except UnicodeEncodeError:      # lone surrogates: not representable, Python path
                enc.append(b"")
                py_rows.append(i)
        offs = np.zeros(n + 1, dtype=np.int64)
        np.cumsum([len(e) for e in enc], out=offs[1:])
        blob = b"".join(enc)
        if ids is None:
            ids = torch.empty(n, max_length, dtype=torch.long, pin_memory=pin)
            mask = torch.empty(n, max_length, dtype=torch.long, pin_memory=pin)
        lens = torch.empty(n, dtype=torch.int32)
        flags = torch.empty(n, dtype=torch.uint8)
        L.check(lib.sprc_tokenize_host(h, blob, offs.ctypes.data, n, max_length, self.threads, L.ptr(ids), L.ptr(mask),
                                       L.ptr(lens), L.ptr(flags)))
        for i in set(py_rows) | set(flags.nonzero().flatten().tolist()):
            self._fill_row(ids, mask, i, text[i], max_length)
            lens[i] = int(mask[i].sum())
        out = TokenBatch(ids, mask)
        out.lens = lens
        return out
