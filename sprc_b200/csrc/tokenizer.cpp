// Host-side caption tokenizer behind the C ABI (include/sprc_b200.h: sprc_tokenizer_*): what the reference's
// `inference` does at blip2_qformer_cir_align_prompt.py:323-329 with the tokenizer of blip2.py:30-34
// (transformers==4.36.2 BertTokenizer("bert-base-uncased") + [DEC]): BasicTokenizer (clean text, isolate CJK
// ideographs, whitespace split, lower-case, NFD + strip Mn, split on punctuation) then greedy longest-match-first
// WordPiece, [CLS] ... [SEP], truncate / pad to max_len.  The algorithm is the third-party library's; it is restated
// from its published source (tokenization_bert.py BasicTokenizer / WordpieceTokenizer) and pinned against the library's
// own legacy tokenizer in tests/test_tokenizer_native.py.
//
// Unicode: every per-character decision comes from tables generated from Python's `unicodedata`
// (tools/gen_unicode_tables.py -> build/unicode_tables.inc), so this code and the library consult the same data.
// Captions holding a character whose rewrite depends on its neighbours (final sigma, combining marks, characters NFC
// rewrites; class COMPLEX) are FLAGGED, not guessed: the caller (sprc_b200/tokenizer.py) runs its exact Python path
// for those.  Threads: captions are independent; `threads` workers take contiguous slices.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/sprc_b200.h"
#include "build/unicode_tables.inc"

namespace sprc { int set_error(int code, const char* fmt, ...); }  // runtime.cu
static void sprc_set_error(const char* msg) { sprc::set_error(-22, "%s", msg); }

namespace {

enum : int { PAD = 0, UNK = 100, CLS = 101, SEP = 102, MASK = 103, DEC = 30522 };
enum : int { C_WORD = 0, C_DROP = 1, C_SPACE = 2, C_CJK = 3, C_COMPLEX = 4 };

template <size_t N>
bool in_ranges(const uint32_t (&r)[N][2], uint32_t cp) {
  size_t lo = 0, hi = N;
  while (lo < hi) {
    size_t mid = (lo + hi) / 2;
    if (cp < r[mid][0]) hi = mid;
    else if (cp > r[mid][1]) lo = mid + 1;
    else return true;
  }
  return false;
}

int classify(uint32_t cp) {
  if (cp < 0x80) {
    if (cp == ' ' || cp == '\t' || cp == '\n' || cp == '\r') return C_SPACE;
    if (cp < 0x20 || cp == 0x7F) return C_DROP;  // cp 0 and the Cc block (\t \n \r were taken above)
    return C_WORD;
  }
  if (in_ranges(kDrop, cp)) return C_DROP;
  if (in_ranges(kSpace, cp)) return C_SPACE;
  if (in_ranges(kComplex, cp)) return C_COMPLEX;
  if (in_ranges(kCjk, cp)) return C_CJK;
  return C_WORD;
}

bool is_punct(uint32_t cp) {
  if (cp < 0x80) return (cp >= 33 && cp <= 47) || (cp >= 58 && cp <= 64) || (cp >= 91 && cp <= 96) || (cp >= 123 && cp <= 126);
  return in_ranges(kPunct, cp);
}

// strip_Mn(NFD(lower(cp))) appended to `out`
void rewrite(uint32_t cp, std::u32string& out) {
  if (cp < 0x80) {
    out.push_back((cp >= 'A' && cp <= 'Z') ? cp + 32 : cp);
    return;
  }
  if (cp >= 0xAC00 && cp <= 0xD7A3) {  // Hangul syllable -> conjoining jamo (Unicode 3.12)
    uint32_t s = cp - 0xAC00;
    out.push_back(0x1100 + s / 588);
    out.push_back(0x1161 + (s % 588) / 28);
    if (s % 28) out.push_back(0x11A7 + s % 28);
    return;
  }
  const size_t n = sizeof(kMap) / sizeof(kMap[0]);
  size_t lo = 0, hi = n;
  while (lo < hi) {
    size_t mid = (lo + hi) / 2;
    if (kMap[mid].cp < cp) lo = mid + 1;
    else hi = mid;
  }
  if (lo < n && kMap[lo].cp == cp) {
    for (uint32_t i = 0; i < kMap[lo].len; ++i) out.push_back(kMapPool[kMap[lo].off + i]);
    return;
  }
  out.push_back(cp);
}

void utf8_append(std::string& s, uint32_t cp) {
  if (cp < 0x80) s.push_back((char)cp);
  else if (cp < 0x800) { s.push_back((char)(0xC0 | (cp >> 6))); s.push_back((char)(0x80 | (cp & 0x3F))); }
  else if (cp < 0x10000) {
    s.push_back((char)(0xE0 | (cp >> 12))); s.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); s.push_back((char)(0x80 | (cp & 0x3F)));
  } else {
    s.push_back((char)(0xF0 | (cp >> 18))); s.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
    s.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); s.push_back((char)(0x80 | (cp & 0x3F)));
  }
}

// strict UTF-8 decode; returns false on malformed input
bool utf8_decode(const char* p, size_t n, std::u32string& out) {
  size_t i = 0;
  while (i < n) {
    uint8_t c = (uint8_t)p[i];
    uint32_t cp; int extra;
    if (c < 0x80) { cp = c; extra = 0; }
    else if ((c & 0xE0) == 0xC0) { cp = c & 0x1F; extra = 1; }
    else if ((c & 0xF0) == 0xE0) { cp = c & 0x0F; extra = 2; }
    else if ((c & 0xF8) == 0xF0) { cp = c & 0x07; extra = 3; }
    else return false;
    for (int k = 1; k <= extra; ++k) {
      if (i + k >= n) return false;
      uint8_t d = (uint8_t)p[i + k];
      if ((d & 0xC0) != 0x80) return false;
      cp = (cp << 6) | (d & 0x3F);
    }
    if (cp > 0x10FFFF) return false;
    out.push_back(cp);
    i += extra + 1;
  }
  return true;
}

uint32_t fnv1a(const std::string& s) {
  uint32_t h = 0x811C9DC5u;
  for (unsigned char b : s) h = (h ^ b) * 0x01000193u;
  return h;
}

}  // namespace

struct sprc_tokenizer {
  bool synthetic = false;
  std::unordered_map<std::string, int32_t> vocab;
  int special_id[6] = {PAD, UNK, CLS, SEP, MASK, DEC};
  int max_word_chars = 100;

  void wordpiece(const std::u32string& w, std::vector<int32_t>& ids) const {
    if (synthetic) {
      std::string s;
      for (uint32_t cp : w) utf8_append(s, cp);
      ids.push_back(1000 + (int32_t)(fnv1a(s) % 29000u));
      return;
    }
    if ((int)w.size() > max_word_chars) { ids.push_back(UNK); return; }
    // byte offsets of every character boundary, so sub-strings are slices of one UTF-8 buffer
    std::string s;
    std::vector<uint32_t> off(w.size() + 1);
    for (size_t i = 0; i < w.size(); ++i) { off[i] = (uint32_t)s.size(); utf8_append(s, w[i]); }
    off[w.size()] = (uint32_t)s.size();
    size_t first = ids.size(), start = 0;
    std::string key;
    while (start < w.size()) {
      size_t end = w.size();
      int32_t cur = -1;
      while (start < end) {
        key.assign(start > 0 ? "##" : "");
        key.append(s, off[start], off[end] - off[start]);
        auto it = vocab.find(key);
        if (it != vocab.end()) { cur = it->second; break; }
        --end;
      }
      if (cur < 0) { ids.resize(first); ids.push_back(UNK); return; }
      ids.push_back(cur);
      start = end;
    }
  }

  // BasicTokenizer + WordPiece over one segment holding no literal special token; returns false if COMPLEX
  bool segment(const std::u32string& t, std::vector<int32_t>& ids, int limit) const {
    std::u32string word;
    auto flush = [&]() {
      if (!word.empty()) { wordpiece(word, ids); word.clear(); }
    };
    for (uint32_t cp : t) {
      if ((int)ids.size() >= limit && word.empty()) break;  // everything past the truncation point is dropped anyway
      int c = classify(cp);
      if (c == C_COMPLEX) return false;
      if (c == C_DROP) continue;
      if (c == C_SPACE) { flush(); continue; }
      size_t before = word.size();
      if (c == C_CJK) { flush(); before = 0; }
      rewrite(cp, word);
      // punctuation produced by the rewrite splits the word: every punctuation character is its own token
      std::u32string tail(word.begin() + before, word.end());
      word.resize(before);
      for (uint32_t r : tail) {
        if (is_punct(r)) {
          flush();
          word.push_back(r);
          flush();
        } else {
          word.push_back(r);
        }
      }
      if (c == C_CJK) flush();
    }
    flush();
    return true;
  }

  // returns the number of live tokens, or -1 when the caption must go to the Python path
  int encode(const char* p, size_t n, int max_len, int64_t* ids_out, int64_t* mask_out) const {
    static const char* kSpecial[6] = {"[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "[DEC]"};
    std::vector<int32_t> ids;
    ids.reserve(64);
    ids.push_back(CLS);
    std::u32string seg;
    size_t i = 0, seg_start = 0;
    auto run_segment = [&](size_t a, size_t b) -> bool {
      if (b <= a) return true;
      seg.clear();
      if (!utf8_decode(p + a, b - a, seg)) return false;
      return segment(seg, ids, max_len);
    };
    while (i < n) {
      if (p[i] == '[') {
        int hit = -1;
        for (int s = 0; s < 6; ++s) {
          size_t L = strlen(kSpecial[s]);
          if (i + L <= n && memcmp(p + i, kSpecial[s], L) == 0) { hit = s; break; }
        }
        if (hit >= 0) {
          if (!run_segment(seg_start, i)) return -1;
          ids.push_back(special_id[hit]);
          i += strlen(kSpecial[hit]);
          seg_start = i;
          continue;
        }
      }
      ++i;
    }
    if (!run_segment(seg_start, n)) return -1;
    if ((int)ids.size() > max_len - 1) ids.resize(max_len - 1);  // truncation keeps [CLS] ... [SEP]
    ids.push_back(SEP);
    int L = (int)ids.size();
    for (int k = 0; k < max_len; ++k) {
      ids_out[k] = k < L ? ids[k] : 0;
      mask_out[k] = k < L ? 1 : 0;   // by length: a literal "[PAD]" in the text is a live token
    }
    return L;
  }
};

extern "C" {

int sprc_tokenizer_create(const char* vocab_utf8, int64_t vocab_bytes, sprc_tokenizer** out) {
  if (!out) { sprc_set_error("sprc_tokenizer_create: null out pointer"); return -22; }
  if ((vocab_utf8 == nullptr) != (vocab_bytes == 0) || vocab_bytes < 0) {
    sprc_set_error("sprc_tokenizer_create: vocab pointer and size must both be given (real vocabulary) or both be null/0 "
                   "(synthetic hashed vocabulary)");
    return -22;
  }
  auto* t = new sprc_tokenizer();
  if (!vocab_utf8) {
    t->synthetic = true;
  } else {
    // one token per line, id = line number (vocab.txt of bert-base-uncased)
    int32_t id = 0;
    const char* p = vocab_utf8; const char* e = vocab_utf8 + vocab_bytes;
    while (p < e) {
      const char* nl = (const char*)memchr(p, '\n', e - p);
      const char* q = nl ? nl : e;
      t->vocab[std::string(p, q - p)] = id++;   // a repeated token keeps its LAST line number, like load_vocab()
      p = nl ? nl + 1 : e;
    }
    static const char* kSpecial[5] = {"[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"};
    for (int s = 0; s < 5; ++s) {
      auto it = t->vocab.find(kSpecial[s]);
      t->special_id[s] = it != t->vocab.end() ? it->second : UNK;
    }
  }
  *out = t;
  return 0;
}

void sprc_tokenizer_destroy(sprc_tokenizer* t) { delete t; }

int sprc_tokenize_host(const sprc_tokenizer* t, const char* texts, const int64_t* offsets, int n, int max_len,
                       int threads, int64_t* ids, int64_t* mask, int32_t* lens, uint8_t* complex_flags) {
  if (!t || !texts || !offsets || !ids || !mask) { sprc_set_error("sprc_tokenize_host: null argument"); return -22; }
  if (n < 0 || max_len < 2) { sprc_set_error("sprc_tokenize_host: n >= 0 and max_len >= 2 required"); return -22; }
  int hw = (int)std::thread::hardware_concurrency();
  if (threads <= 0) threads = std::max(1, std::min(hw > 0 ? hw : 1, 16));
  threads = std::max(1, std::min(threads, (n + 63) / 64));
  auto work = [&](int lo, int hi) {
    for (int i = lo; i < hi; ++i) {
      int L = t->encode(texts + offsets[i], (size_t)(offsets[i + 1] - offsets[i]), max_len, ids + (int64_t)i * max_len,
                        mask + (int64_t)i * max_len);
      if (complex_flags) complex_flags[i] = L < 0;
      if (L < 0) {  // left for the caller's exact path: emit an empty [CLS][SEP] row so the buffers stay well-formed
        for (int k = 0; k < max_len; ++k) { ids[(int64_t)i * max_len + k] = 0; mask[(int64_t)i * max_len + k] = 0; }
      }
      if (lens) lens[i] = L;
    }
  };
  if (threads == 1) {
    work(0, n);
  } else {
    std::vector<std::thread> pool;
    for (int w = 0; w < threads; ++w) pool.emplace_back(work, (int)((int64_t)n * w / threads), (int)((int64_t)n * (w + 1) / threads));
    for (auto& th : pool) th.join();
  }
  return 0;
}

}  // extern "C"
