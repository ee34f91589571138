// DEFLATE (RFC 1951) / zlib (RFC 1950) decoder for the PNG feed (csrc/png.cpp), host code.
//
// Why not zlib's inflate(): the feed is inflate-bound (profiles/r02o_bench.log: 70 % of a 640x480 PNG's decode time),
// and zlib decodes one symbol per loop trip through a 9/6-bit two-level table with a byte-wise bit buffer.  This decoder
// is built for whole-buffer input and output, which is all a PNG needs: a 64-bit bit buffer refilled without branches,
// an 11-bit primary table for literal/length codes (a literal costs one lookup + one store, up to three per refill),
// 8-byte-wide match copies, and one bounds check per loop trip (the fast loop runs while >= 16 input bytes and >= 320
// output bytes remain; a careful tail loop finishes the stream).  Anything irregular - reserved block type, over- or
// under-subscribed code (except a single distance code), distance beyond the output written so far, stream longer or
// shorter than the caller's buffer - is reported as an error; the caller then hands the file to Pillow, whose zlib is
// the authority on malformed streams.  Output is checked against the stream's Adler-32 (zlib's adler32()).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <zlib.h>

namespace sprc_inflate {

struct Entry {
  uint16_t base;   // literal byte, length / distance base, or first index of a subtable
  uint8_t bits;    // code bits to consume (subtable pointer: the primary bits)
  uint8_t op;      // OP_* | extra-bit count (length / distance) or subtable index bits
};
enum : uint8_t { OP_LITERAL = 0x10, OP_EOB = 0x20, OP_SUB = 0x40, OP_INVALID = 0x80, OP_EXTRA_MASK = 0x0F };

constexpr int LL_BITS = 11, D_BITS = 8, PRE_BITS = 7;
constexpr int LL_TABLE = (1 << LL_BITS) + 1024, D_TABLE = (1 << D_BITS) + 512;   // primary + room for subtables

static const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
                                      67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
                                       1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11,
                                       12, 12, 13, 13};

inline uint32_t reverse_bits(uint32_t v, int n) {
  uint32_t r = 0;
  for (int i = 0; i < n; ++i) r |= ((v >> i) & 1u) << (n - 1 - i);
  return r;
}

// kind: 0 = literal/length alphabet, 1 = distance alphabet, 2 = code-length alphabet (symbol value in `base`).
// Returns false for an over-subscribed or incomplete code (one exception, as zlib: a single distance code of one bit).
inline bool build_table(const uint8_t* lens, int nsym, int kind, int primary_bits, Entry* table, int table_cap) {
  int count[16] = {0};
  for (int i = 0; i < nsym; ++i) count[lens[i]]++;
  count[0] = 0;
  int maxlen = 15;
  while (maxlen > 0 && count[maxlen] == 0) --maxlen;
  const Entry invalid = {0, 1, OP_INVALID};
  const int psize = 1 << primary_bits;
  for (int i = 0; i < psize; ++i) table[i] = invalid;
  if (maxlen == 0) return kind == 1;   // no distance codes at all: legal as long as no match occurs (every slot invalid)
  long left = 1;
  for (int l = 1; l <= 15; ++l) {
    left = (left << 1) - count[l];
    if (left < 0) return false;   // over-subscribed
  }
  if (left > 0 && !(kind == 1 && maxlen == 1 && count[1] == 1)) return false;   // incomplete
  uint32_t next_code[16];
  uint32_t code = 0;
  for (int l = 1; l <= 15; ++l) {
    code = (code + count[l - 1]) << 1;
    next_code[l] = code;
  }
  auto make = [&](int sym, int len) {
    Entry e;
    e.bits = static_cast<uint8_t>(len);
    if (kind == 2) {
      e.base = static_cast<uint16_t>(sym);
      e.op = OP_LITERAL;
    } else if (kind == 1) {
      if (sym >= 30) return Entry{0, static_cast<uint8_t>(len), OP_INVALID};
      e.base = kDistBase[sym];
      e.op = kDistExtra[sym];
    } else if (sym < 256) {
      e.base = static_cast<uint16_t>(sym);
      e.op = OP_LITERAL | 1;   // low bits: number of literals this entry emits (pairs are formed below)
    } else if (sym == 256) {
      e.base = 0;
      e.op = OP_EOB;
    } else if (sym <= 285) {
      e.base = kLenBase[sym - 257];
      e.op = kLenExtra[sym - 257];
    } else {
      return Entry{0, static_cast<uint8_t>(len), OP_INVALID};
    }
    return e;
  };
  // codes of at most primary_bits: replicated over the primary table
  // longer codes: one subtable per primary prefix, sized by the longest code under that prefix
  int sub_next = psize;
  // pass 1: longest code per prefix
  static thread_local uint8_t prefix_max[1 << LL_BITS];
  if (maxlen > primary_bits) memset(prefix_max, 0, static_cast<size_t>(psize));
  uint32_t nc[16];
  memcpy(nc, next_code, sizeof(nc));
  if (maxlen > primary_bits) {
    for (int s = 0; s < nsym; ++s) {
      const int l = lens[s];
      if (l == 0) continue;
      const uint32_t c = nc[l]++;
      if (l > primary_bits) {
        const uint32_t rev = reverse_bits(c, l);
        const uint32_t pre = rev & (psize - 1);
        if (l > prefix_max[pre]) prefix_max[pre] = static_cast<uint8_t>(l);
      }
    }
  }
  for (int s = 0; s < nsym; ++s) {
    const int l = lens[s];
    if (l == 0) continue;
    const uint32_t c = next_code[l]++;
    const uint32_t rev = reverse_bits(c, l);
    if (l <= primary_bits) {
      const Entry e = make(s, l);
      for (uint32_t i = rev; i < static_cast<uint32_t>(psize); i += 1u << l) table[i] = e;
    } else {
      const uint32_t pre = rev & (psize - 1);
      const int sub_bits = prefix_max[pre] - primary_bits;
      if (!(table[pre].op & OP_SUB)) {
        if (sub_next + (1 << sub_bits) > table_cap) return false;
        table[pre] = Entry{static_cast<uint16_t>(sub_next), static_cast<uint8_t>(primary_bits),
                           static_cast<uint8_t>(OP_SUB | sub_bits)};
        for (int i = 0; i < (1 << sub_bits); ++i) table[sub_next + i] = invalid;
        sub_next += 1 << sub_bits;
      }
      Entry e = make(s, l);
      e.bits = static_cast<uint8_t>(l - primary_bits);
      const uint32_t start = table[pre].base;
      for (uint32_t i = rev >> primary_bits; i < (1u << sub_bits); i += 1u << (l - primary_bits)) table[start + i] = e;
    }
  }
  if (kind == 0) {
    // Two literals per lookup where both codes fit in the primary index: entry i = (first literal, second literal) when
    // the bits left after the first code already determine a second literal code.  Photo-like PNG streams are mostly
    // literals with 4..7-bit codes, and the decode loop is a serial chain (lookup -> shift -> lookup), so every pair
    // saves one trip of that chain.
    static thread_local Entry orig[1 << LL_BITS];
    memcpy(orig, table, sizeof(Entry) * static_cast<size_t>(psize));
    for (int i = 0; i < psize; ++i) {
      const Entry e1 = orig[i];
      if (e1.op != (OP_LITERAL | 1) || e1.bits >= primary_bits) continue;
      const Entry e2 = orig[i >> e1.bits];
      if (e2.op != (OP_LITERAL | 1) || e1.bits + e2.bits > primary_bits) continue;
      table[i] = Entry{static_cast<uint16_t>(e1.base | (e2.base << 8)), static_cast<uint8_t>(e1.bits + e2.bits),
                       static_cast<uint8_t>(OP_LITERAL | 2)};
    }
  }
  return true;
}

struct Tables {
  Entry ll[LL_TABLE];
  Entry d[D_TABLE];
};

inline uint64_t load64(const uint8_t* p) {
  uint64_t v;
  memcpy(&v, p, 8);
  return v;   // little-endian hosts (x86-64, aarch64)
}

// Raw DEFLATE stream in[0, n) -> out[0, cap).  Returns 0 and *produced on success, nonzero on any irregularity.
inline int inflate_raw(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* produced, size_t* consumed) {
  static thread_local Tables T;
  static thread_local Tables Fixed;
  static thread_local bool fixed_ready = false;
  const uint8_t* ip = in;
  const uint8_t* const in_end = in + n;
  uint8_t* op = out;
  uint8_t* const out_end = out + cap;
  uint64_t bitbuf = 0;
  int bitcnt = 0;

  // bytes past the end of the input read as zeros (the stream is then caught by the "overrun" check)
  auto refill_slow = [&]() {
    while (bitcnt <= 56) {
      if (ip < in_end) bitbuf |= static_cast<uint64_t>(*ip) << bitcnt;
      ++ip;
      bitcnt += 8;
    }
  };
  auto overrun = [&]() { return ip > in_end && static_cast<size_t>(ip - in_end) * 8 > static_cast<size_t>(bitcnt); };
  auto take = [&](int k) {
    const uint32_t v = static_cast<uint32_t>(bitbuf & ((1ull << k) - 1));
    bitbuf >>= k;
    bitcnt -= k;
    return v;
  };

  for (;;) {
    refill_slow();
    const uint32_t final_block = take(1);
    const uint32_t type = take(2);
    const Tables* tb = nullptr;
    if (type == 0) {
      // stored: skip to the byte boundary, LEN / NLEN, raw bytes
      take(bitcnt & 7);
      refill_slow();
      const uint32_t len = take(16), nlen = take(16);
      if ((len ^ 0xFFFFu) != nlen) return 1;
      // give whole bytes of the bit buffer back to the input pointer
      ip -= bitcnt >> 3;
      bitbuf = 0;
      bitcnt = 0;
      if (ip > in_end || static_cast<size_t>(in_end - ip) < len || static_cast<size_t>(out_end - op) < len) return 2;
      memcpy(op, ip, len);
      op += len;
      ip += len;
    } else if (type == 1) {
      if (!fixed_ready) {
        uint8_t l[288 + 32];
        for (int i = 0; i < 144; ++i) l[i] = 8;
        for (int i = 144; i < 256; ++i) l[i] = 9;
        for (int i = 256; i < 280; ++i) l[i] = 7;
        for (int i = 280; i < 288; ++i) l[i] = 8;
        for (int i = 0; i < 32; ++i) l[288 + i] = 5;
        if (!build_table(l, 288, 0, LL_BITS, Fixed.ll, LL_TABLE)) return 3;
        if (!build_table(l + 288, 32, 1, D_BITS, Fixed.d, D_TABLE)) return 3;
        fixed_ready = true;
      }
      tb = &Fixed;
    } else if (type == 2) {
      const int hlit = static_cast<int>(take(5)) + 257, hdist = static_cast<int>(take(5)) + 1;
      const int hclen = static_cast<int>(take(4)) + 4;
      if (hlit > 286 || hdist > 30) return 4;
      static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
      uint8_t pre_lens[19] = {0};
      for (int i = 0; i < hclen; ++i) {
        refill_slow();
        pre_lens[order[i]] = static_cast<uint8_t>(take(3));
      }
      Entry pre[1 << PRE_BITS];
      if (!build_table(pre_lens, 19, 2, PRE_BITS, pre, 1 << PRE_BITS)) return 5;
      uint8_t lens[286 + 30 + 140];
      int i = 0;
      while (i < hlit + hdist) {
        refill_slow();
        const Entry e = pre[bitbuf & ((1u << PRE_BITS) - 1)];
        if (e.op & OP_INVALID) return 6;
        take(e.bits);
        const int sym = e.base;
        if (sym < 16) {
          lens[i++] = static_cast<uint8_t>(sym);
        } else {
          int rep;
          uint8_t v = 0;
          if (sym == 16) {
            if (i == 0) return 7;
            v = lens[i - 1];
            rep = 3 + static_cast<int>(take(2));
          } else if (sym == 17) {
            rep = 3 + static_cast<int>(take(3));
          } else {
            rep = 11 + static_cast<int>(take(7));
          }
          if (i + rep > hlit + hdist) return 8;
          memset(lens + i, v, static_cast<size_t>(rep));
          i += rep;
        }
        if (overrun()) return 9;
      }
      if (lens[256] == 0) return 10;   // no end-of-block code
      if (!build_table(lens, hlit, 0, LL_BITS, T.ll, LL_TABLE)) return 11;
      if (!build_table(lens + hlit, hdist, 1, D_BITS, T.d, D_TABLE)) return 12;
      tb = &T;
    } else {
      return 13;
    }

    if (tb) {
      const Entry* const ll = tb->ll;
      const Entry* const dt = tb->d;
      bool done = false;
      // ---------------- fast loop: no per-symbol bounds checks ----------------
      while (!done && in_end - ip >= 16 && ip <= in_end && out_end - op >= 320) {
        // branch-free refill to >= 56 bits
        bitbuf |= load64(ip) << bitcnt;
        ip += (63 - bitcnt) >> 3;
        bitcnt |= 56;
        Entry e = ll[bitbuf & ((1u << LL_BITS) - 1)];
        if (e.op & OP_LITERAL) {   // up to three lookups per refill (3 x 11 bits < 56), one or two literals each
          bitbuf >>= e.bits;
          bitcnt -= e.bits;
          memcpy(op, &e.base, 2);   // the second byte is only kept when the entry holds a pair
          op += e.op & 3;
          e = ll[bitbuf & ((1u << LL_BITS) - 1)];
          if (e.op & OP_LITERAL) {
            bitbuf >>= e.bits;
            bitcnt -= e.bits;
            memcpy(op, &e.base, 2);
            op += e.op & 3;
            e = ll[bitbuf & ((1u << LL_BITS) - 1)];
            if (e.op & OP_LITERAL) {
              bitbuf >>= e.bits;
              bitcnt -= e.bits;
              memcpy(op, &e.base, 2);
              op += e.op & 3;
              continue;
            }
          }
          // fewer than 56 - 30 = 26 bits may be left: refill before a length + distance (up to 48 bits)
          bitbuf |= load64(ip) << bitcnt;
          ip += (63 - bitcnt) >> 3;
          bitcnt |= 56;
        }
        if (e.op & OP_SUB) {
          bitbuf >>= e.bits;
          bitcnt -= e.bits;
          e = ll[e.base + (bitbuf & ((1u << (e.op & OP_EXTRA_MASK)) - 1))];
          if (e.op & OP_LITERAL) {   // subtable entries are single literals
            bitbuf >>= e.bits;
            bitcnt -= e.bits;
            *op++ = static_cast<uint8_t>(e.base);
            continue;
          }
        }
        if (e.op & (OP_EOB | OP_INVALID)) {
          if (e.op & OP_INVALID) return 14;
          bitbuf >>= e.bits;
          bitcnt -= e.bits;
          done = true;
          break;
        }
        // length
        bitbuf >>= e.bits;
        bitcnt -= e.bits;
        const int lx = e.op & OP_EXTRA_MASK;
        uint32_t len = e.base + static_cast<uint32_t>(bitbuf & ((1u << lx) - 1));
        bitbuf >>= lx;
        bitcnt -= lx;
        // distance (15 + 13 bits at most; at least 56 - 15 - 15 - 5 = 21 bits are left: refill first)
        if (bitcnt < 32) {
          bitbuf |= load64(ip) << bitcnt;
          ip += (63 - bitcnt) >> 3;
          bitcnt |= 56;
        }
        Entry de = dt[bitbuf & ((1u << D_BITS) - 1)];
        if (de.op & OP_SUB) {
          bitbuf >>= de.bits;
          bitcnt -= de.bits;
          de = dt[de.base + (bitbuf & ((1u << (de.op & OP_EXTRA_MASK)) - 1))];
        }
        if (de.op & OP_INVALID) return 15;
        bitbuf >>= de.bits;
        bitcnt -= de.bits;
        const int dx = de.op & OP_EXTRA_MASK;
        const uint32_t dist = de.base + static_cast<uint32_t>(bitbuf & ((1u << dx) - 1));
        bitbuf >>= dx;
        bitcnt -= dx;
        if (dist > static_cast<size_t>(op - out)) return 16;
        const uint8_t* src = op - dist;
        uint8_t* dst = op;
        op += len;
        if (dist >= 8) {   // 8 bytes at a time; may write up to 7 bytes past the match (room is guaranteed: 320)
          do {
            memcpy(dst, src, 8);
            dst += 8;
            src += 8;
          } while (dst < op);
        } else if (dist == 1) {
          memset(dst, *src, len);
        } else {
          do {
            *dst++ = *src++;
          } while (dst < op);
        }
      }
      // ---------------- careful loop: the last bytes of input / output ----------------
      // (the fast loop may leave ip advanced with whole bytes in the bit buffer; refill_slow continues from there)
      while (!done) {
        refill_slow();
        Entry e = ll[bitbuf & ((1u << LL_BITS) - 1)];
        if (e.op & OP_SUB) {
          take(e.bits);
          e = ll[e.base + (bitbuf & ((1u << (e.op & OP_EXTRA_MASK)) - 1))];
        }
        if (e.op & OP_INVALID) return 17;
        take(e.bits);
        if (e.op & OP_LITERAL) {
          const int nl = e.op & 3;
          if (out_end - op < nl) return 18;
          *op++ = static_cast<uint8_t>(e.base);
          if (nl == 2) *op++ = static_cast<uint8_t>(e.base >> 8);
        } else if (e.op & OP_EOB) {
          done = true;
        } else {
          uint32_t len = e.base + take(e.op & OP_EXTRA_MASK);
          refill_slow();
          Entry de = dt[bitbuf & ((1u << D_BITS) - 1)];
          if (de.op & OP_SUB) {
            take(de.bits);
            de = dt[de.base + (bitbuf & ((1u << (de.op & OP_EXTRA_MASK)) - 1))];
          }
          if (de.op & OP_INVALID) return 19;
          take(de.bits);
          const uint32_t dist = de.base + take(de.op & OP_EXTRA_MASK);
          if (dist > static_cast<size_t>(op - out) || static_cast<size_t>(out_end - op) < len) return 20;
          const uint8_t* src = op - dist;
          for (uint32_t k = 0; k < len; ++k) op[k] = src[k];
          op += len;
        }
        if (overrun()) return 21;
      }
    }
    if (overrun()) return 22;
    if (final_block) break;
  }
  // unread whole bytes go back to the caller (the zlib trailer follows)
  take(bitcnt & 7);
  ip -= bitcnt >> 3;
  if (ip > in_end) return 23;
  *produced = static_cast<size_t>(op - out);
  *consumed = static_cast<size_t>(ip - in);
  return 0;
}

// zlib stream (2-byte header, DEFLATE, Adler-32) whose output must fill out[0, cap) EXACTLY (a PNG knows its size).
// Trailing bytes after the Adler-32 are tolerated, as Pillow does.
inline int inflate_zlib_exact(const uint8_t* in, size_t n, uint8_t* out, size_t cap) {
  if (n < 6) return 30;
  const uint32_t cmf = in[0], flg = in[1];
  if ((cmf & 0x0F) != 8 || (cmf >> 4) > 7 || ((cmf << 8) | flg) % 31 != 0 || (flg & 0x20)) return 31;
  size_t produced = 0, consumed = 0;
  const int rc = inflate_raw(in + 2, n - 2, out, cap, &produced, &consumed);
  if (rc != 0) return rc;
  if (produced != cap) return 32;
  if (n - 2 - consumed < 4) return 33;
  const uint8_t* t = in + 2 + consumed;
  const uint32_t want = (uint32_t(t[0]) << 24) | (uint32_t(t[1]) << 16) | (uint32_t(t[2]) << 8) | t[3];
  uLong a = adler32(0L, Z_NULL, 0);
  size_t off = 0;
  while (off < cap) {   // adler32 takes a uInt length
    const size_t chunk = cap - off < (1u << 30) ? cap - off : (1u << 30);
    a = adler32(a, out + off, static_cast<uInt>(chunk));
    off += chunk;
  }
  return static_cast<uint32_t>(a) == want ? 0 : 34;
}

}  // namespace sprc_inflate
