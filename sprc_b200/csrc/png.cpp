// Host-side PNG decoder behind the C ABI (sprc_png_decode_files): the input half of gallery indexing.
//
// What it replaces: the reference feeds its indexer with a 2-worker torch DataLoader whose workers run
// `PIL.Image.open(path)` -> `convert("RGB")` inside `targetpad_transform` (src/utils.py:54-64, src/data_utils.py:91-105,
// 167-186, 253-270; the datasets are stored as PNG files).  Here a batch of files is decoded by `threads` C++ workers
// (no interpreter lock, no inter-process tensor hand-off) straight into ONE caller-provided buffer - in practice a pinned
// arena that is copied to the GPU as it is and consumed by the integer resize kernels (csrc/preprocess.cu).
//
// The algorithm is the PNG specification's (third-party dependency of the reference: Pillow's PngImagePlugin + zlib):
// chunk walk with CRC check, zlib inflate of the concatenated IDAT stream, the five scanline filters, expansion to RGB.
// Pixel semantics are Pillow's `Image.open(p).convert("RGB")`: alpha channels are dropped (no compositing), palettes are
// looked up (tRNS ignored), 1/2/4-bit gray is scaled by 255/85/17, 1/2/4-bit palette indices are unpacked.  Files this
// decoder does not take (16-bit samples, Adam7 interlacing, not a PNG at all) come back with status 1 and are decoded by
// the caller with Pillow itself; corrupt files come back with status 2 (Pillow decides whether that is an exception,
// which the reference's datasets turn into a dropped image).
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sprc_b200.h"
#include "inflate.h"

namespace sprc { int set_error(int code, const char* fmt, ...); }  // runtime.cu

namespace {

enum : int32_t { PNG_OK = 0, PNG_UNSUPPORTED = 1, PNG_CORRUPT = 2, PNG_UNREADABLE = 3 };

struct PngInfo {
  uint32_t w = 0, h = 0;
  int depth = 0, color = 0, interlace = 0;
  int channels = 0;
  bool supported = false;
};

inline uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

const uint8_t kSig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};

bool read_file(const char* path, std::vector<uint8_t>& buf) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  if (fseek(f, 0, SEEK_END) != 0) { fclose(f); return false; }
  const long n = ftell(f);
  if (n < 0) { fclose(f); return false; }
  rewind(f);
  buf.resize(static_cast<size_t>(n));
  const size_t got = n ? fread(buf.data(), 1, static_cast<size_t>(n), f) : 0;
  fclose(f);
  return got == static_cast<size_t>(n);
}

// IHDR only: size and whether this decoder takes the file.
int32_t parse_header(const std::vector<uint8_t>& d, PngInfo& o) {
  if (d.size() < 33 || memcmp(d.data(), kSig, 8) != 0) return PNG_UNSUPPORTED;   // not a PNG: the caller's Pillow path
  if (be32(&d[8]) != 13 || memcmp(&d[12], "IHDR", 4) != 0) return PNG_CORRUPT;
  o.w = be32(&d[16]);
  o.h = be32(&d[20]);
  o.depth = d[24];
  o.color = d[25];
  o.interlace = d[28];
  if (o.w == 0 || o.h == 0 || o.w > (1u << 15) || o.h > (1u << 15) || d[26] != 0 || d[27] != 0) return PNG_CORRUPT;
  switch (o.color) {
    case 0: o.channels = 1; o.supported = o.depth == 1 || o.depth == 2 || o.depth == 4 || o.depth == 8; break;
    case 2: o.channels = 3; o.supported = o.depth == 8; break;
    case 3: o.channels = 1; o.supported = o.depth == 1 || o.depth == 2 || o.depth == 4 || o.depth == 8; break;
    case 4: o.channels = 2; o.supported = o.depth == 8; break;
    case 6: o.channels = 4; o.supported = o.depth == 8; break;
    default: return PNG_CORRUPT;
  }
  if (o.interlace != 0) o.supported = false;
  return o.supported ? PNG_OK : PNG_UNSUPPORTED;
}

// The mode Pillow opens the file in: 0 "RGB", 1 "L", 2 "1", 3 "P", 4 "LA", 5 "RGBA".  Matters to callers that resize:
// Pillow resamples "P" / "1" with NEAREST and "LA" / "RGBA" through premultiplied alpha, so resize-then-convert (the
// reference's transform order) equals convert-then-resize only for "RGB" and "L".
inline int32_t pil_mode(const PngInfo& o) {
  switch (o.color) {
    case 2: return 0;
    case 0: return o.depth == 1 ? 2 : 1;
    case 3: return 3;
    case 4: return 4;
    default: return 5;
  }
}

inline uint8_t paeth(int a, int b, int c) {
  const int p = a + b - c;
  const int pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
  return static_cast<uint8_t>((pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c));
}

// Reverses the scanline filter in place (`cur` holds the filtered bytes, `prev` the reconstructed previous line or
// nullptr for the first one).
bool unfilter(int type, uint8_t* cur, const uint8_t* prev, size_t n, size_t bpp) {
  switch (type) {
    case 0: return true;
    case 1:
      for (size_t i = bpp; i < n; ++i) cur[i] = static_cast<uint8_t>(cur[i] + cur[i - bpp]);
      return true;
    case 2:
      if (prev) for (size_t i = 0; i < n; ++i) cur[i] = static_cast<uint8_t>(cur[i] + prev[i]);
      return true;
    case 3:
      for (size_t i = 0; i < n; ++i) {
        const int a = i >= bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0;
        cur[i] = static_cast<uint8_t>(cur[i] + ((a + b) >> 1));
      }
      return true;
    case 4:
      for (size_t i = 0; i < n; ++i) {
        const int a = i >= bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0, c = (prev && i >= bpp) ? prev[i - bpp] : 0;
        cur[i] = static_cast<uint8_t>(cur[i] + paeth(a, b, c));
      }
      return true;
    default: return false;
  }
}

// One reconstructed scanline -> w RGB pixels.
void expand_row(const PngInfo& o, const uint8_t* s, const uint8_t* pal, uint8_t* dst) {
  const uint32_t w = o.w;
  if (o.color == 2) {
    memcpy(dst, s, size_t(w) * 3);
  } else if (o.color == 6) {
    for (uint32_t x = 0; x < w; ++x) { dst[3 * x] = s[4 * x]; dst[3 * x + 1] = s[4 * x + 1]; dst[3 * x + 2] = s[4 * x + 2]; }
  } else if (o.color == 4) {
    for (uint32_t x = 0; x < w; ++x) dst[3 * x] = dst[3 * x + 1] = dst[3 * x + 2] = s[2 * x];
  } else {
    const int d = o.depth;
    const uint32_t mask = (1u << d) - 1;
    // gray: Pillow's "1" -> 0 / 255, "L;2" -> v * 85, "L;4" -> v * 17, "L" as is
    const uint32_t scale = d == 1 ? 255u : d == 2 ? 85u : d == 4 ? 17u : 1u;
    for (uint32_t x = 0; x < w; ++x) {
      uint32_t v;
      if (d == 8) {
        v = s[x];
      } else {
        const uint32_t bit = x * d;
        v = (s[bit >> 3] >> (8 - d - (bit & 7))) & mask;
      }
      if (o.color == 0) {
        const uint8_t g = static_cast<uint8_t>(v * scale);
        dst[3 * x] = dst[3 * x + 1] = dst[3 * x + 2] = g;
      } else {
        dst[3 * x] = pal[3 * v];
        dst[3 * x + 1] = pal[3 * v + 1];
        dst[3 * x + 2] = pal[3 * v + 2];
      }
    }
  }
}

int32_t decode(const std::vector<uint8_t>& d, const PngInfo& o, uint8_t* out, std::vector<uint8_t>& idat,
               std::vector<uint8_t>& raw) {
  uint8_t pal[768];
  memset(pal, 0, sizeof(pal));   // Pillow pads a short palette with zeros
  bool have_pal = false, have_end = false;
  idat.clear();
  size_t pos = 8;
  while (pos + 12 <= d.size()) {
    const uint32_t len = be32(&d[pos]);
    if (len > d.size() - pos - 12) return PNG_CORRUPT;
    const uint8_t* type = &d[pos + 4];
    const uint8_t* body = &d[pos + 8];
    const uint32_t crc = be32(&d[pos + 8 + len]);
    if (static_cast<uint32_t>(crc32(crc32(0L, type, 4), body, len)) != crc) return PNG_CORRUPT;
    if (memcmp(type, "PLTE", 4) == 0) {
      if (len % 3 != 0 || len > 768) return PNG_CORRUPT;
      memcpy(pal, body, len);
      have_pal = true;
    } else if (memcmp(type, "IDAT", 4) == 0) {
      idat.insert(idat.end(), body, body + len);
    } else if (memcmp(type, "IEND", 4) == 0) {
      have_end = true;
      break;
    } else if (!(type[0] & 0x20) && memcmp(type, "IHDR", 4) != 0) {
      return PNG_UNSUPPORTED;   // an unknown CRITICAL chunk: leave the file to Pillow
    }
    pos += 12 + size_t(len);
  }
  if (!have_end || idat.empty() || (o.color == 3 && !have_pal)) return PNG_CORRUPT;
  const size_t bits = size_t(o.depth) * o.channels;
  const size_t stride = (size_t(o.w) * bits + 7) / 8;
  const size_t bpp = std::max<size_t>(1, bits / 8);
  raw.resize((stride + 1) * o.h);
  // own DEFLATE decoder first (inflate.h: ~2x zlib on photo-like PNGs); whatever it refuses goes through zlib, so the
  // set of accepted streams is zlib's
  if (sprc_inflate::inflate_zlib_exact(idat.data(), idat.size(), raw.data(), raw.size()) != 0) {
    uLongf got = static_cast<uLongf>(raw.size());
    const int zr = uncompress(raw.data(), &got, idat.data(), static_cast<uLong>(idat.size()));
    // Z_BUF_ERROR with a full output buffer = trailing data after the last scanline, which Pillow tolerates
    if (!(zr == Z_OK || (zr == Z_BUF_ERROR && got == raw.size())) || got != raw.size()) return PNG_CORRUPT;
  }
  const uint8_t* prev = nullptr;
  for (uint32_t y = 0; y < o.h; ++y) {
    uint8_t* line = raw.data() + size_t(y) * (stride + 1);
    if (!unfilter(line[0], line + 1, prev, stride, bpp)) return PNG_CORRUPT;
    expand_row(o, line + 1, pal, out + size_t(y) * o.w * 3);
    prev = line + 1;
  }
  return PNG_OK;
}

template <class F>
void parallel_for(int n, int threads, F&& fn) {
  if (threads <= 1 || n <= 1) {
    for (int i = 0; i < n; ++i) fn(i);
    return;
  }
  std::atomic<int> next{0};
  std::vector<std::thread> pool;
  pool.reserve(threads);
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&] {
      for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
    });
  for (auto& th : pool) th.join();
}

}  // namespace

extern "C" int sprc_png_decode_files(const char* paths, const int64_t* path_offsets, int n, int threads, uint8_t* out,
                                     int64_t out_capacity, int64_t* pixel_offsets, int32_t* wh, int32_t* status) {
  if (!paths || !path_offsets || !pixel_offsets || !wh || !status || n < 0)
    return sprc::set_error(-22, "sprc_png_decode_files: null argument");
  if (threads <= 0) {
    const int hw = static_cast<int>(std::thread::hardware_concurrency());
    threads = std::max(1, std::min(hw > 0 ? hw : 1, 32));
  }
  threads = std::max(1, std::min(threads, n));
  std::vector<std::vector<uint8_t>> files(static_cast<size_t>(n));
  std::vector<PngInfo> info(static_cast<size_t>(n));
  // pass 1: read + IHDR (sizes are needed before the images can be placed in the arena)
  parallel_for(n, threads, [&](int i) {
    const std::string path(paths + path_offsets[i], paths + path_offsets[i + 1]);
    if (!read_file(path.c_str(), files[i])) {
      status[i] = PNG_UNREADABLE;
      return;
    }
    status[i] = parse_header(files[i], info[i]);
  });
  int64_t off = 0;
  for (int i = 0; i < n; ++i) {
    pixel_offsets[i] = off;
    if (status[i] == PNG_OK) {
      wh[3 * i] = static_cast<int32_t>(info[i].w);
      wh[3 * i + 1] = static_cast<int32_t>(info[i].h);
      wh[3 * i + 2] = pil_mode(info[i]);
      off += int64_t(info[i].w) * info[i].h * 3;
    } else {
      wh[3 * i] = wh[3 * i + 1] = 0;
      wh[3 * i + 2] = -1;
    }
  }
  pixel_offsets[n] = off;
  if (off > out_capacity || (off > 0 && !out))
    return sprc::set_error(-34, "sprc_png_decode_files: the batch needs %lld bytes, the buffer holds %lld",
                           static_cast<long long>(off), static_cast<long long>(out_capacity));
  // pass 2: inflate + unfilter + expand, every image into its own slice
  parallel_for(n, threads, [&](int i) {
    if (status[i] != PNG_OK) return;
    thread_local std::vector<uint8_t> idat, raw;
    status[i] = decode(files[i], info[i], out + pixel_offsets[i], idat, raw);
    std::vector<uint8_t>().swap(files[i]);
  });
  return 0;
}

// Test / micro-benchmark entry: the DEFLATE decoder of inflate.h alone (no zlib fallback).  Returns 0 when `in` is a
// zlib stream that inflates to exactly out_bytes bytes with a matching Adler-32.
extern "C" int sprc_op_inflate_zlib(const uint8_t* in, int64_t in_bytes, uint8_t* out, int64_t out_bytes) {
  if (!in || !out || in_bytes < 0 || out_bytes < 0) return sprc::set_error(-22, "sprc_op_inflate_zlib: bad argument");
  return sprc_inflate::inflate_zlib_exact(in, static_cast<size_t>(in_bytes), out, static_cast<size_t>(out_bytes));
}
