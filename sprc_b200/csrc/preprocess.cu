// GPU image preprocessing for gallery indexing: the reference's `targetpad_transform`
// (src/data_utils.py:52-72 TargetPad, :91-105 Compose[TargetPad, Resize(dim, BICUBIC), CenterCrop(dim), RGB, ToTensor,
// Normalize]) applied to decoded RGB uint8 images, bit-exact with the PIL/torchvision pipeline it replaces.
//
// Third-party algorithm restated here: Pillow's 8-bit resampling (src/libImaging/Resample.c, Pillow 12.2 as installed:
// precompute_coeffs / normalize_coeffs_8bpc / ImagingResampleHorizontal_8bpc / ...Vertical_8bpc): separable, horizontal
// pass first into a uint8 intermediate, fixed-point coefficients with 22 fractional bits, accumulator seeded with
// 1 << 21, result clip8(acc >> 22).  The coefficient tables are computed on the host in double precision exactly as
// Pillow does (sprc_b200/preprocess.py); the kernels below only do the integer passes, so results are bit-identical.
// Zero padding (TargetPad) is virtual: padded pixels contribute 0 to the sums and are never materialised; only the
// columns / rows that survive the centre crop are computed.
#include "common.h"
#include "ops.h"

namespace sprc {

// per-image descriptor (int64 fields; filled by sprc_b200/preprocess.py)
enum {
  PD_SRC_OFF = 0,  // byte offset of the image's RGB pixels (H x W x 3, row-major) in the packed pixel buffer
  PD_W,            // source width
  PD_H,            // source height
  PD_HP,           // left padding (TargetPad)
  PD_VP,           // top padding
  PD_ROW0,         // first padded-image row the vertical pass reads
  PD_NROWS,        // number of such rows (rows of the intermediate)
  PD_KH,           // taps per output column (table row pitch)
  PD_KV,           // taps per output row
  PD_OFF_HB,       // offsets (in int32 elements) into the coefficient buffer: horizontal bounds [dim][2] (xmin, n)
  PD_OFF_HK,       //   horizontal coefficients [dim][KH]
  PD_OFF_VB,       //   vertical bounds [dim][2] (ymin relative to ROW0, n)
  PD_OFF_VK,       //   vertical coefficients [dim][KV]
  PD_TMP_OFF,      // byte offset of the image's intermediate [NROWS][dim][3] in the workspace
  PD_FIELDS = 16
};

static constexpr int PRECISION_BITS = 32 - 8 - 2;   // Resample.c

__device__ __forceinline__ int clip8(int v) {
  v >>= PRECISION_BITS;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: one thread per (intermediate row, output column); 3 channels
__global__ void __launch_bounds__(256)
preprocess_horizontal_kernel(const uint8_t* __restrict__ pixels, const long long* __restrict__ desc,
                             const int* __restrict__ tables, int dim, uint8_t* __restrict__ tmp) {
  const long long* d = desc + static_cast<size_t>(blockIdx.y) * PD_FIELDS;
  const int nrows = static_cast<int>(d[PD_NROWS]);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nrows * dim) return;
  const int r = idx / dim, x = idx % dim;
  const int W = static_cast<int>(d[PD_W]), H = static_cast<int>(d[PD_H]);
  const int sy = static_cast<int>(d[PD_ROW0]) + r - static_cast<int>(d[PD_VP]);   // source row of this padded row
  uint8_t* out = tmp + d[PD_TMP_OFF] + (static_cast<size_t>(r) * dim + x) * 3;
  int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
  if (sy >= 0 && sy < H) {
    const int* hb = tables + d[PD_OFF_HB] + 2 * x;
    const int* hk = tables + d[PD_OFF_HK] + static_cast<size_t>(x) * d[PD_KH];
    const int xmin = hb[0] - static_cast<int>(d[PD_HP]), n = hb[1];
    const uint8_t* row = pixels + d[PD_SRC_OFF] + static_cast<size_t>(sy) * W * 3;
    for (int j = 0; j < n; ++j) {
      const int sx = xmin + j;
      if (sx >= 0 && sx < W) {
        const int k = hk[j];
        a0 += row[sx * 3] * k;
        a1 += row[sx * 3 + 1] * k;
        a2 += row[sx * 3 + 2] * k;
      }
    }
  }
  out[0] = static_cast<uint8_t>(clip8(a0));
  out[1] = static_cast<uint8_t>(clip8(a1));
  out[2] = static_cast<uint8_t>(clip8(a2));
}

// vertical pass + ToTensor + Normalize: one thread per output pixel, writes the three planes
__global__ void __launch_bounds__(256)
preprocess_vertical_kernel(const long long* __restrict__ desc, const int* __restrict__ tables, int dim,
                           const uint8_t* __restrict__ tmp, float3 mean, float3 stdv, float* __restrict__ out) {
  const long long* d = desc + static_cast<size_t>(blockIdx.y) * PD_FIELDS;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= dim * dim) return;
  const int y = idx / dim, x = idx % dim;
  const int* vb = tables + d[PD_OFF_VB] + 2 * y;
  const int* vk = tables + d[PD_OFF_VK] + static_cast<size_t>(y) * d[PD_KV];
  const int ymin = vb[0], n = vb[1];
  const uint8_t* col = tmp + d[PD_TMP_OFF] + (static_cast<size_t>(ymin) * dim + x) * 3;
  int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
  for (int j = 0; j < n; ++j) {
    const int k = vk[j];
    const uint8_t* px = col + static_cast<size_t>(j) * dim * 3;
    a0 += px[0] * k;
    a1 += px[1] * k;
    a2 += px[2] * k;
  }
  // ToTensor: uint8 -> float32 / 255 ; Normalize: (v - mean) / std, both IEEE fp32 like torch
  const float v0 = __fdiv_rn(static_cast<float>(clip8(a0)), 255.f);
  const float v1 = __fdiv_rn(static_cast<float>(clip8(a1)), 255.f);
  const float v2 = __fdiv_rn(static_cast<float>(clip8(a2)), 255.f);
  float* o = out + static_cast<size_t>(blockIdx.y) * 3 * dim * dim + idx;
  o[0] = __fdiv_rn(__fsub_rn(v0, mean.x), stdv.x);
  o[static_cast<size_t>(dim) * dim] = __fdiv_rn(__fsub_rn(v1, mean.y), stdv.y);
  o[2 * static_cast<size_t>(dim) * dim] = __fdiv_rn(__fsub_rn(v2, mean.z), stdv.z);
}

int preprocess_targetpad(const uint8_t* pixels, const long long* desc, const int* tables, int n, int dim, int max_rows,
                         uint8_t* tmp, const float* mean, const float* stdv, float* out, cudaStream_t st) {
  SPRC_REQUIRE(n > 0 && dim > 0 && max_rows > 0, "preprocess: empty batch (n=%d dim=%d rows=%d)", n, dim, max_rows);
  SPRC_REQUIRE(pixels && desc && tables && tmp && out && mean && stdv, "preprocess: null argument");
  dim3 gh((static_cast<unsigned>(max_rows) * dim + 255) / 256, n);
  preprocess_horizontal_kernel<<<gh, 256, 0, st>>>(pixels, desc, tables, dim, tmp);
  count_launch();
  dim3 gv((static_cast<unsigned>(dim) * dim + 255) / 256, n);
  preprocess_vertical_kernel<<<gv, 256, 0, st>>>(desc, tables, dim, tmp, make_float3(mean[0], mean[1], mean[2]),
                                                 make_float3(stdv[0], stdv[1], stdv[2]), out);
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sprc
