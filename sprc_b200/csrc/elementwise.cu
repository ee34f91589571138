// HBM-bound elementwise / row-reduction kernels of the hot path: LayerNorm, patch extraction,
// token assembly, Q-Former embeddings, dtype conversion, row gather, L2 normalisation, ITM head.
// All use 128-bit accesses and warp-shuffle reductions; none re-reads its input.
#include <atomic>

#include "ops.h"
#include "ptx.cuh"

namespace sprc {

static std::atomic<int64_t> g_launches{0};
int64_t launch_count() { return g_launches.load(); }
void count_launch(int n) { g_launches.fetch_add(n); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row kept in registers (width <= 32*4*MAXV)
// ------------------------------------------------------------------------------------------------
template <int MAXV>  // float4 vectors per lane
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int rows, int width, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, int grp_rows, int grp_stride, float* out_f32,
                 bf16* out_bf16, int fp16, int rev) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  griddep_wait();
  griddep_launch();
  if (row >= rows) return;
  if (rev) row = rows - 1 - row;  // low block indices (scheduled first) take the last rows
  const int lane = threadIdx.x & 31;
  long long prow = row;
  if (grp_rows > 0) prow = (long long)(row / grp_rows) * grp_stride + (row % grp_rows);
  const int nvec = width >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)prow * width);
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      v[i] = xr[c];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float mean = warp_sum(s) / (float)width;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      ss += (a * a + b * b) + (cc * cc + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / (float)width + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      const float4 g = __ldg(g4 + c), b = __ldg(b4 + c);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      if (out_f32) reinterpret_cast<float4*>(out_f32 + (size_t)prow * width)[c] = o;
      if (out_bf16)
        reinterpret_cast<uint2*>(out_bf16 + (size_t)prow * width)[c] =
            make_uint2(pack_act(o.x, o.y, fp16), pack_act(o.z, o.w, fp16));
    }
  }
}

int layernorm(const float* x, int rows, int width, const float* gamma, const float* beta, float eps, int grp_rows,
              int grp_stride, float* out_f32, bf16* out_bf16, cudaStream_t st) {
  SPRC_REQUIRE(width % 4 == 0 && width <= 32 * 4 * 12, "layernorm: unsupported width %d", width);
  if (rows <= 0) return 0;
  const int wpb = 8;
  const int grid = (rows + wpb - 1) / wpb;
  const int rev = next_sweep_reverse();
  prof_begin(st);
  if (width <= 32 * 4 * 6)
    SPRC_CUDA(launch_pdl(layernorm_kernel<6>, dim3(grid), dim3(wpb * 32), 0, st, x, rows, width, gamma, beta, eps,
                         grp_rows, grp_stride, out_f32, out_bf16, act_fp16(), rev));
  else
    SPRC_CUDA(launch_pdl(layernorm_kernel<12>, dim3(grid), dim3(wpb * 32), 0, st, x, rows, width, gamma, beta, eps,
                         grp_rows, grp_stride, out_f32, out_bf16, act_fp16(), rev));
  prof_end(PROF_ELEMWISE, 0.0, (double)rows * width * (4.0 + (out_f32 ? 4.0 : 0.0) + (out_bf16 ? 2.0 : 0.0)), st);
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// patch extraction: one CTA per (image, patch-row of 16 patches); coalesced reads of 14 image rows
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
im2col_kernel(const float* __restrict__ img, bf16* __restrict__ patches_, int ldp, int fp16) {
  unsigned short* patches = reinterpret_cast<unsigned short*>(patches_);
  const int b = blockIdx.x >> 4;
  const int py = blockIdx.x & 15;
  // element (c, ky, x) with x in [0,224): source img[b][c][py*14+ky][x]; dest patch (py*16 + x/14),
  // column c*196 + ky*14 + x%14
  for (int i = threadIdx.x; i < 3 * 14 * 224; i += blockDim.x) {
    const int x = i % 224;
    const int ky = (i / 224) % 14;
    const int c = i / (224 * 14);
    const float v = img[(((size_t)b * 3 + c) * 224 + (py * 14 + ky)) * 224 + x];
    const int px = x / 14, kx = x % 14;
    patches[((size_t)b * 256 + py * 16 + px) * ldp + c * 196 + ky * 14 + kx] = to_act(v, fp16);
  }
  // zero the K padding columns [588, ldp)
  const int pad = ldp - 588;
  for (int i = threadIdx.x; i < 16 * pad; i += blockDim.x) {
    const int px = i / pad, j = i % pad;
    patches[((size_t)b * 256 + py * 16 + px) * ldp + 588 + j] = 0;
  }
}

int im2col_patches(const float* images, int B, bf16* patches, int ldp, cudaStream_t st) {
  SPRC_REQUIRE(ldp >= 588, "im2col: ldp %d < 588", ldp);
  if (B <= 0) return 0;
  im2col_kernel<<<B * 16, 256, 0, st>>>(images, patches, ldp, act_fp16());
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256)
vit_assemble_kernel(const float4* __restrict__ patch_out, const float4* __restrict__ cls,
                    const float4* __restrict__ pos, int B, int w4, float4* __restrict__ x) {
  const size_t total = (size_t)B * 257 * w4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % w4);
    const size_t r = i / w4;
    const int t = (int)(r % 257);
    const size_t b = r / 257;
    float4 v = t == 0 ? __ldg(cls + c) : patch_out[(b * 256 + (t - 1)) * w4 + c];
    const float4 p = __ldg(pos + (size_t)t * w4 + c);
    v.x += p.x;
    v.y += p.y;
    v.z += p.z;
    v.w += p.w;
    x[i] = v;
  }
}

int vit_assemble_tokens(const float* patch_out, const float* cls, const float* pos, int B, int width, float* x,
                        cudaStream_t st) {
  if (B <= 0) return 0;
  const size_t total = (size_t)B * 257 * (width / 4);
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  vit_assemble_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(patch_out),
                                            reinterpret_cast<const float4*>(cls),
                                            reinterpret_cast<const float4*>(pos), B, width / 4,
                                            reinterpret_cast<float4*>(x));
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Q-Former embeddings (pre-LayerNorm rows), width fixed at 768 = 192 float4
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(192)
qformer_embed_kernel(const float4* __restrict__ qe, int q_batch_rows, const int64_t* __restrict__ ids, int ids_div,
                     const float4* __restrict__ word, const float4* __restrict__ pos, int vocab, int S,
                     float4* __restrict__ out) {
  const int r = blockIdx.x;  // row over B*S
  const int b = r / S, s = r % S;
  const int c = threadIdx.x;
  float4 v;
  if (s < 32) {
    v = q_batch_rows == 0 ? __ldg(qe + (size_t)s * 192 + c) : qe[((size_t)b * q_batch_rows + s) * 192 + c];
  } else {
    long long id = ids[(size_t)(b / ids_div) * 32 + (s - 32)];
    if (id < 0) id = 0;
    if (id >= vocab) id = vocab - 1;
    const float4 w = __ldg(word + (size_t)id * 192 + c);
    const float4 p = __ldg(pos + (size_t)(s - 32) * 192 + c);
    v = make_float4(w.x + p.x, w.y + p.y, w.z + p.z, w.w + p.w);
  }
  out[(size_t)r * 192 + c] = v;
}

int qformer_embed_rows(const float* query_embeds, int q_batch_rows, const int64_t* ids, int ids_div,
                       const float* word_emb, const float* pos_emb, int vocab, int B, float* out, cudaStream_t st) {
  if (ids_div < 1) ids_div = 1;
  if (B <= 0) return 0;
  const int S = ids ? 64 : 32;
  qformer_embed_kernel<<<B * S, 192, 0, st>>>(reinterpret_cast<const float4*>(query_embeds), q_batch_rows, ids, ids_div,
                                              reinterpret_cast<const float4*>(word_emb),
                                              reinterpret_cast<const float4*>(pos_emb), vocab, S,
                                              reinterpret_cast<float4*>(out));
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

// Ragged layout (attention_qfr.cu): rows [0, 32 B) query rows, rows [32 B + toff[b], + L[b]) text rows of sample b.
// One block per query row, then one block per text row (row_sample[slot] = owning sample); slack rows (t >= L[b]) are
// zero.
__global__ void __launch_bounds__(192)
qformer_embed_ragged_kernel(const float4* __restrict__ qe, int q_is_batched, const int64_t* __restrict__ ids,
                            int ids_div, const int* __restrict__ slot_sample, const int* __restrict__ toff,
                            const int* __restrict__ len, const float4* __restrict__ word,
                            const float4* __restrict__ pos, int vocab, int B, float4* __restrict__ out) {
  const int r = blockIdx.x;
  const int c = threadIdx.x;
  float4 v;
  if (r < 32 * B) {
    v = q_is_batched ? qe[(size_t)r * 192 + c] : __ldg(qe + (size_t)(r & 31) * 192 + c);
  } else {
    const int slot = r - 32 * B;            // row inside the text region
    const int b = slot_sample[slot];        // sample owning this row (slack rows: the pair's second sample)
    const int t = slot - toff[b];
    if (t < len[b]) {
      long long id = ids[(size_t)(b / ids_div) * 32 + t];   // rerank repeats each caption for its T candidates
      if (id < 0) id = 0;
      if (id >= vocab) id = vocab - 1;
      const float4 w = __ldg(word + (size_t)id * 192 + c);
      const float4 p = __ldg(pos + (size_t)t * 192 + c);
      v = make_float4(w.x + p.x, w.y + p.y, w.z + p.z, w.w + p.w);
    } else {
      v = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  out[(size_t)r * 192 + c] = v;
}

int qformer_embed_ragged(const float* query_embeds, int q_is_batched, const int64_t* ids, int ids_div,
                         const int* slot_sample, const int* toff, const int* len, const float* word_emb,
                         const float* pos_emb, int vocab, int B, int rows_total, float* out, cudaStream_t st) {
  if (B <= 0) return 0;
  if (ids_div < 1) ids_div = 1;
  qformer_embed_ragged_kernel<<<rows_total, 192, 0, st>>>(
      reinterpret_cast<const float4*>(query_embeds), q_is_batched, ids, ids_div, slot_sample, toff, len,
      reinterpret_cast<const float4*>(word_emb), reinterpret_cast<const float4*>(pos_emb), vocab, B,
      reinterpret_cast<float4*>(out));
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

// dst[i, :] = src[base + rows[i], :] for 768-wide rows: fp32 and (optionally) the 16-bit copy in one launch
__global__ void __launch_bounds__(192)
gather_rows768_kernel(const float4* __restrict__ src32, const uint2* __restrict__ src16, const int* __restrict__ rows,
                      int base, float4* __restrict__ dst32, uint2* __restrict__ dst16) {
  const size_t r = (size_t)(base + rows[blockIdx.x]);
  const int c = threadIdx.x;
  if (src32) dst32[(size_t)blockIdx.x * 192 + c] = src32[r * 192 + c];
  if (src16) dst16[(size_t)blockIdx.x * 192 + c] = src16[r * 192 + c];
}

int gather_rows768(const float* src32, const bf16* src16, const int* rows, int base, int n, float* dst32, bf16* dst16,
                   cudaStream_t st) {
  if (n <= 0) return 0;
  gather_rows768_kernel<<<n, 192, 0, st>>>(reinterpret_cast<const float4*>(src32),
                                           reinterpret_cast<const uint2*>(src16), rows, base,
                                           reinterpret_cast<float4*>(dst32), reinterpret_cast<uint2*>(dst16));
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

__global__ void qformer_key_mask_kernel(const int64_t* __restrict__ am, int div, int B, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 64) return;
  const int b = i >> 6, j = i & 63;
  out[i] = j < 32 ? 0.f : (1.0f - (float)am[(size_t)(b / div) * 32 + (j - 32)]) * -10000.0f;
}

int qformer_key_mask(const int64_t* attention_mask, int div, int B, float* out, cudaStream_t st) {
  if (B <= 0) return 0;
  if (div < 1) div = 1;
  qformer_key_mask_kernel<<<(B * 64 + 255) / 256, 256, 0, st>>>(attention_mask, div, B, out);
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// conversions and gathers
// ------------------------------------------------------------------------------------------------
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) convert_kernel(const TIn* __restrict__ in, TOut* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = static_cast<TOut>(static_cast<float>(in[i]));
}

template <typename TIn, typename TOut>
static int launch_convert(const TIn* in, TOut* out, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  convert_kernel<TIn, TOut><<<(int)blocks, 256, 0, st>>>(in, out, n);
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

int convert_f32_to_bf16(const float* in, bf16* out, size_t n, cudaStream_t st) {
  if (act_fp16()) return launch_convert<float, __half>(in, reinterpret_cast<__half*>(out), n, st);
  return launch_convert<float, bf16>(in, out, n, st);
}
int convert_f16_to_bf16(const void* in, bf16* out, size_t n, cudaStream_t st) {
  if (act_fp16()) return launch_convert<__half, __half>(static_cast<const __half*>(in), reinterpret_cast<__half*>(out), n, st);
  return launch_convert<__half, bf16>(static_cast<const __half*>(in), out, n, st);
}
int convert_f16_to_f32(const void* in, float* out, size_t n, cudaStream_t st) {
  return launch_convert<__half, float>(static_cast<const __half*>(in), out, n, st);
}

template <typename TIn>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const TIn* __restrict__ table, const int32_t* __restrict__ rows, size_t row_elems,
                   bf16* __restrict__ out, int fp16) {
  const size_t r = blockIdx.x;
  const TIn* src = table + (size_t)rows[r] * row_elems;
  bf16* dst = out + r * row_elems;
  if (sizeof(TIn) == 2) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (size_t i = threadIdx.x; i < row_elems / 8; i += blockDim.x) d4[i] = s4[i];
  } else {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    uint2* d2 = reinterpret_cast<uint2*>(dst);
    for (size_t i = threadIdx.x; i < row_elems / 4; i += blockDim.x) {
      const float4 v = s4[i];
      d2[i] = make_uint2(pack_act(v.x, v.y, fp16), pack_act(v.z, v.w, fp16));
    }
  }
}

int gather_rows_bf16(const void* table, int table_dtype, const int32_t* rows, int n_rows, size_t row_elems,
                     bf16* out, cudaStream_t st) {
  SPRC_REQUIRE(row_elems % 8 == 0, "gather_rows: row_elems %zu not a multiple of 8", row_elems);
  if (n_rows <= 0) return 0;
  if (table_dtype == 2)
    gather_rows_kernel<bf16><<<n_rows, 256, 0, st>>>(static_cast<const bf16*>(table), rows, row_elems, out, act_fp16());
  else if (table_dtype == 0)
    gather_rows_kernel<float><<<n_rows, 256, 0, st>>>(static_cast<const float*>(table), rows, row_elems, out, act_fp16());
  else
    return set_error(-22, "gather_rows: unsupported table dtype %d", table_dtype);
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// L2 normalisation of 256-wide rows: one warp per row, 8 floats per lane
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
l2norm256_kernel(const float* __restrict__ in, size_t in_row_stride, int rows, float* out_f32, bf16* out_bf16,
                 int fp16) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* src = reinterpret_cast<const float4*>(in + (size_t)row * in_row_stride) + lane * 2;
  const float4 a = src[0], b = src[1];
  float ss = (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w) + (b.x * b.x + b.y * b.y) + (b.z * b.z + b.w * b.w);
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize: x / max(||x||, eps)
  const float o[8] = {a.x * inv, a.y * inv, a.z * inv, a.w * inv, b.x * inv, b.y * inv, b.z * inv, b.w * inv};
  if (out_f32) {
    float4* d = reinterpret_cast<float4*>(out_f32 + (size_t)row * 256) + lane * 2;
    d[0] = make_float4(o[0], o[1], o[2], o[3]);
    d[1] = make_float4(o[4], o[5], o[6], o[7]);
  }
  if (out_bf16) {
    uint4* d = reinterpret_cast<uint4*>(out_bf16 + (size_t)row * 256) + lane;
    *d = make_uint4(pack_act(o[0], o[1], fp16), pack_act(o[2], o[3], fp16), pack_act(o[4], o[5], fp16),
                    pack_act(o[6], o[7], fp16));
  }
}

int l2norm_rows256(const float* in, size_t in_row_stride, int rows, float* out_f32, bf16* out_bf16,
                   cudaStream_t st) {
  if (rows <= 0) return 0;
  l2norm256_kernel<<<(rows + 7) / 8, 256, 0, st>>>(in, in_row_stride, rows, out_f32, out_bf16, act_fp16());
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// ITM head: logits[pair] = mean_{r<32} (W h_r + b), p = softmax(logits)[1]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
itm_head_kernel(const float* __restrict__ h, int rows_per_pair, const float* __restrict__ w,
                const float* __restrict__ b, float* __restrict__ p) {
  // mean over rows commutes with the linear head: first average the 32 rows, then two dot products
  __shared__ float s0[8], s1[8];
  const size_t pair = blockIdx.x;
  const float* base = h + pair * rows_per_pair * 768;
  float a0 = 0.f, a1 = 0.f;
  for (int c = threadIdx.x; c < 768; c += blockDim.x) {
    float m = 0.f;
#pragma unroll 4
    for (int r = 0; r < 32; ++r) m += base[(size_t)r * 768 + c];
    m *= (1.0f / 32.0f);
    a0 += m * w[c];
    a1 += m * w[768 + c];
  }
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  if ((threadIdx.x & 31) == 0) {
    s0[threadIdx.x >> 5] = a0;
    s1[threadIdx.x >> 5] = a1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l0 = b[0], l1 = b[1];
    for (int i = 0; i < 8; ++i) {
      l0 += s0[i];
      l1 += s1[i];
    }
    const float mx = fmaxf(l0, l1);
    const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
    p[pair] = e1 / (e0 + e1);
  }
}

int itm_head_prob(const float* h, int rows_per_pair, int pairs, const float* w, const float* b, float* p,
                  cudaStream_t st) {
  if (pairs <= 0) return 0;
  itm_head_kernel<<<pairs, 256, 0, st>>>(h, rows_per_pair, w, b, p);
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sprc
