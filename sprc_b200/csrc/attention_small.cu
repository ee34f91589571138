// Attention for the Q-Former's small shapes (dh = 64, Lq <= 64):
//   self-attention   Lq = Lk = 32 / 64 with the additive pad mask      Qformer.py:211-256
//   cross-attention  Lq = 32, Lk = 257 (514 = two KV segments, rerank)  Qformer.py:191-194,438-450
// These are latency-bound, not FLOP-bound (a (sample, head) item is ~1 MFLOP over 24-70 KB), so the kernel is
// persistent: each CTA walks over items with a two-stage shared-memory ring and prefetches the next item's
// Q, K, V with cp.async (16 B per request) while the tensor cores work on the current one.
// The 8 warps of a CTA are split as (items per iteration) x (16-row query tiles) x (key splits):
//   self  S = 64 : 2 items x 4 tiles x 1 split      self S = 32 : 4 items x 2 tiles x 1 split
//   cross Lq = 32: 1 item  x 2 tiles x 4 splits - the four warps of a tile take alternate 64-key chunks and
//                  merge their online-softmax partials (m, l, O) through shared memory.
// bf16 mma.sync.m16n8k16 with fp32 softmax; the score matrix never leaves the SM.
#include <math.h>
#include <stdio.h>

#include "mma_sync.cuh"
#include "ops.h"
#include "ptx.cuh"

namespace sprc {

static constexpr int SA_WARPS = 8;
static constexpr int SA_LDS = 72;  // 64 + 8 bf16: conflict-free ldmatrix rows
static constexpr int SA_PART = 16 * 64 + 32;  // floats per warp in the split-merge buffer

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct SmallAttnLayout {
  int Lkp;          // keys padded to 64
  int q_rows;       // query rows staged per item (Lq padded to 16*NQT)
  int item_bytes;   // K + V + Q + mask of one item
  int stage_bytes;  // ITEMS items
};

template <int NQT, int KSPLIT, bool FP16>
__global__ void __launch_bounds__(SA_WARPS * 32)
attention_small_kernel(const AttnDesc a, const SmallAttnLayout lay, const int nstages) {
  constexpr int ITEMS = SA_WARPS / (NQT * KSPLIT);
  extern __shared__ __align__(16) uint8_t sa_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Lkp = lay.Lkp;
  const int n_items = a.B * a.H;
  const int n_iters = (n_items + ITEMS - 1) / ITEMS;
  float* sPart = reinterpret_cast<float*>(sa_smem + (size_t)nstages * lay.stage_bytes);

  auto item_ptr = [&](int stage, int slot) { return sa_smem + (size_t)stage * lay.stage_bytes + (size_t)slot * lay.item_bytes; };

  // issue the loads of iteration `it` into `stage`
  auto prefetch = [&](int it, int stage) {
    for (int slot = 0; slot < ITEMS; ++slot) {
      const int item = it * ITEMS + slot;
      if (item >= n_items) break;
      const int b = item / a.H, h = item % a.H;
      bf16* sK = reinterpret_cast<bf16*>(item_ptr(stage, slot));
      bf16* sV = sK + (size_t)Lkp * SA_LDS;
      bf16* sQ = sV + (size_t)Lkp * SA_LDS;
      float* sMask = reinterpret_cast<float*>(sQ + (size_t)lay.q_rows * SA_LDS);
      // thread t serves 16-byte column chunk (t & 7) of rows (t >> 3) + 32*k: pointers advance by a constant
      // per request (address generation was a third of the instruction stream in the first version)
      const int vc = threadIdx.x & 7, j0 = threadIdx.x >> 3;
      {
        bf16* dk = sK + (size_t)j0 * SA_LDS + vc * 8;
        bf16* dv = sV + (size_t)j0 * SA_LDS + vc * 8;
        if (!a.kv_idx0) {
          const bf16* gk = a.K + ((long long)b * a.kv_batch_rows + j0) * a.ldk + h * 64 + vc * 8;
          const bf16* gv = a.V + ((long long)b * a.kv_batch_rows + j0) * a.ldv + h * 64 + vc * 8;
          const long long sk = 32LL * a.ldk, sv = 32LL * a.ldv;
          for (int j = j0; j < Lkp; j += 32, dk += 32 * SA_LDS, dv += 32 * SA_LDS, gk += sk, gv += sv) {
            if (j < a.Lk) {
              cp_async16(dk, gk);
              cp_async16(dv, gv);
            } else {
              *reinterpret_cast<uint4*>(dk) = make_uint4(0, 0, 0, 0);
              *reinterpret_cast<uint4*>(dv) = make_uint4(0, 0, 0, 0);
            }
          }
        } else {
          const long long r0 = (long long)a.kv_idx0[b] * a.kv_batch_rows, r1 = (long long)a.kv_idx1[b] * a.kv_batch_rows;
          for (int j = j0; j < Lkp; j += 32, dk += 32 * SA_LDS, dv += 32 * SA_LDS) {
            if (j < a.Lk) {
              const long long row = j < a.Lk1 ? r0 + j : r1 + (j - a.Lk1);
              cp_async16(dk, a.K + row * a.ldk + h * 64 + vc * 8);
              cp_async16(dv, a.V + row * a.ldv + h * 64 + vc * 8);
            } else {
              *reinterpret_cast<uint4*>(dk) = make_uint4(0, 0, 0, 0);
              *reinterpret_cast<uint4*>(dv) = make_uint4(0, 0, 0, 0);
            }
          }
        }
        bf16* dq = sQ + (size_t)j0 * SA_LDS + vc * 8;
        const bf16* gq = a.Q + ((long long)b * a.q_batch_rows + j0) * a.ldq + h * 64 + vc * 8;
        for (int r = j0; r < lay.q_rows; r += 32, dq += 32 * SA_LDS, gq += 32LL * a.ldq) {
          if (r < a.Lq)
            cp_async16(dq, gq);
          else
            *reinterpret_cast<uint4*>(dq) = make_uint4(0, 0, 0, 0);
        }
      }
      for (int j = threadIdx.x; j < Lkp; j += blockDim.x)
        sMask[j] = j < a.Lk ? (a.key_mask ? a.key_mask[(size_t)b * a.Lk + j] * 1.4426950408889634f : 0.f) : -INFINITY;
    }
    cp_async_commit();
  };

  const int slot = warp / (NQT * KSPLIT);
  const int qt = warp % NQT;
  const int ks = (warp / NQT) % KSPLIT;
  const float sc = a.scale * 1.4426950408889634f;
  const int nqt_valid = (a.Lq + 15) >> 4;

  griddep_wait();
  griddep_launch();
  if (blockIdx.x < n_iters) prefetch(blockIdx.x, 0);
  int iter = 0;
  for (int it = blockIdx.x; it < n_iters; it += gridDim.x, ++iter) {
    const int stage = nstages == 2 ? (iter & 1) : 0;
    const int it_next = it + gridDim.x;
    if (nstages == 2 && it_next < n_iters) {
      prefetch(it_next, stage ^ 1);
      cp_async_wait_group<1>();
    } else {
      if (nstages == 1 && iter > 0) prefetch(it, 0);
      cp_async_wait_group<0>();
    }
    __syncthreads();

    const int item = it * ITEMS + slot;
    const bool active = item < n_items && qt < nqt_valid;
    const int b = active ? item / a.H : 0, h = active ? item % a.H : 0;
    bf16* sK = reinterpret_cast<bf16*>(item_ptr(stage, slot));
    bf16* sV = sK + (size_t)Lkp * SA_LDS;
    bf16* sQ = sV + (size_t)Lkp * SA_LDS;
    const float* sMask = reinterpret_cast<const float*>(sQ + (size_t)lay.q_rows * SA_LDS);

    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;

    if (active) {
      uint32_t aq[4][4];
#pragma unroll
      for (int kq = 0; kq < 4; ++kq) {
        const int r = qt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = kq * 16 + (lane >> 4) * 8;
        ldsm_x4(smem_u32(sQ + r * SA_LDS + c), aq[kq][0], aq[kq][1], aq[kq][2], aq[kq][3]);
      }
      for (int c0 = ks * 64; c0 < Lkp; c0 += KSPLIT * 64) {
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
        for (int kq = 0; kq < 4; ++kq) {
#pragma unroll
          for (int np = 0; np < 4; ++np) {
            const int key = c0 + (np * 2 + (lane >> 4)) * 8 + (lane & 7);
            const int col = kq * 16 + ((lane >> 3) & 1) * 8;
            uint32_t r0, r1, r2, r3;
            ldsm_x4(smem_u32(sK + (size_t)key * SA_LDS + col), r0, r1, r2, r3);
            mma_16816<FP16>(s[np * 2], aq[kq], r0, r1);
            mma_16816<FP16>(s[np * 2 + 1], aq[kq], r2, r3);
          }
        }
        float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int col = c0 + nt * 8 + (lane & 3) * 2;
          const float k0 = sMask[col], k1 = sMask[col + 1];
          s[nt][0] = s[nt][0] * sc + k0;
          s[nt][1] = s[nt][1] * sc + k1;
          s[nt][2] = s[nt][2] * sc + k0;
          s[nt][3] = s[nt][3] * sc + k1;
          mx_lo = fmaxf(mx_lo, fmaxf(s[nt][0], s[nt][1]));
          mx_hi = fmaxf(mx_hi, fmaxf(s[nt][2], s[nt][3]));
        }
        mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
        mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
        mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
        mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
        const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);
        // every 64-key chunk below Lkp holds at least one finite score, so mn_* is finite
        const float al_lo = exp2f(m_lo - mn_lo), al_hi = exp2f(m_hi - mn_hi);
        m_lo = mn_lo;
        m_hi = mn_hi;
        float rs_lo = 0.f, rs_hi = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          s[nt][0] = exp2f(s[nt][0] - mn_lo);
          s[nt][1] = exp2f(s[nt][1] - mn_lo);
          s[nt][2] = exp2f(s[nt][2] - mn_hi);
          s[nt][3] = exp2f(s[nt][3] - mn_hi);
          rs_lo += s[nt][0] + s[nt][1];
          rs_hi += s[nt][2] + s[nt][3];
        }
        l_lo = l_lo * al_lo + rs_lo;
        l_hi = l_hi * al_hi + rs_hi;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          o[i][0] *= al_lo;
          o[i][1] *= al_lo;
          o[i][2] *= al_hi;
          o[i][3] *= al_hi;
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint32_t pa[4];
          pa[0] = pack_act(s[2 * kk][0], s[2 * kk][1], FP16);
          pa[1] = pack_act(s[2 * kk][2], s[2 * kk][3], FP16);
          pa[2] = pack_act(s[2 * kk + 1][0], s[2 * kk + 1][1], FP16);
          pa[3] = pack_act(s[2 * kk + 1][2], s[2 * kk + 1][3], FP16);
#pragma unroll
          for (int dp = 0; dp < 4; ++dp) {
            const int key = c0 + kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
            const int col = (dp * 2 + (lane >> 4)) * 8;
            uint32_t r0, r1, r2, r3;
            ldsm_x4_t(smem_u32(sV + (size_t)key * SA_LDS + col), r0, r1, r2, r3);
            mma_16816<FP16>(o[dp * 2], pa, r0, r1);
            mma_16816<FP16>(o[dp * 2 + 1], pa, r2, r3);
          }
        }
      }
      l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
      l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
      l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
      l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    }

    if constexpr (KSPLIT > 1) {
      // ---- merge the key splits of each query tile (split 0 of the tile is the owner) ----
      float* part = sPart + warp * SA_PART;
      if (active && ks > 0) {
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
          const int col = dt * 8 + (lane & 3) * 2;
          *reinterpret_cast<float2*>(part + (lane >> 2) * 64 + col) = make_float2(o[dt][0], o[dt][1]);
          *reinterpret_cast<float2*>(part + ((lane >> 2) + 8) * 64 + col) = make_float2(o[dt][2], o[dt][3]);
        }
        if ((lane & 3) == 0) {
          part[1024 + (lane >> 2)] = m_lo;
          part[1024 + 8 + (lane >> 2)] = m_hi;
          part[1024 + 16 + (lane >> 2)] = l_lo;
          part[1024 + 24 + (lane >> 2)] = l_hi;
        }
      }
      __syncthreads();
      if (active && ks == 0) {
        for (int s2 = 1; s2 < KSPLIT; ++s2) {
          const float* op = sPart + (warp + s2 * NQT) * SA_PART;
          const float pm_lo = op[1024 + (lane >> 2)], pm_hi = op[1024 + 8 + (lane >> 2)];
          const float pl_lo = op[1024 + 16 + (lane >> 2)], pl_hi = op[1024 + 24 + (lane >> 2)];
          if (pm_lo == -INFINITY) continue;  // that split had no chunk (warp-uniform: depends on Lkp only)
          const float mn_lo = fmaxf(m_lo, pm_lo), mn_hi = fmaxf(m_hi, pm_hi);
          const float a_lo = exp2f(m_lo - mn_lo), a_hi = exp2f(m_hi - mn_hi);
          const float b_lo = exp2f(pm_lo - mn_lo), b_hi = exp2f(pm_hi - mn_hi);
          m_lo = mn_lo;
          m_hi = mn_hi;
          l_lo = l_lo * a_lo + pl_lo * b_lo;
          l_hi = l_hi * a_hi + pl_hi * b_hi;
#pragma unroll
          for (int dt = 0; dt < 8; ++dt) {
            const int col = dt * 8 + (lane & 3) * 2;
            const float2 x = *reinterpret_cast<const float2*>(op + (lane >> 2) * 64 + col);
            const float2 y = *reinterpret_cast<const float2*>(op + ((lane >> 2) + 8) * 64 + col);
            o[dt][0] = o[dt][0] * a_lo + x.x * b_lo;
            o[dt][1] = o[dt][1] * a_lo + x.y * b_lo;
            o[dt][2] = o[dt][2] * a_hi + y.x * b_hi;
            o[dt][3] = o[dt][3] * a_hi + y.y * b_hi;
          }
        }
      }
    }
    if (active && ks == 0) {
      const float inv_lo = 1.0f / l_lo, inv_hi = 1.0f / l_hi;
      const int r_lo = qt * 16 + (lane >> 2), r_hi = r_lo + 8;
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        const int col = dt * 8 + (lane & 3) * 2;
        if (r_lo < a.Lq)
          *reinterpret_cast<uint32_t*>(a.O + ((long long)b * a.q_batch_rows + r_lo) * a.ldo + h * 64 + col) =
              pack_act(o[dt][0] * inv_lo, o[dt][1] * inv_lo, FP16);
        if (r_hi < a.Lq)
          *reinterpret_cast<uint32_t*>(a.O + ((long long)b * a.q_batch_rows + r_hi) * a.ldo + h * 64 + col) =
              pack_act(o[dt][2] * inv_hi, o[dt][3] * inv_hi, FP16);
      }
    }
    __syncthreads();  // every warp is done with this stage (and with sPart) before it is refilled
  }
}

template <int NQT, int KSPLIT, bool FP16>
static int launch_small(const AttnDesc& a, cudaStream_t st) {
  constexpr int ITEMS = SA_WARPS / (NQT * KSPLIT);
  SmallAttnLayout lay;
  lay.Lkp = (a.Lk + 63) & ~63;
  lay.q_rows = NQT * 16;
  lay.item_bytes = (2 * lay.Lkp + lay.q_rows) * SA_LDS * (int)sizeof(bf16) + lay.Lkp * (int)sizeof(float);
  lay.item_bytes = (lay.item_bytes + 15) & ~15;
  lay.stage_bytes = ITEMS * lay.item_bytes;
  const size_t part_bytes = KSPLIT > 1 ? (size_t)SA_WARPS * SA_PART * sizeof(float) : 0;
  int nstages = 2;
  if (2 * (size_t)lay.stage_bytes + part_bytes > 227 * 1024) nstages = 1;
  const size_t smem = (size_t)nstages * lay.stage_bytes + part_bytes;
  SPRC_REQUIRE(smem <= 227 * 1024, "attention_small: Lk=%d needs %zu B of shared memory", a.Lk, smem);
  static size_t configured = 0;
  if (smem > configured) {
    SPRC_CUDA(cudaFuncSetAttribute(attention_small_kernel<NQT, KSPLIT, FP16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    configured = smem;
  }
  const int n_iters = (a.B * a.H + ITEMS - 1) / ITEMS;
  const int ctas_per_sm = smem <= 112 * 1024 ? 2 : 1;
  int grid = device_sm_count() * ctas_per_sm;
  if (grid > n_iters) grid = n_iters;
  prof_begin(st);
  SPRC_CUDA(launch_pdl(attention_small_kernel<NQT, KSPLIT, FP16>, dim3(grid), dim3(SA_WARPS * 32), smem, st, a, lay,
                       nstages));
  if (prof_enabled()) {
    char tag[56];
    snprintf(tag, sizeof(tag), "B%d H%d dh%d Lq%d Lk%d", a.B, a.H, a.dh, a.Lq, a.Lk);
    prof_end(PROF_ATTN, 4.0 * a.B * a.H * (double)a.Lq * a.Lk * a.dh,
             2.0 * a.B * a.H * a.dh * (2.0 * a.Lq + 2.0 * a.Lk), st, tag);
  }
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

int attention_small(const AttnDesc& a, cudaStream_t st) {
  if (act_fp16()) {
    if (a.Lk > 64 && a.Lq > 32) return launch_small<4, 2, true>(a, st);
    if (a.Lk > 64) return launch_small<2, 4, true>(a, st);
    if (a.Lq > 32) return launch_small<4, 1, true>(a, st);
    return launch_small<2, 1, true>(a, st);
  }
  if (a.Lk > 64 && a.Lq > 32) return launch_small<4, 2, false>(a, st);  // (not on the reference path)
  if (a.Lk > 64) return launch_small<2, 4, false>(a, st);            // cross-attention, Lq <= 32
  if (a.Lq > 32) return launch_small<4, 1, false>(a, st);            // self-attention S = 64
  return launch_small<2, 1, false>(a, st);                           // self-attention S = 32
}

}  // namespace sprc
