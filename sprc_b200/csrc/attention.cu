// Fused attention  O = softmax(Q K^T * scale + mask) V  for the three attention shapes of the path:
//   ViT MHSA        257 x 257, 16 heads x 64 (CLIP-L) or x 88 (EVA-g)   eva_vit.py:128-145, clip_vit.py:134
//   Q-Former self   32/64 x 32/64, 12 x 64, additive -10000 pad mask    Qformer.py:211-256
//   Q-Former cross  32 x 257 (514 for rerank), 12 x 64                  Qformer.py:191-194,438-450
// One CTA per (sample, head): K and V of that head are staged once in shared memory (zero padded to a
// multiple of 64 keys and to DHP = 64/96 columns), then each warp owns 16-row query tiles and runs an
// online-softmax pass over 64-key chunks with bf16 tensor-core MMAs (m16n8k16, fp32 accumulate) —
// the score matrix never touches HBM (the reference materialises [B,H,257,257] in HBM).
// The attention core is 2.8 % of the ViT FLOPs; the GEMMs around it are tcgen05.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "mma_sync.cuh"
#include "ops.h"
#include "ptx.cuh"

namespace sprc {

static constexpr int ATT_WARPS = 6;

template <int DHP, bool FP16>
__global__ void __launch_bounds__(ATT_WARPS * 32) attention_kernel(const AttnDesc a) {
  constexpr int LDS = DHP + 8;    // padded smem row (elements): conflict-free ldmatrix
  constexpr int KS = DHP / 16;    // k-steps of Q K^T
  constexpr int DT = DHP / 8;     // output n-tiles
  constexpr int VPR = DHP / 8;    // 16-byte vectors per padded row
  extern __shared__ __align__(16) uint8_t att_smem[];
  const int b = blockIdx.y, h = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Lkp = (a.Lk + 63) & ~63;
  bf16* sK = reinterpret_cast<bf16*>(att_smem);
  bf16* sV = sK + (size_t)Lkp * LDS;
  bf16* sQ = sV + (size_t)Lkp * LDS;
  float* sMask = reinterpret_cast<float*>(sQ + ATT_WARPS * 16 * LDS);
  const int dvec = a.dh >> 3;  // valid 16-byte vectors per row (8 or 11)

  // ---- stage K, V (zero padded) and the additive key mask ----
  for (int i = threadIdx.x; i < Lkp * VPR; i += blockDim.x) {
    const int j = i / VPR, vc = i % VPR;
    uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
    if (j < a.Lk && vc < dvec) {
      long long row;
      if (a.kv_idx0) {
        row = j < a.Lk1 ? (long long)a.kv_idx0[b] * a.kv_batch_rows + j
                        : (long long)a.kv_idx1[b] * a.kv_batch_rows + (j - a.Lk1);
      } else {
        row = (long long)b * a.kv_batch_rows + j;
      }
      kv = *reinterpret_cast<const uint4*>(a.K + row * a.ldk + h * a.dh + vc * 8);
      vv = *reinterpret_cast<const uint4*>(a.V + row * a.ldv + h * a.dh + vc * 8);
    }
    *reinterpret_cast<uint4*>(sK + (size_t)j * LDS + vc * 8) = kv;
    *reinterpret_cast<uint4*>(sV + (size_t)j * LDS + vc * 8) = vv;
  }
  for (int j = threadIdx.x; j < Lkp; j += blockDim.x)
    sMask[j] = j < a.Lk ? (a.key_mask ? a.key_mask[(size_t)b * a.Lk + j] : 0.f) : -INFINITY;
  __syncthreads();

  const float sc = a.scale * 1.4426950408889634f;  // fold log2(e): softmax via exp2
  bf16* myQ = sQ + warp * 16 * LDS;
  const int nqt = (a.Lq + 15) >> 4;
  for (int qt = warp; qt < nqt; qt += ATT_WARPS) {
    // ---- this warp's 16 query rows -> smem -> A fragments ----
    __syncwarp();
    for (int i = lane; i < 16 * VPR; i += 32) {
      const int r = i / VPR, vc = i % VPR;
      const int qrow = qt * 16 + r;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (qrow < a.Lq && vc < dvec)
        v = *reinterpret_cast<const uint4*>(a.Q + ((long long)b * a.q_batch_rows + qrow) * a.ldq + h * a.dh +
                                            vc * 8);
      *reinterpret_cast<uint4*>(myQ + r * LDS + vc * 8) = v;
    }
    __syncwarp();
    uint32_t aq[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int r = (lane & 7) + ((lane >> 3) & 1) * 8;
      const int c = ks * 16 + (lane >> 4) * 8;
      ldsm_x4(smem_u32(myQ + r * LDS + c), aq[ks][0], aq[ks][1], aq[ks][2], aq[ks][3]);
    }

    float o[DT][4];
#pragma unroll
    for (int i = 0; i < DT; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;

    for (int c0 = 0; c0 < Lkp; c0 += 64) {
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      // ---- S = Q K^T for 64 keys ----
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          const int key = c0 + (np * 2 + (lane >> 4)) * 8 + (lane & 7);
          const int col = ks * 16 + ((lane >> 3) & 1) * 8;
          uint32_t r0, r1, r2, r3;
          ldsm_x4(smem_u32(sK + (size_t)key * LDS + col), r0, r1, r2, r3);
          mma_16816<FP16>(s[np * 2], aq[ks], r0, r1);
          mma_16816<FP16>(s[np * 2 + 1], aq[ks], r2, r3);
        }
      }
      // ---- scale + mask, chunk row max ----
      float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int col = c0 + nt * 8 + (lane & 3) * 2;
        const float k0 = sMask[col] * 1.4426950408889634f, k1 = sMask[col + 1] * 1.4426950408889634f;
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
      // chunk 0 always holds at least one finite score, so mn_* is finite here
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
      for (int i = 0; i < DT; ++i) {
        o[i][0] *= al_lo;
        o[i][1] *= al_lo;
        o[i][2] *= al_hi;
        o[i][3] *= al_hi;
      }
      // ---- O += P V ----
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t pa[4];
        pa[0] = pack_act(s[2 * kk][0], s[2 * kk][1], FP16);
        pa[1] = pack_act(s[2 * kk][2], s[2 * kk][3], FP16);
        pa[2] = pack_act(s[2 * kk + 1][0], s[2 * kk + 1][1], FP16);
        pa[3] = pack_act(s[2 * kk + 1][2], s[2 * kk + 1][3], FP16);
#pragma unroll
        for (int dp = 0; dp < DT / 2; ++dp) {
          const int key = c0 + kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
          const int col = (dp * 2 + (lane >> 4)) * 8;
          uint32_t r0, r1, r2, r3;
          ldsm_x4_t(smem_u32(sV + (size_t)key * LDS + col), r0, r1, r2, r3);
          mma_16816<FP16>(o[dp * 2], pa, r0, r1);
          mma_16816<FP16>(o[dp * 2 + 1], pa, r2, r3);
        }
      }
    }
    // ---- finalize and store ----
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    const float inv_lo = 1.0f / l_lo, inv_hi = 1.0f / l_hi;
    const int r_lo = qt * 16 + (lane >> 2), r_hi = r_lo + 8;
#pragma unroll
    for (int dt = 0; dt < DT; ++dt) {
      const int col = dt * 8 + (lane & 3) * 2;
      if (col < a.dh) {
        if (r_lo < a.Lq)
          *reinterpret_cast<uint32_t*>(a.O + ((long long)b * a.q_batch_rows + r_lo) * a.ldo + h * a.dh + col) =
              pack_act(o[dt][0] * inv_lo, o[dt][1] * inv_lo, FP16);
        if (r_hi < a.Lq)
          *reinterpret_cast<uint32_t*>(a.O + ((long long)b * a.q_batch_rows + r_hi) * a.ldo + h * a.dh + col) =
              pack_act(o[dt][2] * inv_hi, o[dt][3] * inv_hi, FP16);
      }
    }
  }
}

int attention_small(const AttnDesc& a, cudaStream_t st);  // attention_small.cu
bool attention_cross2_eligible(const AttnDesc& a);          // attention_cross.cu
int attention_cross2(const AttnDesc& a, cudaStream_t st);
bool attention_vit_eligible(const AttnDesc& a);             // attention_vit.cu
int attention_vit(const AttnDesc& a, cudaStream_t st);
bool attention_tc_eligible(const AttnDesc& a);               // attention_tc.cu
bool attention_qf_eligible(const AttnDesc& a);               // attention_qf.cu
int attention_qf(const AttnDesc& a, cudaStream_t st);
int attention_tc(const AttnDesc& a, cudaStream_t st);

template <int DHP, bool FP16>
static int launch_attention(const AttnDesc& a, cudaStream_t st) {
  const int Lkp = (a.Lk + 63) & ~63;
  const size_t smem = ((size_t)2 * Lkp + ATT_WARPS * 16) * (DHP + 8) * sizeof(bf16) + (size_t)Lkp * sizeof(float);
  SPRC_REQUIRE(smem <= 227 * 1024, "attention: Lk=%d needs %zu B of shared memory", a.Lk, smem);
  static size_t configured = 0;
  if (smem > configured) {
    SPRC_CUDA(cudaFuncSetAttribute(attention_kernel<DHP, FP16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid(a.H, a.B);
  prof_begin(st);
  attention_kernel<DHP, FP16><<<grid, ATT_WARPS * 32, smem, st>>>(a);
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

int attention(const AttnDesc& a, cudaStream_t st) {
  SPRC_REQUIRE(a.B > 0 && a.H > 0 && a.Lq > 0 && a.Lk > 0, "attention: empty problem");
  SPRC_REQUIRE(a.dh % 8 == 0 && a.dh <= 96, "attention: head dim %d unsupported", a.dh);
  SPRC_REQUIRE(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 2 == 0,
               "attention: row pitches must keep 16-byte alignment");
  SPRC_REQUIRE(a.B <= 65535, "attention: B=%d exceeds grid limit", a.B);
  static const bool legacy = getenv("SPRC_ATTN_MMA_SYNC") != nullptr;  // A/B switch for tests
  static const bool cross_v1 = getenv("SPRC_CROSS_ATTN_V1") != nullptr;  // A/B switch: first-generation cross kernel
  if (!legacy && !cross_v1 && attention_cross2_eligible(a)) return attention_cross2(a, st);  // 257 / 514 keys
  if (!legacy && attention_qf_eligible(a)) return attention_qf(a, st);  // Q-Former self (/ cross, first generation)
  SPRC_REQUIRE(a.kv_head_stride == 0, "attention: head-major K/V is only read by the tcgen05 cross-attention kernel");
  if (a.dh == 64 && a.Lq <= 64) return attention_small(a, st);  // remaining small shapes (rerank two-segment keys)
  static const bool vit_v1 = getenv("SPRC_VIT_ATTN_V1") != nullptr;     // A/B switch: first-generation ViT kernel
  if (!legacy && !vit_v1 && attention_vit_eligible(a)) return attention_vit(a, st);  // ViT: two tiles in flight
  if (!legacy && attention_tc_eligible(a)) return attention_tc(a, st);  // ViT, first generation
  if (act_fp16()) return a.dh <= 64 ? launch_attention<64, true>(a, st) : launch_attention<96, true>(a, st);
  if (a.dh <= 64) return launch_attention<64, false>(a, st);
  return launch_attention<96, false>(a, st);
}

}  // namespace sprc
