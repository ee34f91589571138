// Q-Former attention on tcgen05 (sm_100a), dh = 64, 12 heads:
//   self-attention   S = 32 (gallery pass) or 64 (fusion / text pass) rows per sample, additive pad mask
//                    Qformer.py:211-256 (scores, mask, softmax, context), :133-139 (q/k/v projections)
//   cross-attention  32 query rows per sample over the 257 visual tokens of that sample
//                    Qformer.py:191-194,438-450
// Both kernels are persistent and warp-specialised like the ViT attention (attention_tc.cu):
//   warp 0      TMA producer (2-stage ring of Q, K, V tiles, 128-byte swizzle)
//   warp 1      tcgen05.mma issuer, accumulators in TMEM
//   warps 2..5  softmax + epilogue: one thread per query row (TMEM lane), ex2.approx with the scale folded in,
//               output tile staged in shared memory and written with ONE TMA store.
// Self-attention: 128 consecutive rows of the packed [rows, 3*768] QKV activation hold 128/S samples; ONE
// 128x128x64 MMA computes all their score blocks (the off-diagonal blocks are never read), P is written to
// TMEM as a block-diagonal bf16 matrix (off-diagonal columns zeroed once per CTA) and O = P V is a TMEM-A MMA
// with V as an MN-major operand - no per-sample MMAs, no shuffles: each thread's softmax row is in registers.
// Cross-attention: the 32 query rows are replicated into the four TMEM lane quarters (four 4 KB TMA loads of
// the same box), so the four softmax warps each read the SAME rows from their own quarter and split the 272
// (padded) keys between them; P goes to shared memory as a K-major A operand whose rows 0..31 are the real
// query rows, O = P V accumulates all keys, and the lane-quarter-0 warp writes the 32 output rows.
#include <math.h>
#include <stdio.h>

#include "ops.h"
#include "ptx.cuh"

namespace sprc {

int make_tmap_bf16(CUtensorMap* tm, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
                   uint64_t stride2, uint32_t b0, uint32_t b1, uint32_t b2, int rank);

static constexpr float kLog2e = 1.4426950408889634f;
static constexpr int QF_THREADS = 6 * 32;

// ================================================================================================
// self-attention
// ================================================================================================
static constexpr int QS_TILE = 128 * 128;     // 128 rows x 64 bf16
static constexpr int QS_STAGE = 3 * QS_TILE;  // Q, K, V
static constexpr int QS_COL_S = 0, QS_COL_P = 128, QS_COL_O = 192;  // TMEM columns (256 allocated)
static constexpr int QS_SMEM = 2 * QS_STAGE + 256 + 1024;

struct QfSelfParams {
  int rows, B, H;
  float scale_log2;
  int fp16;
  int rev;  // sweep the items from the last one down (next_sweep_reverse)
  const float* key_mask;  // additive [B, S] or null
};

template <int S>
__global__ void __launch_bounds__(QF_THREADS, 2)
qf_self_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                            const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                            const QfSelfParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * QS_STAGE);
  uint64_t* full = bars;        // [2]
  uint64_t* empty = bars + 2;   // [2]
  uint64_t* s_full = bars + 4;
  uint64_t* p_full = bars + 5;
  uint64_t* o_full = bars + 6;
  uint64_t* o_empty = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = ((p.rows + 127) / 128) * p.H;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 4);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();
  griddep_launch();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        const int itm = p.rev ? n_items - 1 - item : item;
        const int g = itm / p.H, h = itm % p.H;
        uint8_t* st = smem + s * QS_STAGE;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], QS_STAGE);
        tma_load_2d(&tmQ, &full[s], st, h * 64, g * 128, kEvictFirst);
        tma_load_2d(&tmK, &full[s], st + QS_TILE, h * 64, g * 128, kEvictFirst);
        tma_load_2d(&tmV, &full[s], st + 2 * QS_TILE, h * 64, g * 128, kEvictFirst);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_s = umma_idesc_16(128, 128, p.fp16);
    const uint32_t idesc_pv = umma_idesc_16(128, 64, p.fp16) | (1u << 16);  // B operand MN-major
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int s = it & 1;
      uint8_t* st = smem + s * QS_STAGE;
      mbar_wait(&full[s], (it >> 1) & 1);
      tc_fence_after();
      // S is free: the softmax warps signalled p_full for the previous item before its PV was issued
      if (elect_one()) {
        const uint64_t da = umma_desc_k_sw128(smem_u32(st));
        const uint64_t db = umma_desc_k_sw128(smem_u32(st + QS_TILE));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16(tmem_base + QS_COL_S, da + 2 * kk, db + 2 * kk, idesc_s, kk != 0 ? 1u : 0u);
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, it & 1);          // P(it) is in TMEM, S(it) consumed
      mbar_wait(o_empty, (it & 1) ^ 1);   // the epilogue of the previous item has read O
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t db = umma_desc_mn_sw128(smem_u32(st + 2 * QS_TILE + ks * 16 * 128), QS_TILE);
          umma_bf16_ts(tmem_base + QS_COL_O, tmem_base + QS_COL_P + ks * 8, db, idesc_pv, ks != 0 ? 1u : 0u);
        }
        umma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax + epilogue (warps 2..5) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;        // row of the 128-row tile = TMEM lane
    const int blk = row / S;              // this row's sample inside the tile
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    {
      // block-diagonal P: the columns outside this row's own block stay zero for the whole kernel
      uint32_t z[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) z[j] = 0u;
      tmem_st32(tmem_base + lane_addr + QS_COL_P, z);
      tmem_st32(tmem_base + lane_addr + QS_COL_P + 32, z);
      tmem_st_wait();
    }
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int s = it & 1;
      const int itm = p.rev ? n_items - 1 - item : item;
        const int g = itm / p.H, h = itm % p.H;
      const int sample = (g * 128 + row) / S;
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      uint32_t sr[S];
      {
        const uint32_t s_addr = tmem_base + lane_addr + QS_COL_S + blk * S;
        uint32_t(&a0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sr[0]);
        tmem_ld32(s_addr, a0);
        if constexpr (S == 64) {
          uint32_t(&a1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sr[32]);
          tmem_ld32(s_addr + 32, a1);
        }
        tmem_ld_wait();
      }
      float v[S];
      float mx = -INFINITY;
      if (p.key_mask && sample < p.B) {
        const float4* mk = reinterpret_cast<const float4*>(p.key_mask + static_cast<size_t>(sample) * S);
#pragma unroll
        for (int j = 0; j < S; j += 4) {
          const float4 m4 = __ldg(mk + j / 4);
          v[j] = fmaf(__uint_as_float(sr[j]), p.scale_log2, m4.x * kLog2e);
          v[j + 1] = fmaf(__uint_as_float(sr[j + 1]), p.scale_log2, m4.y * kLog2e);
          v[j + 2] = fmaf(__uint_as_float(sr[j + 2]), p.scale_log2, m4.z * kLog2e);
          v[j + 3] = fmaf(__uint_as_float(sr[j + 3]), p.scale_log2, m4.w * kLog2e);
        }
      } else {
#pragma unroll
        for (int j = 0; j < S; ++j) v[j] = __uint_as_float(sr[j]) * p.scale_log2;
      }
#pragma unroll
      for (int j = 0; j < S; ++j) mx = fmaxf(mx, v[j]);
      float sum = 0.f;
      uint32_t pk[S / 2];
#pragma unroll
      for (int j = 0; j < S; j += 2) {
        const float e0 = ex2_approx(v[j] - mx), e1 = ex2_approx(v[j + 1] - mx);
        sum += e0 + e1;
        pk[j / 2] = pack_act(e0, e1, p.fp16);
      }
      {
        const uint32_t p_addr = tmem_base + lane_addr + QS_COL_P + blk * (S / 2);
        if constexpr (S == 64) {
          tmem_st32(p_addr, pk);
        } else {
          tmem_st16(p_addr, pk);
        }
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // ---- epilogue: O / l -> bf16 -> swizzled staging (the Q tile of this stage, dead since S was formed) ----
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      uint32_t r[64];
      {
        uint32_t(&a0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
        uint32_t(&a1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[32]);
        tmem_ld32(tmem_base + lane_addr + QS_COL_O, a0);
        tmem_ld32(tmem_base + lane_addr + QS_COL_O + 32, a1);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      const float inv = 1.0f / sum;
      const uint32_t stg = smem_u32(smem + s * QS_STAGE) + row * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        sts128(stg + ((c ^ (row & 7)) << 4),
               pack_act(__uint_as_float(r[8 * c]) * inv, __uint_as_float(r[8 * c + 1]) * inv, p.fp16),
               pack_act(__uint_as_float(r[8 * c + 2]) * inv, __uint_as_float(r[8 * c + 3]) * inv, p.fp16),
               pack_act(__uint_as_float(r[8 * c + 4]) * inv, __uint_as_float(r[8 * c + 5]) * inv, p.fp16),
               pack_act(__uint_as_float(r[8 * c + 6]) * inv, __uint_as_float(r[8 * c + 7]) * inv, p.fp16));
      fence_proxy_async();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == 2 && lane == 0) {
        tma_store_2d(&tmO, smem_u32(smem + s * QS_STAGE), h * 64, g * 128);  // rows >= p.rows are clipped
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(&empty[s]);  // Q (staging), K, V of this stage may be refilled
      }
    }
    if (warp == 2 && lane == 0) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

template <int S>
static int launch_qf_self(const AttnDesc& a, cudaStream_t st) {
  CUtensorMap tmQ, tmK, tmV, tmO;
  const uint64_t rows = (uint64_t)a.B * S;
  const uint64_t w = (uint64_t)a.H * 64;
  SPRC_TRY(make_tmap_bf16(&tmQ, a.Q, w, rows, 1, a.ldq, 0, 64, 128, 1, 2));
  SPRC_TRY(make_tmap_bf16(&tmK, a.K, w, rows, 1, a.ldk, 0, 64, 128, 1, 2));
  SPRC_TRY(make_tmap_bf16(&tmV, a.V, w, rows, 1, a.ldv, 0, 64, 128, 1, 2));
  SPRC_TRY(make_tmap_bf16(&tmO, a.O, w, rows, 1, a.ldo, 0, 64, 128, 1, 2));
  QfSelfParams p;
  p.rows = (int)rows;
  p.B = a.B;
  p.H = a.H;
  p.scale_log2 = a.scale * kLog2e;
  p.fp16 = act_fp16();
  p.rev = next_sweep_reverse();
  p.key_mask = a.key_mask;
  static bool attr_set = false;
  if (!attr_set) {
    SPRC_CUDA(cudaFuncSetAttribute(qf_self_attention_tc_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   QS_SMEM));
    attr_set = true;
  }
  const int items = (int)((rows + 127) / 128) * a.H;
  int grid = 2 * device_sm_count();
  if (grid > items) grid = items;
  prof_begin(st);
  SPRC_CUDA(launch_pdl(qf_self_attention_tc_kernel<S>, dim3(grid), dim3(QF_THREADS), QS_SMEM, st, tmQ, tmK, tmV, tmO,
                       p));
  if (prof_enabled()) {
    char tag[56];
    snprintf(tag, sizeof(tag), "qf-self tc B%d H%d S%d", a.B, a.H, S);
    prof_end(PROF_ATTN, 4.0 * a.B * a.H * (double)S * S * 64, 2.0 * a.B * a.H * 64 * 4.0 * S, st, tag);
  }
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

// ================================================================================================
// cross-attention
// ================================================================================================
static constexpr int QC_LK = 272;              // keys padded to the MMA K step
static constexpr int QC_HALF = QC_LK / 2;      // TMA box rows (<= 256)
static constexpr int QC_QBYTES = 128 * 128;    // four replicas of the 32-row query tile
static constexpr int QC_KBYTES = QC_LK * 128;
static constexpr int QC_STAGE = QC_QBYTES + 2 * QC_KBYTES;
static constexpr int QC_PBYTES = 32 * 1024;    // five 64-key blocks 4 KB apart + the rows 32..127 the MMA also reads
static constexpr int QC_OBYTES = 32 * 128;     // output staging
static constexpr int QC_COL_S = 0, QC_COL_O = 288;
static constexpr int QC_SMEM = QC_PBYTES + 2 * QC_STAGE + QC_OBYTES + 2048 + 1024;

struct QfCrossParams {
  int B, H, Lk;
  int q_batch_rows, kv_batch_rows;
  float scale_log2;
  int fp16;
  int rev;  // sweep the items from the last one down (next_sweep_reverse)
};

__global__ void __launch_bounds__(QF_THREADS, 1)
qf_cross_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                             const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                             const QfCrossParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sP = smem;
  uint8_t* stages = smem + QC_PBYTES;
  uint8_t* sO = stages + 2 * QC_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO + QC_OBYTES);
  uint64_t* full = bars;        // [2]
  uint64_t* empty = bars + 2;   // [2]
  uint64_t* s_full = bars + 4;
  uint64_t* p_full = bars + 5;
  uint64_t* o_full = bars + 6;
  uint64_t* o_empty = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const uint32_t xch = smem_u32(bars + 10);  // [4 key blocks][32 rows] row max, then [4][32] row sum

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.B * p.H;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();
  griddep_launch();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int s = it & 1;
        const int itm = p.rev ? n_items - 1 - item : item;
        const int b = itm / p.H, h = itm % p.H;
        uint8_t* st = stages + s * QC_STAGE;
        mbar_wait(&empty[s], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&full[s], QC_STAGE);
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
          tma_load_2d(&tmQ, &full[s], st + rep * 4096, h * 64, b * p.q_batch_rows, kEvictNormal);
        const int kr = b * p.kv_batch_rows;
        // K/V maps are (d, row, head): heads are column slices of wide rows or contiguous [rows, 64] blocks
        tma_load_3d(&tmK, &full[s], st + QC_QBYTES, 0, kr, h, kEvictFirst);
        tma_load_3d(&tmK, &full[s], st + QC_QBYTES + QC_HALF * 128, 0, kr + QC_HALF, h, kEvictFirst);
        tma_load_3d(&tmV, &full[s], st + QC_QBYTES + QC_KBYTES, 0, kr, h, kEvictFirst);
        tma_load_3d(&tmV, &full[s], st + QC_QBYTES + QC_KBYTES + QC_HALF * 128, 0, kr + QC_HALF, h, kEvictFirst);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_s256 = umma_idesc_16(128, 256, p.fp16);
    const uint32_t idesc_s16 = umma_idesc_16(128, 16, p.fp16);
    const uint32_t idesc_pv = umma_idesc_16(128, 64, p.fp16) | (1u << 16);  // B operand MN-major
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int s = it & 1;
      uint8_t* st = stages + s * QC_STAGE;
      mbar_wait(&full[s], (it >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = umma_desc_k_sw128(smem_u32(st));
        const uint64_t db = umma_desc_k_sw128(smem_u32(st + QC_QBYTES));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          umma_bf16(tmem_base + QC_COL_S, da + 2 * kk, db + 2 * kk, idesc_s256, kk != 0 ? 1u : 0u);
          umma_bf16(tmem_base + QC_COL_S + 256, da + 2 * kk, db + 2 * kk + ((256 * 128) >> 4), idesc_s16,
                    kk != 0 ? 1u : 0u);
        }
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, it & 1);          // P(it) is in shared memory, S(it) consumed
      mbar_wait(o_empty, (it & 1) ^ 1);   // the previous item's O has been read
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < QC_LK / 16; ++ks) {
          const uint64_t da = umma_desc_k_sw128(smem_u32(sP + (ks >> 2) * 4096)) + 2 * (ks & 3);
          const uint64_t db = umma_desc_mn_sw128(smem_u32(st + QC_QBYTES + QC_KBYTES + ks * 16 * 128), QC_KBYTES);
          umma_bf16(tmem_base + QC_COL_O, da, db, idesc_pv, ks != 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(&empty[s]);  // Q, K, V of this stage are free once these MMAs have read them
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax + epilogue (warps 2..5) =====================
    const int q = warp & 3;   // TMEM lane quarter = 64-key block this warp owns (quarter 3 also takes keys 256..271)
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t prow = smem_u32(sP) + lane * 128;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int itm = p.rev ? n_items - 1 - item : item;
        const int b = itm / p.H, h = itm % p.H;
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      uint32_t sr[80];
      {
        uint32_t(&a0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sr[0]);
        uint32_t(&a1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sr[32]);
        uint32_t(&a2)[16] = *reinterpret_cast<uint32_t(*)[16]>(&sr[64]);
        const uint32_t s_addr = tmem_base + lane_addr + QC_COL_S + q * 64;
        tmem_ld32(s_addr, a0);
        tmem_ld32(s_addr + 32, a1);
        if (q == 3) tmem_ld16(tmem_base + lane_addr + QC_COL_S + 256, a2);
        tmem_ld_wait();
      }
      const int nk = q == 3 ? 80 : 64;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 80; ++j) {
        const int key = j < 64 ? q * 64 + j : 256 + (j - 64);
        if (j < nk && key < p.Lk) mx = fmaxf(mx, __uint_as_float(sr[j]));
      }
      sts32f(xch + (q * 32 + lane) * 4, mx);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mx = fmaxf(fmaxf(lds32f(xch + lane * 4), lds32f(xch + (32 + lane) * 4)),
                 fmaxf(lds32f(xch + (64 + lane) * 4), lds32f(xch + (96 + lane) * 4)));
      const float moff = mx * p.scale_log2;
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 10; ++c) {
        if (c < 8 || q == 3) {
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            const int jj = c * 8 + j;
            const int key = jj < 64 ? q * 64 + jj : 256 + (jj - 64);
            float e0 = ex2_approx(fmaf(__uint_as_float(sr[jj]), p.scale_log2, -moff));
            float e1 = ex2_approx(fmaf(__uint_as_float(sr[jj + 1]), p.scale_log2, -moff));
            if (key >= p.Lk) e0 = 0.f;
            if (key + 1 >= p.Lk) e1 = 0.f;
            sum += e0 + e1;
            pk[j / 2] = pack_act(e0, e1, p.fp16);
          }
          // K-major A operand: key block kb at sP + kb*4096, row `lane`, 16-byte chunk cc (8 keys), 128B swizzle
          const int kb = c < 8 ? q : 4, cc = c < 8 ? c : c - 8;
          sts128(prow + kb * 4096 + ((cc ^ (lane & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
        }
      }
      sts32f(xch + (128 + q * 32 + lane) * 4, sum);
      fence_proxy_async();   // P (generic-proxy writes) -> tcgen05.mma (async proxy)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      tc_fence_before();
      if (lane == 0) mbar_arrive(p_full);
      if (q == 0) {
        // ---- epilogue (lanes 0..31 hold the real rows): O / l -> bf16 -> staging -> TMA store of 32 rows ----
        mbar_wait(o_full, it & 1);
        tc_fence_after();
        uint32_t r[64];
        {
          uint32_t(&a0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
          uint32_t(&a1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[32]);
          tmem_ld32(tmem_base + QC_COL_O, a0);
          tmem_ld32(tmem_base + QC_COL_O + 32, a1);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(o_empty);
          bulk_wait_read0();  // the previous item's store has read the staging tile
        }
        __syncwarp();
        const float inv = 1.0f / ((lds32f(xch + (128 + lane) * 4) + lds32f(xch + (160 + lane) * 4)) +
                                  (lds32f(xch + (192 + lane) * 4) + lds32f(xch + (224 + lane) * 4)));
        const uint32_t stg = smem_u32(sO) + lane * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          sts128(stg + ((c ^ (lane & 7)) << 4),
                 pack_act(__uint_as_float(r[8 * c]) * inv, __uint_as_float(r[8 * c + 1]) * inv, p.fp16),
                 pack_act(__uint_as_float(r[8 * c + 2]) * inv, __uint_as_float(r[8 * c + 3]) * inv, p.fp16),
                 pack_act(__uint_as_float(r[8 * c + 4]) * inv, __uint_as_float(r[8 * c + 5]) * inv, p.fp16),
                 pack_act(__uint_as_float(r[8 * c + 6]) * inv, __uint_as_float(r[8 * c + 7]) * inv, p.fp16));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmO, smem_u32(sO), h * 64, b * p.q_batch_rows);
          bulk_commit();
        }
      }
      // the row sums of this item are read by the quarter-0 warp before the next item's sums are written:
      // it reaches the next bar.sync (row max exchange) only after its epilogue
    }
    if (q == 0 && lane == 0) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int launch_qf_cross(const AttnDesc& a, cudaStream_t st) {
  CUtensorMap tmQ, tmK, tmV, tmO;
  const uint64_t w = (uint64_t)a.H * 64;
  const uint64_t qrows = (uint64_t)(a.B - 1) * a.q_batch_rows + a.Lq;
  const uint64_t krows = (uint64_t)(a.B - 1) * a.kv_batch_rows + a.Lk;
  SPRC_TRY(make_tmap_bf16(&tmQ, a.Q, w, qrows, 1, a.ldq, 0, 64, 32, 1, 2));
  const uint64_t hstride = a.kv_head_stride > 0 ? (uint64_t)a.kv_head_stride : 64;
  SPRC_TRY(make_tmap_bf16(&tmK, a.K, 64, krows, a.H, a.ldk, hstride, 64, QC_HALF, 1, 3));
  SPRC_TRY(make_tmap_bf16(&tmV, a.V, 64, krows, a.H, a.ldv, hstride, 64, QC_HALF, 1, 3));
  SPRC_TRY(make_tmap_bf16(&tmO, a.O, w, qrows, 1, a.ldo, 0, 64, 32, 1, 2));
  QfCrossParams p;
  p.B = a.B;
  p.H = a.H;
  p.Lk = a.Lk;
  p.q_batch_rows = a.q_batch_rows;
  p.kv_batch_rows = a.kv_batch_rows;
  p.scale_log2 = a.scale * kLog2e;
  p.fp16 = act_fp16();
  p.rev = next_sweep_reverse();
  static bool attr_set = false;
  if (!attr_set) {
    SPRC_CUDA(cudaFuncSetAttribute(qf_cross_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   QC_SMEM));
    attr_set = true;
  }
  const int items = a.B * a.H;
  const int grid = items < device_sm_count() ? items : device_sm_count();
  prof_begin(st);
  SPRC_CUDA(launch_pdl(qf_cross_attention_tc_kernel, dim3(grid), dim3(QF_THREADS), QC_SMEM, st, tmQ, tmK, tmV, tmO,
                       p));
  if (prof_enabled()) {
    char tag[56];
    snprintf(tag, sizeof(tag), "qf-cross tc B%d H%d Lq%d Lk%d", a.B, a.H, a.Lq, a.Lk);
    prof_end(PROF_ATTN, 4.0 * a.B * a.H * (double)a.Lq * a.Lk * 64, 2.0 * a.B * a.H * 64 * (2.0 * a.Lq + 2.0 * a.Lk), st,
             tag);
  }
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

// Eligible: the Q-Former's own shapes (everything else stays on attention_small / attention_kernel).
bool attention_qf_eligible(const AttnDesc& a) {
  if (a.dh != 64 || a.kv_idx0 || a.kv_idx1) return false;
  const bool self = a.Lq == a.Lk && (a.Lq == 32 || a.Lq == 64) && a.q_batch_rows == a.Lq && a.kv_batch_rows == a.Lk;
  const bool cross = a.Lq == 32 && a.Lk > 64 && a.Lk <= QC_LK && !a.key_mask && a.q_batch_rows >= 32 &&
                     a.kv_batch_rows >= a.Lk;
  return self || cross;
}

int attention_qf(const AttnDesc& a, cudaStream_t st) {
  if (a.Lq == a.Lk) return a.Lq == 64 ? launch_qf_self<64>(a, st) : launch_qf_self<32>(a, st);
  return launch_qf_cross(a, st);
}

}  // namespace sprc
