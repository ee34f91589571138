// Q-Former self-attention over RAGGED text rows (sm_100a tcgen05), dh = 64, 12 heads.
//
// The reference pads every caption to 32 tokens and runs all 64 rows of every sample through both Q-Former passes
// (Qformer.py:211-256 with the additive pad mask of :807); padded text rows are never attended to (their softmax
// weight is exp(-10000) = 0 in fp32) and their own outputs are never read (align_prompt.py:343 keeps the query rows,
// :348 row 32), so they are dead work: 19 of 64 rows per sample for CIRR-length captions.  The ragged layout
// drops them.  Hidden-state rows of a B-sample batch:
//     rows [0, 32 B)                 the 32 query rows of sample 0, 1, ...
//     rows [32 B, 32 B + T8)         live text rows, sample b at toff[b] .. toff[b] + L[b]; the two samples of a
//                                    pair are adjacent and each PAIR owns a slot of round_up(L0 + L1, 8) rows (the
//                                    <= 7 slack rows at its end hold finite don't-care values)
// so every GEMM / LayerNorm of the Q-Former becomes a plain dense row range, and only this kernel has to know
// which rows belong together.  One work item = (pair of samples, head): a 128-row tile of four 32-row quarters =
// query rows of sample 0 | query rows of sample 1 | 32 rows from sample 0's first text row | 32 rows from sample 1's,
// so that each of the four softmax warps (= TMEM lane quarters) serves ONE sample (tcgen05.ld/.st take one address per
// warp).  As in attention_qf.cu ONE 128x128x64 MMA forms all scores, each softmax thread owns one row and reads only
// the two 32-column blocks of its own sample (32 query keys + L live text keys), P is written to TMEM as a sparse
// bf16 matrix (the other sample's blocks stay zero) and O = P V is a TMEM-A MMA with V MN-major.  Text rows past a
// sample's rows belong to other samples: they are loaded and computed (finite) but never stored - the epilogue
// compacts the two text quarters into the pair's slot order in the staging tile and stores the slot in 8-row boxes.
#include <math.h>
#include <stdio.h>

#include "ops.h"
#include "ptx.cuh"

namespace sprc {

int make_tmap_bf16(CUtensorMap* tm, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
                   uint64_t stride2, uint32_t b0, uint32_t b1, uint32_t b2, int rank);

namespace {

constexpr float kLog2e = 1.4426950408889634f;
constexpr int THREADS = 6 * 32;
constexpr int TILE = 128 * 128;        // 128 rows x 64 x 16-bit
constexpr int QUARTER = 32 * 128;   // bytes of a 32-row quarter of a tile
constexpr int STAGE = 3 * TILE;        // Q, K, V
constexpr int COL_S = 0, COL_P = 128, COL_O = 192;   // TMEM columns (256 allocated)
constexpr int SMEM = 2 * STAGE + 256 + 1024;

struct RaggedParams {
  int n_pairs, H;
  int text_base;          // first text row = 32 * B
  const int4* pairs;      // per pair: {toff0, L0, toff1, L1} (text row offsets relative to text_base; L1 = 0: no sample 1)
  float scale_log2;
  int fp16;
  int rev;
};

__global__ void __launch_bounds__(THREADS, 2)
qf_self_attention_ragged_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO32,
                                const __grid_constant__ CUtensorMap tmO8, const RaggedParams p) {
  // tmQ / tmK / tmV / tmO32: 32-row boxes; tmO8: 8-row boxes
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * STAGE);
  uint64_t* full = bars;        // [2]
  uint64_t* empty = bars + 2;   // [2]
  uint64_t* s_full = bars + 4;
  uint64_t* p_full = bars + 5;
  uint64_t* o_full = bars + 6;
  uint64_t* o_empty = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.n_pairs * p.H;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO32);
    tma_prefetch_desc(&tmO8);
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
        const int4 pr = __ldg(&p.pairs[g]);
        const int r[4] = {g * 64, g * 64 + 32, p.text_base + pr.x, p.text_base + pr.z};
        uint8_t* st = smem + s * STAGE;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], STAGE);
        // (rows past the end of the activation are zero-filled)
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          tma_load_2d(&tmQ, &full[s], st + qd * QUARTER, h * 64, r[qd], kEvictFirst);
          tma_load_2d(&tmK, &full[s], st + TILE + qd * QUARTER, h * 64, r[qd], kEvictFirst);
          tma_load_2d(&tmV, &full[s], st + 2 * TILE + qd * QUARTER, h * 64, r[qd], kEvictFirst);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_s = umma_idesc_16(128, 128, p.fp16);
    const uint32_t idesc_pv = umma_idesc_16(128, 64, p.fp16) | (1u << 16);  // B operand MN-major
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int s = it & 1;
      uint8_t* st = smem + s * STAGE;
      mbar_wait(&full[s], (it >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = umma_desc_k_sw128(smem_u32(st));
        const uint64_t db = umma_desc_k_sw128(smem_u32(st + TILE));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16(tmem_base + COL_S, da + 2 * kk, db + 2 * kk, idesc_s, kk != 0 ? 1u : 0u);
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, it & 1);          // P(it) is in TMEM, S(it) consumed
      mbar_wait(o_empty, (it & 1) ^ 1);   // the epilogue of the previous item has read O
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t db = umma_desc_mn_sw128(smem_u32(st + 2 * TILE + ks * 16 * 128), TILE);
          umma_bf16_ts(tmem_base + COL_O, tmem_base + COL_P + ks * 8, db, idesc_pv, ks != 0 ? 1u : 0u);
        }
        umma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax + epilogue (warps 2..5): one thread per tile row =====================
    const int q = warp & 3;                 // TMEM lane quarter = tile quarter: 0/1 query rows, 2/3 text rows
    const int row = q * 32 + lane;
    const int smp = q & 1;                  // the sample (of the pair) this warp serves
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + COL_S;
    const uint32_t p_addr = tmem_base + lane_addr + COL_P;
    {
      // P row = 64 packed columns (128 keys): this row only ever writes [16 smp, +16) (query keys) and
      // [32 + 16 smp, +16) (text keys); the other sample's blocks stay zero for the whole kernel
      uint32_t z[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) z[j] = 0u;
      tmem_st16(p_addr + (smp ^ 1) * 16, z);
      tmem_st16(p_addr + 32 + (smp ^ 1) * 16, z);
      tmem_st_wait();
    }
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int s = it & 1;
      const int itm = p.rev ? n_items - 1 - item : item;
      const int g = itm / p.H, h = itm % p.H;
      const int4 pr = __ldg(&p.pairs[g]);             // {toff0, L0, toff1, L1}
      const int L = smp ? pr.w : pr.y;                // live text keys of this row's sample
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      uint32_t sq[32], stx[32];
      tmem_ld32(s_addr + smp * 32, sq);               // scores against the sample's query rows
      tmem_ld32(s_addr + 64 + smp * 32, stx);         // ... against its text quarter (first L columns live)
      tmem_ld_wait();
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        mx = fmaxf(mx, __uint_as_float(sq[j]));
        if (j < L) mx = fmaxf(mx, __uint_as_float(stx[j]));
      }
      mx *= p.scale_log2;   // scale > 0
      float sum = 0.f;
      uint32_t pq[16], pt[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float e0 = ex2_approx(fmaf(__uint_as_float(sq[j]), p.scale_log2, -mx));
        const float e1 = ex2_approx(fmaf(__uint_as_float(sq[j + 1]), p.scale_log2, -mx));
        sum += e0 + e1;
        pq[j / 2] = pack_act(e0, e1, p.fp16);
      }
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float f0 = j < L ? ex2_approx(fmaf(__uint_as_float(stx[j]), p.scale_log2, -mx)) : 0.f;
        const float f1 = j + 1 < L ? ex2_approx(fmaf(__uint_as_float(stx[j + 1]), p.scale_log2, -mx)) : 0.f;
        sum += f0 + f1;
        pt[j / 2] = pack_act(f0, f1, p.fp16);
      }
      tmem_st16(p_addr + smp * 16, pq);
      tmem_st16(p_addr + 32 + smp * 16, pt);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // ---- epilogue: O / l -> 16-bit -> swizzled staging (the Q tile of this stage, dead since S was formed) ----
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      uint32_t r[64];
      {
        uint32_t(&a0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
        uint32_t(&a1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[32]);
        tmem_ld32(tmem_base + lane_addr + COL_O, a0);
        tmem_ld32(tmem_base + lane_addr + COL_O + 32, a1);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      const float inv = 1.0f / sum;
      // staging row: query quarters in place; text rows in slot order = sample 0's L0 live rows, then sample 1's rows
      // (its L1 live rows followed by whatever fills the slot up to a multiple of 8)
      const int slot8 = (pr.y + pr.w + 7) & ~7;
      int srow = row;
      bool wr = true;
      if (q == 2) wr = lane < pr.y;
      if (q == 3) {
        srow = 64 + pr.y + lane;
        wr = pr.y + lane < slot8;
      }
      const uint32_t stg = smem_u32(smem + s * STAGE) + srow * 128;
      if (wr) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          sts128(stg + ((c ^ (srow & 7)) << 4),
                 pack_act(__uint_as_float(r[8 * c]) * inv, __uint_as_float(r[8 * c + 1]) * inv, p.fp16),
                 pack_act(__uint_as_float(r[8 * c + 2]) * inv, __uint_as_float(r[8 * c + 3]) * inv, p.fp16),
                 pack_act(__uint_as_float(r[8 * c + 4]) * inv, __uint_as_float(r[8 * c + 5]) * inv, p.fp16),
                 pack_act(__uint_as_float(r[8 * c + 6]) * inv, __uint_as_float(r[8 * c + 7]) * inv, p.fp16));
      }
      fence_proxy_async();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == 2 && lane == 0) {
        const uint32_t base = smem_u32(smem + s * STAGE);
        // query rows of sample 0 and (if it exists) sample 1: 32-row boxes
        tma_store_2d(&tmO32, base, h * 64, g * 64);
        if (pr.w > 0) tma_store_2d(&tmO32, base + QUARTER, h * 64, g * 64 + 32);
        // the pair's text slot, 8 rows at a time (nothing beyond it: those rows belong to other pairs)
        for (int i = 0; i < (slot8 >> 3); ++i)
          tma_store_2d(&tmO8, base + 2 * QUARTER + i * 8 * 128, h * 64, p.text_base + pr.x + 8 * i);
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

}  // namespace

// qkv: packed [rows, 3 * 768] activation (Q | K | V, head-major columns); out: [rows, 768]; rows = 32 B + T8.
// pairs_dev: [ceil(B / 2)] int4 {toff0, L0, toff1, L1 (0 if the pair has one sample)}.
int attention_qf_ragged(const bf16* qkv, int ldqkv, bf16* out, int ldo, int B, int rows_total, const int4* pairs_dev,
                        float scale, cudaStream_t st) {
  SPRC_REQUIRE(B > 0 && rows_total >= 32 * B, "ragged attention: B=%d rows=%d", B, rows_total);
  const int H = 12;
  CUtensorMap tmQ, tmK, tmV, tmO32, tmO8;
  const uint64_t w = (uint64_t)H * 64, rows = (uint64_t)rows_total;
  SPRC_TRY(make_tmap_bf16(&tmQ, qkv, w, rows, 1, ldqkv, 0, 64, 32, 1, 2));
  SPRC_TRY(make_tmap_bf16(&tmK, qkv + 768, w, rows, 1, ldqkv, 0, 64, 32, 1, 2));
  SPRC_TRY(make_tmap_bf16(&tmV, qkv + 1536, w, rows, 1, ldqkv, 0, 64, 32, 1, 2));
  SPRC_TRY(make_tmap_bf16(&tmO32, out, w, rows, 1, ldo, 0, 64, 32, 1, 2));
  SPRC_TRY(make_tmap_bf16(&tmO8, out, w, rows, 1, ldo, 0, 64, 8, 1, 2));
  RaggedParams p;
  p.n_pairs = (B + 1) / 2;
  p.H = H;
  p.text_base = 32 * B;
  p.pairs = pairs_dev;
  p.scale_log2 = scale * kLog2e;
  p.fp16 = act_fp16();
  p.rev = next_sweep_reverse();
  static bool attr_set = false;
  if (!attr_set) {
    SPRC_CUDA(cudaFuncSetAttribute(qf_self_attention_ragged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  const int items = p.n_pairs * H;
  int grid = 2 * device_sm_count();
  if (grid > items) grid = items;
  prof_begin(st);
  SPRC_CUDA(launch_pdl(qf_self_attention_ragged_kernel, dim3(grid), dim3(THREADS), SMEM, st, tmQ, tmK, tmV, tmO32,
                       tmO8, p));
  if (prof_enabled()) {
    char tag[56];
    snprintf(tag, sizeof(tag), "qf-self ragged B%d rows%d", B, rows_total);
    prof_end(PROF_ATTN, 4.0 * p.n_pairs * H * 128.0 * 128 * 64 / 2, 2.0 * rows_total * 768 * 4.0, st, tag);
  }
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sprc
