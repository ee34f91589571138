// ViT multi-head self-attention on tcgen05 (sm_100a): softmax(Q K^T * scale) V with L = 257 tokens,
// 16 heads x 64 (CLIP-L, clip_vit.py:134 nn.MultiheadAttention) or x 88 (EVA-g, eva_vit.py:118-148).
//
// Persistent kernel, one CTA per SM, one (image, head) at a time, 128 query rows per tile (3 tiles):
//   warp 0      TMA producer: K and V of the head (272 rows, zero / next-image rows masked later) once per
//               item, the Q tile per M-tile; 3-D tensor maps (d, token row, head) straight over the packed
//               [B*257, 3*Dv] QKV activation - a head dim of 88 is clipped by the map (columns >= 88 are
//               zero-filled by TMA), so no padding copies exist.
//   warp 1      tcgen05.mma issuer.  S[128 x 272] = Q K^T (K-major smem operands, fp32 in TMEM);
//               O[128 x dh] = P V with P read straight from TMEM (A operand in TMEM, bf16) and V as an
//               MN-major smem operand (its natural [token, d] layout - no transpose).
//   warps 2..17 softmax + epilogue: each thread owns one query row (TMEM lane) and a 64/80-key segment that it
//               reads from TMEM ONCE into registers; row max and row sum are exchanged between the four
//               segment warps of a row through shared memory; ex2.approx with the scale folded in; P written
//               back to TMEM as packed bf16 with tcgen05.st; O scaled by 1/l.  Warps whose 32 rows all lie
//               beyond L (most of the third tile, L = 257) skip the exponentials.
// The score matrix never leaves the SM (the reference materialises [B,16,257,257] in HBM, eva_vit.py:128-141).
// TMEM columns: O [0,96) | S [96,368) | P [368,504).
#include <math.h>
#include <stdio.h>

#include "ops.h"
#include "ptx.cuh"

namespace sprc {

int make_tmap_bf16(CUtensorMap* tm, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
                   uint64_t stride2, uint32_t b0, uint32_t b1, uint32_t b2, int rank);

static constexpr int TA_LK = 272;             // keys padded to a multiple of 16 (and of the 8-row swizzle atom)
static constexpr int TA_HALF = TA_LK / 2;     // 136 keys per softmax warp
static constexpr int TA_QBYTES = 128 * 128;   // one 64-column block of a 128-row Q tile
static constexpr int TA_KBYTES = TA_LK * 128; // one 64-column block of K or V
static constexpr int TA_COL_O = 0, TA_COL_S = 96, TA_COL_P = 368;
static constexpr int TA_SM_WARPS = 16;       // softmax warps: 4 per TMEM lane quarter
static constexpr int TA_THREADS = (2 + TA_SM_WARPS) * 32;

struct TcAttnParams {
  int B, H, L;        // images, heads, tokens (257)
  int ldo;            // output row pitch (elements)
  float scale_log2;   // scale * log2(e)
  int fp16;           // operand format
  int rev;            // sweep the items from the last one down (next_sweep_reverse)
  bf16* O;
};

template <int DH>
__global__ void __launch_bounds__(TA_THREADS, 1)
vit_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const TcAttnParams p) {
  constexpr int DHB = (DH + 63) / 64;         // 64-column blocks per row (1 or 2)
  constexpr int DHP = (DH + 15) / 16 * 16;    // head dim padded to the MMA K step (64 / 96)
  constexpr int KSTEPS = DHP / 16;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + DHB * TA_QBYTES;
  uint8_t* sV = sK + DHB * TA_KBYTES;
  uint8_t* tail = sV + DHB * TA_KBYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint64_t* kv_full = bars + 0;
  uint64_t* kv_empty = bars + 1;
  uint64_t* q_full = bars + 2;
  uint64_t* q_empty = bars + 3;
  uint64_t* s_full = bars + 4;
  uint64_t* p_full = bars + 5;
  uint64_t* o_full = bars + 6;
  uint64_t* o_empty = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const uint32_t xch = smem_u32(bars + 10);  // [4 segments][128 rows] row max, then [4][128] row sum (byte address)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.B * p.H;
  const int n_mt = (p.L + 127) / 128;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, TA_SM_WARPS);
    mbar_init(o_full, 1);
    mbar_init(o_empty, TA_SM_WARPS);
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
      uint32_t kv_ph = 0, q_ph = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int itm = p.rev ? n_items - 1 - item : item;
        const int b = itm / p.H, h = itm % p.H;
        const int row0 = b * p.L;
        mbar_wait(kv_empty, kv_ph ^ 1);
        mbar_expect_tx(kv_full, 2 * DHB * TA_KBYTES);
#pragma unroll
        for (int kb = 0; kb < DHB; ++kb) {
          tma_load_3d(&tmK, kv_full, sK + kb * TA_KBYTES, kb * 64, row0, h, kEvictNormal);
          tma_load_3d(&tmK, kv_full, sK + kb * TA_KBYTES + TA_HALF * 128, kb * 64, row0 + TA_HALF, h, kEvictNormal);
          tma_load_3d(&tmV, kv_full, sV + kb * TA_KBYTES, kb * 64, row0, h, kEvictNormal);
          tma_load_3d(&tmV, kv_full, sV + kb * TA_KBYTES + TA_HALF * 128, kb * 64, row0 + TA_HALF, h, kEvictNormal);
        }
        kv_ph ^= 1;
        for (int mt = 0; mt < n_mt; ++mt) {
          mbar_wait(q_empty, q_ph ^ 1);
          mbar_expect_tx(q_full, DHB * TA_QBYTES);
#pragma unroll
          for (int kb = 0; kb < DHB; ++kb)
            tma_load_3d(&tmQ, q_full, sQ + kb * TA_QBYTES, kb * 64, row0 + mt * 128, h, kEvictNormal);
          q_ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_s256 = umma_idesc_16(128, 256, p.fp16);
    const uint32_t idesc_s16 = umma_idesc_16(128, 16, p.fp16);
    const uint32_t idesc_pv = umma_idesc_16(128, DHP, p.fp16) | (1u << 16);  // B operand MN-major
    uint32_t kv_ph = 0, q_ph = 0, p_ph = 0, oe_ph = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      mbar_wait(kv_full, kv_ph);
      kv_ph ^= 1;
      for (int mt = 0; mt < n_mt; ++mt) {
        mbar_wait(q_full, q_ph);
        q_ph ^= 1;
        tc_fence_after();
        // S is free: the softmax warps signalled p_full for the previous tile before we issued its PV
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < KSTEPS; ++ks) {
            const int kb = ks >> 2, kk = ks & 3;
            const uint64_t da = umma_desc_k_sw128(smem_u32(sQ + kb * TA_QBYTES)) + 2 * kk;
            const uint64_t db = umma_desc_k_sw128(smem_u32(sK + kb * TA_KBYTES)) + 2 * kk;
            umma_bf16(tmem_base + TA_COL_S, da, db, idesc_s256, ks != 0 ? 1u : 0u);
            umma_bf16(tmem_base + TA_COL_S + 256, da, db + ((256 * 128) >> 4), idesc_s16, ks != 0 ? 1u : 0u);
          }
          umma_commit(q_empty);
          umma_commit(s_full);
        }
        __syncwarp();
        mbar_wait(p_full, p_ph);   // P(mt) is in TMEM, S(mt) consumed
        p_ph ^= 1;
        mbar_wait(o_empty, oe_ph ^ 1);  // epilogue of the previous tile has read O
        oe_ph ^= 1;
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < TA_LK / 16; ++ks) {
            const uint64_t db = umma_desc_mn_sw128(smem_u32(sV + ks * 16 * 128), TA_KBYTES);
            umma_bf16_ts(tmem_base + TA_COL_O, tmem_base + TA_COL_P + ks * 8, db, idesc_pv, ks != 0 ? 1u : 0u);
          }
          umma_commit(o_full);
          if (mt == n_mt - 1) umma_commit(kv_empty);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== softmax + epilogue (warps 2..17) =====================
    const int q = warp & 3;              // TMEM lane quarter
    const int seg = (warp - 2) >> 2;     // key segment: seg 0 -> keys [0,80), seg s>0 -> [80 + 64(s-1), +64)
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int key0 = seg == 0 ? 0 : 80 + 64 * (seg - 1);
    const int nkeys = seg == 0 ? 80 : 64;
    uint32_t s_ph = 0, o_ph = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int itm = p.rev ? n_items - 1 - item : item;
        const int b = itm / p.H, h = itm % p.H;
      for (int mt = 0; mt < n_mt; ++mt) {
        const bool warp_has_rows = mt * 128 + q * 32 < p.L;  // warp-uniform
        mbar_wait(s_full, s_ph);
        s_ph ^= 1;
        tc_fence_after();
        const uint32_t s_addr = tmem_base + lane_addr + TA_COL_S + key0;
        const uint32_t p_addr = tmem_base + lane_addr + TA_COL_P + key0 / 2;
        float sum = 0.f;
        if (warp_has_rows) {
          // ---- one TMEM read of this thread's segment ----
          uint32_t sr[80];
          {
            uint32_t(&a0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sr[0]);
            uint32_t(&a1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sr[32]);
            uint32_t(&a2)[16] = *reinterpret_cast<uint32_t(*)[16]>(&sr[64]);
            tmem_ld32(s_addr, a0);
            tmem_ld32(s_addr + 32, a1);
            if (seg == 0) tmem_ld16(s_addr + 64, a2);
            tmem_ld_wait();
          }
          float mx = -INFINITY;
          if (key0 + nkeys <= p.L) {
#pragma unroll
            for (int j = 0; j < 64; ++j) mx = fmaxf(mx, __uint_as_float(sr[j]));
            if (seg == 0) {
#pragma unroll
              for (int j = 64; j < 80; ++j) mx = fmaxf(mx, __uint_as_float(sr[j]));
            }
          } else {  // the last segment holds the padded keys >= L
#pragma unroll
            for (int j = 0; j < 64; ++j)
              if (key0 + j < p.L) mx = fmaxf(mx, __uint_as_float(sr[j]));
          }
          sts32f(xch + (seg * 128 + row_in_tile) * 4, mx);
          asm volatile("bar.sync 1, 512;" ::: "memory");
          mx = fmaxf(fmaxf(lds32f(xch + row_in_tile * 4), lds32f(xch + (128 + row_in_tile) * 4)),
                     fmaxf(lds32f(xch + (256 + row_in_tile) * 4), lds32f(xch + (384 + row_in_tile) * 4)));
          const float moff = mx * p.scale_log2;
          // ---- p = 2^(s*scale - max*scale), packed bf16 -> TMEM, 32 keys (16 columns) at a time ----
          const int nblk = nkeys / 32;  // 2 (+ a 16-key tail for segment 0)
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) {
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const int key = key0 + blk * 32 + j;
              float e0 = ex2_approx(fmaf(__uint_as_float(sr[blk * 32 + j]), p.scale_log2, -moff));
              float e1 = ex2_approx(fmaf(__uint_as_float(sr[blk * 32 + j + 1]), p.scale_log2, -moff));
              if (key >= p.L) e0 = 0.f;
              if (key + 1 >= p.L) e1 = 0.f;
              sum += e0 + e1;
              pk[j / 2] = pack_act(e0, e1, p.fp16);
            }
            tmem_st16(p_addr + blk * 16, pk);
          }
          (void)nblk;
          if (seg == 0) {
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float e0 = ex2_approx(fmaf(__uint_as_float(sr[64 + j]), p.scale_log2, -moff));
              const float e1 = ex2_approx(fmaf(__uint_as_float(sr[64 + j + 1]), p.scale_log2, -moff));
              sum += e0 + e1;
              pk[j / 2] = pack_act(e0, e1, p.fp16);
            }
            tmem_st8(p_addr + 32, pk);
          }
          tmem_st_wait();
        } else {
          // rows beyond L: P content is irrelevant (rows are never stored) - only keep the barriers in step
          asm volatile("bar.sync 1, 512;" ::: "memory");
        }
        sts32f(xch + (512 + seg * 128 + row_in_tile) * 4, sum);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        // ---- epilogue: O / l ; the four segment warps of a row each store a quarter of the columns ----
        mbar_wait(o_full, o_ph);
        o_ph ^= 1;
        tc_fence_after();
        asm volatile("bar.sync 1, 512;" ::: "memory");  // all row sums are visible
        const int qrow = mt * 128 + row_in_tile;
        constexpr int OC = DHP / 4;  // output columns per warp (16 or 24)
        if (warp_has_rows) {
          const float inv = 1.0f / ((lds32f(xch + (512 + row_in_tile) * 4) + lds32f(xch + (640 + row_in_tile) * 4)) +
                                    (lds32f(xch + (768 + row_in_tile) * 4) + lds32f(xch + (896 + row_in_tile) * 4)));
          const uint32_t o_addr = tmem_base + lane_addr + TA_COL_O + seg * OC;
          bf16* orow = p.O + (static_cast<size_t>(b) * p.L + qrow) * p.ldo + h * DH + seg * OC;
          uint32_t r[16];
          tmem_ld16(o_addr, r);
          uint32_t r2[8];
          if constexpr (OC == 24) tmem_ld8(o_addr + 16, r2);
          tmem_ld_wait();
          if (qrow < p.L) {
#pragma unroll
            for (int j = 0; j < 16; j += 8)
              *reinterpret_cast<uint4*>(orow + j) = make_uint4(
                  pack_act(__uint_as_float(r[j]) * inv, __uint_as_float(r[j + 1]) * inv, p.fp16),
                  pack_act(__uint_as_float(r[j + 2]) * inv, __uint_as_float(r[j + 3]) * inv, p.fp16),
                  pack_act(__uint_as_float(r[j + 4]) * inv, __uint_as_float(r[j + 5]) * inv, p.fp16),
                  pack_act(__uint_as_float(r[j + 6]) * inv, __uint_as_float(r[j + 7]) * inv, p.fp16));
            if constexpr (OC == 24) {
              if (seg * OC + 16 < DH)
                *reinterpret_cast<uint4*>(orow + 16) = make_uint4(
                    pack_act(__uint_as_float(r2[0]) * inv, __uint_as_float(r2[1]) * inv, p.fp16),
                    pack_act(__uint_as_float(r2[2]) * inv, __uint_as_float(r2[3]) * inv, p.fp16),
                    pack_act(__uint_as_float(r2[4]) * inv, __uint_as_float(r2[5]) * inv, p.fp16),
                    pack_act(__uint_as_float(r2[6]) * inv, __uint_as_float(r2[7]) * inv, p.fp16));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int DH>
static int launch_tc(const AttnDesc& a, cudaStream_t st) {
  constexpr int DHB = (DH + 63) / 64;
  const size_t smem = (size_t)DHB * (TA_QBYTES + 2 * TA_KBYTES) + 10 * 8 + 8 * 128 * 4 + 1024;
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t rows = (uint64_t)a.B * a.Lq;
  SPRC_TRY(make_tmap_bf16(&tmQ, a.Q, DH, rows, a.H, a.ldq, DH, 64, 128, 1, 3));
  SPRC_TRY(make_tmap_bf16(&tmK, a.K, DH, rows, a.H, a.ldk, DH, 64, TA_HALF, 1, 3));
  SPRC_TRY(make_tmap_bf16(&tmV, a.V, DH, rows, a.H, a.ldv, DH, 64, TA_HALF, 1, 3));
  TcAttnParams p;
  p.B = a.B;
  p.H = a.H;
  p.L = a.Lq;
  p.ldo = a.ldo;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.fp16 = act_fp16();
  p.rev = next_sweep_reverse();
  p.O = a.O;
  static bool attr_set = false;
  if (!attr_set) {
    SPRC_CUDA(cudaFuncSetAttribute(vit_attention_tc_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    attr_set = true;
  }
  const int items = a.B * a.H;
  const int grid = items < device_sm_count() ? items : device_sm_count();
  prof_begin(st);
  SPRC_CUDA(launch_pdl(vit_attention_tc_kernel<DH>, dim3(grid), dim3(TA_THREADS), smem, st, tmQ, tmK, tmV, p));
  if (prof_enabled()) {
    char tag[56];
    snprintf(tag, sizeof(tag), "tc B%d H%d dh%d L%d", a.B, a.H, a.dh, a.Lq);
    prof_end(PROF_ATTN, 4.0 * a.B * a.H * (double)a.Lq * a.Lk * a.dh, 2.0 * a.B * a.H * a.dh * 4.0 * a.Lq, st, tag);
  }
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

// Eligible: plain self-attention over packed per-image rows (ViT), L <= 272, dh 64 or 88.
bool attention_tc_eligible(const AttnDesc& a) {
  return (a.dh == 64 || a.dh == 88) && a.Lq == a.Lk && a.Lq > 128 && a.Lq <= TA_LK && !a.key_mask && !a.kv_idx0 &&
         a.q_batch_rows == a.Lq && a.kv_batch_rows == a.Lk && a.ldq == a.ldk && a.ldk == a.ldv;
}

int attention_tc(const AttnDesc& a, cudaStream_t st) {
  if (a.dh == 64) return launch_tc<64>(a, st);
  return launch_tc<88>(a, st);
}

}  // namespace sprc
