// ViT multi-head self-attention on tcgen05 (sm_100a), second generation: softmax(Q K^T * scale) V with L = 257 tokens,
// 16 heads x 64 (CLIP-L, clip_vit.py:134 nn.MultiheadAttention) or x 88 (EVA-g, eva_vit.py:118-148).
//
// 257 = 2 * 128 + 1 on BOTH axes is what made the first kernel (attention_tc.cu) slow: a third 128-row tile for one
// query row, a 272-column score tile that forces all 512 TMEM columns into one buffer, and therefore a strictly
// serial TMA -> MMA -> softmax -> MMA -> store chain per tile.  Here the "+1" never reaches the tensor core:
//   * queries 0..255 are TWO 128-row tiles (A, B) that are in flight together; query 256 is computed on CUDA cores by
//     two tail warps straight from the K / V tiles in shared memory (257 x dh MACs twice);
//   * keys 0..255 are ONE N = 256 MMA per tile; the score of key 256 is a dh-long dot product per query row, split
//     over the four softmax threads of a row (operands read through L2), its probability enters the row sum and its
//     value row enters O as a rank-1 update in the epilogue.
// TMEM (512 columns) = two regions of 256, one per tile.  Inside a region: S fp32 [0,256); once every thread has
// read its scores, P (16-bit, packed) overwrites columns [0,64) (keys 0..127) and [192,256) (keys 128..255) and O
// accumulates in [64, 64 + dh).  So while the 16 softmax warps work on tile A the tensor core computes S of tile B,
// P V of tile A runs under the softmax of tile B, and the next item's S under this item's epilogues.
// Shared memory is single-buffered but released early: K after the second S MMA, V after the second P V MMA, each Q
// tile after its S MMA - the next item's operands stream in behind the MMAs that read the current ones.
//   warp 0       TMA producer        warp 1   tcgen05.mma issuer
//   warps 2..17  softmax + epilogue: thread = (query row, 64-key segment), as attention_tc.cu
//   warps 18,19  tail: query row 256
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "mma_sync.cuh"
#include "ops.h"
#include "ptx.cuh"

namespace sprc {

int make_tmap_bf16(CUtensorMap* tm, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
                   uint64_t stride2, uint32_t b0, uint32_t b1, uint32_t b2, int rank);

namespace {

constexpr int VA_L = 257;
constexpr int VA_LK = 272;             // K / V rows staged per item (row 256 = the odd key; the rest is never read)
constexpr int VA_HALF = VA_LK / 2;     // TMA box rows
constexpr int VA_QBYTES = 128 * 128;   // one 64-column block of a 128-row Q tile
constexpr int VA_KBYTES = VA_LK * 128; // one 64-column block of K or V
constexpr int VA_SM_WARPS = 16;
constexpr int VA_TAIL_WARPS = 2;
constexpr int VA_THREADS = (2 + VA_SM_WARPS + VA_TAIL_WARPS) * 32;
constexpr int VA_REGION = 256;         // TMEM columns per tile
constexpr int VA_COL_O = 64;           // O inside a region (after the scores have been consumed)
constexpr int VA_COL_PHI = 192;        // P of keys 128..255 inside a region
constexpr int VA_XCH_FLOATS = 2 * 12 * 128;   // per tile: max[4][128], part[4][128], sum[4][128]
constexpr int VA_TAIL_FLOATS = 2 * 96 + 8;

struct VitAttnParams {
  int B, H;
  int ld;             // row pitch of the packed QKV activation (elements)
  int ldo;            // output row pitch (elements)
  float scale_log2;   // scale * log2(e)
  int fp16;
  int rev;
  const unsigned short* Q;   // = qkv; K = Q + Dv, V = Q + 2 Dv (16-bit elements)
  const unsigned short* K;
  const unsigned short* V;
  unsigned short* O;
};

__device__ __forceinline__ float2 unpack2(uint32_t w, int fp16) {
  if (fp16) {
    const __half2 h = *reinterpret_cast<const __half2*>(&w);
    return __half22float2(h);
  }
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
}
__device__ __forceinline__ float dot8(const uint4& a, const uint4& b, int fp16) {
  const float2 a0 = unpack2(a.x, fp16), a1 = unpack2(a.y, fp16), a2 = unpack2(a.z, fp16), a3 = unpack2(a.w, fp16);
  const float2 b0 = unpack2(b.x, fp16), b1 = unpack2(b.y, fp16), b2 = unpack2(b.z, fp16), b3 = unpack2(b.w, fp16);
  float s = a0.x * b0.x;
  s = fmaf(a0.y, b0.y, s);
  s = fmaf(a1.x, b1.x, s);
  s = fmaf(a1.y, b1.y, s);
  s = fmaf(a2.x, b2.x, s);
  s = fmaf(a2.y, b2.y, s);
  s = fmaf(a3.x, b3.x, s);
  s = fmaf(a3.y, b3.y, s);
  return s;
}
__device__ __forceinline__ uint32_t lds32u(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

template <int DH>
__global__ void __launch_bounds__(VA_THREADS, 1)
vit_attention_v2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const VitAttnParams p) {
  constexpr int DHB = (DH + 63) / 64;         // 64-column blocks per row (1 or 2)
  constexpr int DHP = (DH + 15) / 16 * 16;    // head dim padded to the MMA K step (64 / 96)
  constexpr int KSTEPS = DHP / 16;
  constexpr int NCH = DH / 8;                 // 16-byte chunks per head row (8 / 11)
  constexpr int CH = DHP / 32;                // chunks of the key-256 dot product per softmax thread (2 / 3)
  constexpr int OC = DHP / 4;                 // output columns per softmax thread (16 / 24)
  constexpr int NBUF = DH <= 64 ? 2 : 1;      // operand buffers: the whole NEXT item streams in behind this one
  constexpr int ITEM_BYTES = DHB * (2 * VA_QBYTES + 2 * VA_KBYTES);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // per buffer: Q [2 tiles][DHB blocks] | K [DHB] | V [DHB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NBUF * ITEM_BYTES);
  uint64_t* k_full = bars + 0;    // [2]
  uint64_t* k_empty = bars + 2;   // [2]
  uint64_t* v_full = bars + 4;    // [2]
  uint64_t* v_empty = bars + 6;   // [2]
  uint64_t* q_full = bars + 8;    // [2 buffers][2 tiles]
  uint64_t* q_empty = bars + 12;  // [2][2]
  uint64_t* s_full = bars + 16;   // [2 tiles]
  uint64_t* p_full = bars + 18;
  uint64_t* o_full = bars + 20;
  uint64_t* o_empty = bars + 22;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  const uint32_t xch = smem_u32(bars + 26);                       // VA_XCH_FLOATS floats
  const uint32_t tsm = xch + VA_XCH_FLOATS * 4;                   // tail warps: partial O[2][96], max[2], sum[2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.B * p.H;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int bf = 0; bf < 2; ++bf) {
      mbar_init(&k_full[bf], 1);
      mbar_init(&k_empty[bf], 2 + VA_SM_WARPS);   // MMA commit + tail warps + every softmax warp (key 256 row)
      mbar_init(&v_full[bf], 1);
      mbar_init(&v_empty[bf], 2 + VA_SM_WARPS);   // MMA commit + tail warps + every softmax warp (value 256 row)
      for (int t = 0; t < 2; ++t) {
        mbar_init(&q_full[bf * 2 + t], 1);
        mbar_init(&q_empty[bf * 2 + t], 1 + VA_SM_WARPS);   // MMA commit + every softmax warp (its query rows)
      }
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], VA_SM_WARPS);
      mbar_init(&o_full[t], 1);
      mbar_init(&o_empty[t], VA_SM_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();
  griddep_launch();

  // buffer of item iteration `it`, and the parity of its n-th use
#define VA_BUF(it) (NBUF == 2 ? ((it) & 1) : 0)
#define VA_USE(it) (NBUF == 2 ? (((it) >> 1) & 1) : ((it) & 1))

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int itm = p.rev ? n_items - 1 - item : item;
        const int b = itm / p.H, h = itm % p.H;
        const int row0 = b * VA_L;
        const int bf = VA_BUF(it);
        const uint32_t ph = VA_USE(it) ^ 1;
        uint8_t* sQ = smem + bf * ITEM_BYTES;
        uint8_t* sK = sQ + 2 * DHB * VA_QBYTES;
        uint8_t* sV = sK + DHB * VA_KBYTES;
        mbar_wait(&k_empty[bf], ph);
        mbar_expect_tx(&k_full[bf], DHB * VA_KBYTES);
#pragma unroll
        for (int kb = 0; kb < DHB; ++kb) {
          tma_load_3d(&tmK, &k_full[bf], sK + kb * VA_KBYTES, kb * 64, row0, h, kEvictNormal);
          tma_load_3d(&tmK, &k_full[bf], sK + kb * VA_KBYTES + VA_HALF * 128, kb * 64, row0 + VA_HALF, h, kEvictNormal);
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&q_empty[bf * 2 + t], ph);
          mbar_expect_tx(&q_full[bf * 2 + t], DHB * VA_QBYTES);
#pragma unroll
          for (int kb = 0; kb < DHB; ++kb)
            tma_load_3d(&tmQ, &q_full[bf * 2 + t], sQ + (t * DHB + kb) * VA_QBYTES, kb * 64, row0 + t * 128, h,
                        kEvictNormal);
        }
        mbar_wait(&v_empty[bf], ph);
        mbar_expect_tx(&v_full[bf], DHB * VA_KBYTES);
#pragma unroll
        for (int kb = 0; kb < DHB; ++kb) {
          tma_load_3d(&tmV, &v_full[bf], sV + kb * VA_KBYTES, kb * 64, row0, h, kEvictNormal);
          tma_load_3d(&tmV, &v_full[bf], sV + kb * VA_KBYTES + VA_HALF * 128, kb * 64, row0 + VA_HALF, h, kEvictNormal);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_s = umma_idesc_16(128, 256, p.fp16);
    const uint32_t idesc_pv = umma_idesc_16(128, DHP, p.fp16) | (1u << 16);  // B operand MN-major
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int bf = VA_BUF(it);
      const uint32_t bph = VA_USE(it);
      uint8_t* sQ = smem + bf * ITEM_BYTES;
      uint8_t* sK = sQ + 2 * DHB * VA_QBYTES;
      uint8_t* sV = sK + DHB * VA_KBYTES;
      mbar_wait(&k_full[bf], bph);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&q_full[bf * 2 + t], bph);
        mbar_wait(&o_empty[t], ph ^ 1);   // the previous item's epilogue has read this region's O
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < KSTEPS; ++ks) {
            const int kb = ks >> 2, kk = ks & 3;
            const uint64_t da = umma_desc_k_sw128(smem_u32(sQ + (t * DHB + kb) * VA_QBYTES)) + 2 * kk;
            const uint64_t db = umma_desc_k_sw128(smem_u32(sK + kb * VA_KBYTES)) + 2 * kk;
            umma_bf16(tmem_base + t * VA_REGION, da, db, idesc_s, ks != 0 ? 1u : 0u);
          }
          umma_commit(&q_empty[bf * 2 + t]);
          umma_commit(&s_full[t]);
          if (t == 1) umma_commit(&k_empty[bf]);
        }
        __syncwarp();
      }
      mbar_wait(&v_full[bf], bph);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&p_full[t], ph);   // P of this tile is in TMEM, its scores are consumed
        tc_fence_after();
        if (elect_one()) {
          const uint32_t reg = tmem_base + t * VA_REGION;
#pragma unroll
          for (int ks = 0; ks < 16; ++ks) {
            const uint32_t pa = reg + (ks < 8 ? ks * 8 : VA_COL_PHI + (ks - 8) * 8);
            const uint64_t db = umma_desc_mn_sw128(smem_u32(sV + ks * 16 * 128), VA_KBYTES);
            umma_bf16_ts(reg + VA_COL_O, pa, db, idesc_pv, ks != 0 ? 1u : 0u);
          }
          umma_commit(&o_full[t]);
          if (t == 1) umma_commit(&v_empty[bf]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 2 + VA_SM_WARPS) {
    // ===================== softmax + epilogue (warps 2..17) =====================
    const int q = warp & 3;              // TMEM lane quarter
    const int seg = (warp - 2) >> 2;     // 64-key segment
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t p_col = seg < 2 ? seg * 32 : VA_COL_PHI + (seg - 2) * 32;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int itm = p.rev ? n_items - 1 - item : item;
      const int b = itm / p.H, h = itm % p.H;
      const size_t row0 = static_cast<size_t>(b) * VA_L;
      const uint32_t ph = it & 1;
      const int bf = VA_BUF(it);
      const uint32_t bph = VA_USE(it);
      const uint32_t sQa = smem_u32(smem + bf * ITEM_BYTES);
      const uint32_t sKa = sQa + 2 * DHB * VA_QBYTES;
      const uint32_t sVa = sKa + DHB * VA_KBYTES;
      // ---- this thread's share of the key-256 scores of its two query rows, from the staged tiles ----
      float part[2] = {0.f, 0.f};
      mbar_wait(&k_full[bf], bph);
      uint4 kc[CH];
#pragma unroll
      for (int cc = 0; cc < CH; ++cc) {
        const int c = seg * CH + cc;   // row 256 of a swizzled tile: 256 & 7 == 0, chunks in place
        kc[cc] = c < NCH ? lds128(sKa + (c >> 3) * VA_KBYTES + 256 * 128 + ((c & 7) << 4)) : make_uint4(0, 0, 0, 0);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&k_empty[bf]);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&q_full[bf * 2 + t], bph);
#pragma unroll
        for (int cc = 0; cc < CH; ++cc) {
          const int c = seg * CH + cc;
          if (c < NCH)
            part[t] += dot8(lds128(sQa + (t * DHB + (c >> 3)) * VA_QBYTES + row_in_tile * 128 +
                                   (((c & 7) ^ (row_in_tile & 7)) << 4)), kc[cc], p.fp16);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_empty[bf * 2 + t]);
      }
      float p256[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const uint32_t xt = xch + t * (12 * 128 * 4);
        mbar_wait(&s_full[t], ph);
        tc_fence_after();
        const uint32_t reg = tmem_base + lane_addr + t * VA_REGION;
        uint32_t sr[64];
        {
          uint32_t(&a0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sr[0]);
          uint32_t(&a1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sr[32]);
          tmem_ld32(reg + seg * 64, a0);
          tmem_ld32(reg + seg * 64 + 32, a1);
          tmem_ld_wait();
        }
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 64; ++j) mx = fmaxf(mx, __uint_as_float(sr[j]));
        sts32f(xt + (seg * 128 + row_in_tile) * 4, mx);
        sts32f(xt + ((4 + seg) * 128 + row_in_tile) * 4, part[t]);
        tc_fence_before();
        asm volatile("bar.sync 1, 512;" ::: "memory");   // every score of this tile has been read: P may overwrite S
        tc_fence_after();
        const float s256 = (lds32f(xt + (4 * 128 + row_in_tile) * 4) + lds32f(xt + (5 * 128 + row_in_tile) * 4)) +
                           (lds32f(xt + (6 * 128 + row_in_tile) * 4) + lds32f(xt + (7 * 128 + row_in_tile) * 4));
        mx = fmaxf(fmaxf(lds32f(xt + row_in_tile * 4), lds32f(xt + (128 + row_in_tile) * 4)),
                   fmaxf(lds32f(xt + (256 + row_in_tile) * 4), lds32f(xt + (384 + row_in_tile) * 4)));
        mx = fmaxf(mx, s256);
        const float moff = mx * p.scale_log2;
        float sum = 0.f;
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float e0 = ex2_approx(fmaf(__uint_as_float(sr[blk * 32 + j]), p.scale_log2, -moff));
            const float e1 = ex2_approx(fmaf(__uint_as_float(sr[blk * 32 + j + 1]), p.scale_log2, -moff));
            sum += e0 + e1;
            pk[j / 2] = pack_act(e0, e1, p.fp16);
          }
          tmem_st16(reg + p_col + blk * 16, pk);
        }
        p256[t] = ex2_approx(fmaf(s256, p.scale_log2, -moff));
        if (seg == 0) sum += p256[t];
        sts32f(xt + ((8 + seg) * 128 + row_in_tile) * 4, sum);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");     // the row sums of both tiles are visible
      // value row 256 (rank-1 update of O) from the staged V tile, then this warp is done with the tile
      mbar_wait(&v_full[bf], bph);
      uint4 vv[OC / 8];
#pragma unroll
      for (int c = 0; c < OC / 8; ++c) {
        const int col = seg * OC + c * 8;
        vv[c] = col < DH ? lds128(sVa + (col >> 6) * VA_KBYTES + 256 * 128 + (((col >> 3) & 7) << 4))
                         : make_uint4(0, 0, 0, 0);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&v_empty[bf]);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const uint32_t xt = xch + t * (12 * 128 * 4);
        const float inv = 1.0f / ((lds32f(xt + (8 * 128 + row_in_tile) * 4) + lds32f(xt + (9 * 128 + row_in_tile) * 4)) +
                                  (lds32f(xt + (10 * 128 + row_in_tile) * 4) + lds32f(xt + (11 * 128 + row_in_tile) * 4)));
        mbar_wait(&o_full[t], ph);
        tc_fence_after();
        uint32_t r[OC];
        {
          uint32_t(&a0)[16] = *reinterpret_cast<uint32_t(*)[16]>(&r[0]);
          tmem_ld16(tmem_base + lane_addr + t * VA_REGION + VA_COL_O + seg * OC, a0);
          if constexpr (OC == 24) {
            uint32_t(&a1)[8] = *reinterpret_cast<uint32_t(*)[8]>(&r[16]);
            tmem_ld8(tmem_base + lane_addr + t * VA_REGION + VA_COL_O + seg * OC + 16, a1);
          }
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[t]);   // O is in registers: the region may take the next item's scores
        unsigned short* orow = p.O + (row0 + t * 128 + row_in_tile) * p.ldo + h * DH + seg * OC;
        const float pw = p256[t];
#pragma unroll
        for (int c = 0; c < OC / 8; ++c) {
          if (seg * OC + c * 8 < DH) {
            const uint32_t vw[4] = {vv[c].x, vv[c].y, vv[c].z, vv[c].w};
            uint32_t o4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 v2 = unpack2(vw[j], p.fp16);
              o4[j] = pack_act(fmaf(pw, v2.x, __uint_as_float(r[c * 8 + 2 * j])) * inv,
                               fmaf(pw, v2.y, __uint_as_float(r[c * 8 + 2 * j + 1])) * inv, p.fp16);
            }
            *reinterpret_cast<uint4*>(orow + c * 8) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
          }
        }
      }
    }
  } else {
    // ===================== tail warps: query row 256 on warp-level MMAs (m16n8k16, row 0 of the A tile) =====================
    constexpr int DT = DHP / 8;                    // output n-tiles
    const int tw = warp - (2 + VA_SM_WARPS);       // 0 / 1: keys [0,136) / [136,272)
    const int kbase = tw * VA_HALF;
    const uint32_t t_o = tsm;                      // float partial O [2][96]
    const uint32_t t_x = t_o + 2 * 96 * 4;         // float max[2], sum[2]
    const bool arow = lane < 4;                    // lanes holding row 0 of the fragments
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int itm = p.rev ? n_items - 1 - item : item;
      const int b = itm / p.H, h = itm % p.H;
      const size_t row0 = static_cast<size_t>(b) * VA_L;
      const int bf = VA_BUF(it);
      const uint32_t bph = VA_USE(it);
      const uint32_t sKa = smem_u32(smem + bf * ITEM_BYTES) + 2 * DHB * VA_QBYTES;
      const uint32_t sVa = sKa + DHB * VA_KBYTES;
      // A fragments of query row 256: a[0] = (row 0, k 2*(lane%4)..+1), a[2] = (row 0, k + 8); rows 8.. are zero
      const uint32_t* qw = reinterpret_cast<const uint32_t*>(p.Q + (row0 + 256) * p.ld + h * DH);
      uint32_t aq[KSTEPS][4];
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        const int w0 = ks * 8 + (lane & 3), w1 = w0 + 4;
        aq[ks][0] = (arow && 2 * w0 < DH) ? __ldg(qw + w0) : 0u;
        aq[ks][1] = 0u;
        aq[ks][2] = (arow && 2 * w1 < DH) ? __ldg(qw + w1) : 0u;
        aq[ks][3] = 0u;
      }
      mbar_wait(&k_full[bf], bph);
      // ---- scores of this warp's 136 keys: 9 pairs of n-tiles (16 keys); lanes 0..3 hold row 0 ----
      float sc[9][4];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
        int key = kbase + 16 * j + (lane & 7) + ((lane >> 4) << 3);
        key = key < VA_LK ? key : VA_LK - 1;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
          const int c = 2 * ks + ((lane >> 3) & 1);
          uint32_t r0, r1, r2, r3;
          ldsm_x4(sKa + (c >> 3) * VA_KBYTES + key * 128 + (((c & 7) ^ (key & 7)) << 4), r0, r1, r2, r3);
          if (p.fp16) {
            mma_16816<true>(s0, aq[ks], r0, r1);
            mma_16816<true>(s1, aq[ks], r2, r3);
          } else {
            mma_16816<false>(s0, aq[ks], r0, r1);
            mma_16816<false>(s1, aq[ks], r2, r3);
          }
        }
        // this lane's keys: n-tile 0 -> kbase + 16 j + 2 (lane % 4) + {0, 1}; n-tile 1 -> + 8
        const int k0 = kbase + 16 * j + 2 * (lane & 3);
        const int khi = tw == 0 ? VA_HALF : VA_L;
        sc[j][0] = (arow && k0 < khi) ? s0[0] * p.scale_log2 : -INFINITY;
        sc[j][1] = (arow && k0 + 1 < khi) ? s0[1] * p.scale_log2 : -INFINITY;
        sc[j][2] = (arow && k0 + 8 < khi) ? s1[0] * p.scale_log2 : -INFINITY;
        sc[j][3] = (arow && k0 + 9 < khi) ? s1[1] * p.scale_log2 : -INFINITY;
        mx = fmaxf(fmaxf(mx, fmaxf(sc[j][0], sc[j][1])), fmaxf(sc[j][2], sc[j][3]));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (lane == 0) sts32f(t_x + tw * 4, mx);
      asm volatile("bar.sync 2, 64;" ::: "memory");     // both tail warps have read K
      if (tw == 0 && lane == 0) mbar_arrive(&k_empty[bf]);
      mx = fmaxf(lds32f(t_x), lds32f(t_x + 4));
      float sum = 0.f;
      uint32_t pa[9][4];
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const float e0 = ex2_approx(sc[j][0] - mx), e1 = ex2_approx(sc[j][1] - mx);
        const float e2 = ex2_approx(sc[j][2] - mx), e3 = ex2_approx(sc[j][3] - mx);
        sum += (e0 + e1) + (e2 + e3);
        pa[j][0] = pack_act(e0, e1, p.fp16);   // the C layout of two n-tiles is the A layout of one k-step
        pa[j][1] = 0u;
        pa[j][2] = pack_act(e2, e3, p.fp16);
        pa[j][3] = 0u;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) sts32f(t_x + 8 + tw * 4, sum);
      mbar_wait(&v_full[bf], bph);
      // ---- O[row 0] = P V over this warp's keys ----
      float oacc[DT][4];
#pragma unroll
      for (int d = 0; d < DT; ++d) oacc[d][0] = oacc[d][1] = oacc[d][2] = oacc[d][3] = 0.f;
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        int key = kbase + 16 * j + ((lane >> 3) & 1) * 8 + (lane & 7);
        key = key < VA_LK ? key : VA_LK - 1;
#pragma unroll
        for (int dp = 0; dp < DT / 2; ++dp) {
          const int c = 2 * dp + (lane >> 4);
          uint32_t r0, r1, r2, r3;
          ldsm_x4_t(sVa + (c >> 3) * VA_KBYTES + key * 128 + (((c & 7) ^ (key & 7)) << 4), r0, r1, r2, r3);
          if (p.fp16) {
            mma_16816<true>(oacc[2 * dp], pa[j], r0, r1);
            mma_16816<true>(oacc[2 * dp + 1], pa[j], r2, r3);
          } else {
            mma_16816<false>(oacc[2 * dp], pa[j], r0, r1);
            mma_16816<false>(oacc[2 * dp + 1], pa[j], r2, r3);
          }
        }
      }
      if (arow) {
#pragma unroll
        for (int d = 0; d < DT; ++d) {
          sts32f(t_o + (tw * 96 + d * 8 + 2 * lane) * 4, oacc[d][0]);
          sts32f(t_o + (tw * 96 + d * 8 + 2 * lane + 1) * 4, oacc[d][1]);
        }
      }
      asm volatile("bar.sync 2, 64;" ::: "memory");     // both tail warps have read V; partial sums are visible
      if (tw == 0) {
        if (lane == 0) mbar_arrive(&v_empty[bf]);
        const float inv = 1.0f / (lds32f(t_x + 8) + lds32f(t_x + 12));
        unsigned short* orow = p.O + (row0 + 256) * p.ldo + h * DH;
#pragma unroll
        for (int c0 = 0; c0 < DHP; c0 += 64) {
          const int col = c0 + 2 * lane;
          if (col < DH) {
            const float o0 = (lds32f(t_o + col * 4) + lds32f(t_o + (96 + col) * 4)) * inv;
            const float o1 = (lds32f(t_o + (col + 1) * 4) + lds32f(t_o + (96 + col + 1) * 4)) * inv;
            *reinterpret_cast<uint32_t*>(orow + col) = pack_act(o0, o1, p.fp16);
          }
        }
      }
      asm volatile("bar.sync 2, 64;" ::: "memory");     // the tail buffers are free for the next item
    }
  }
#undef VA_BUF
#undef VA_USE

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int DH>
int launch_v2(const AttnDesc& a, cudaStream_t st) {
  constexpr int DHB = (DH + 63) / 64;
  constexpr int NBUF = DH <= 64 ? 2 : 1;
  const size_t smem = (size_t)NBUF * DHB * (2 * VA_QBYTES + 2 * VA_KBYTES) + 26 * 8 +
                      (VA_XCH_FLOATS + VA_TAIL_FLOATS) * 4 + 1024;
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t rows = (uint64_t)a.B * a.Lq;
  SPRC_TRY(make_tmap_bf16(&tmQ, a.Q, DH, rows, a.H, a.ldq, DH, 64, 128, 1, 3));
  SPRC_TRY(make_tmap_bf16(&tmK, a.K, DH, rows, a.H, a.ldk, DH, 64, VA_HALF, 1, 3));
  SPRC_TRY(make_tmap_bf16(&tmV, a.V, DH, rows, a.H, a.ldv, DH, 64, VA_HALF, 1, 3));
  VitAttnParams p;
  p.B = a.B;
  p.H = a.H;
  p.ld = a.ldq;
  p.ldo = a.ldo;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.fp16 = act_fp16();
  p.rev = next_sweep_reverse();
  p.Q = reinterpret_cast<const unsigned short*>(a.Q);
  p.K = reinterpret_cast<const unsigned short*>(a.K);
  p.V = reinterpret_cast<const unsigned short*>(a.V);
  p.O = reinterpret_cast<unsigned short*>(a.O);
  static bool attr_set = false;
  if (!attr_set) {
    SPRC_CUDA(cudaFuncSetAttribute(vit_attention_v2_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int items = a.B * a.H;
  const int grid = items < device_sm_count() ? items : device_sm_count();
  prof_begin(st);
  SPRC_CUDA(launch_pdl(vit_attention_v2_kernel<DH>, dim3(grid), dim3(VA_THREADS), smem, st, tmQ, tmK, tmV, p));
  if (prof_enabled()) {
    char tag[56];
    snprintf(tag, sizeof(tag), "vit2 B%d H%d dh%d L%d", a.B, a.H, a.dh, a.Lq);
    prof_end(PROF_ATTN, 4.0 * a.B * a.H * (double)a.Lq * a.Lk * a.dh, 2.0 * a.B * a.H * a.dh * 4.0 * a.Lq, st, tag);
  }
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

// Eligible: the ViT's own shape - packed per-image rows, L = 257, dh 64 or 88, 16-byte aligned head slices.
bool attention_vit_eligible(const AttnDesc& a) {
  return (a.dh == 64 || a.dh == 88) && a.Lq == VA_L && a.Lk == VA_L && !a.key_mask && !a.kv_idx0 &&
         a.q_batch_rows == VA_L && a.kv_batch_rows == VA_L && a.ldq == a.ldk && a.ldk == a.ldv && a.ldq % 8 == 0 &&
         a.ldo % 8 == 0 && a.kv_head_stride == 0;
}

int attention_vit(const AttnDesc& a, cudaStream_t st) {
  if (a.dh == 64) return launch_v2<64>(a, st);
  return launch_v2<88>(a, st);
}

}  // namespace sprc
