// Q-Former cross-attention on tcgen05, second generation: 32 query rows per (sample, head) over the 257 visual tokens
// of one image (gallery / fusion passes, Qformer.py:191-194,438-450) or over cat(reference image, candidate image) =
// 514 tokens (inference_rerank, blip2_qformer_cir_rerank.py:419-436), dh = 64, 12 heads.
//
// As in attention_vit.cu the odd 257th key of every image never reaches the tensor core: keys 0..255 of a segment are
// ONE N = 256 MMA (the 32 query rows replicated into the four TMEM lane quarters, so the four softmax warps each own
// 64 keys of the SAME rows), key 256 is a 64-long dot product per row split over the four warps (operands are in the
// staged tiles), and its value row enters O as a rank-1 update in the epilogue.  That leaves two 256-column TMEM
// regions:
//   one segment per item  (NSEG = 1): regions alternate per item - S of item i+1 is computed while the softmax warps
//                         work on item i; O of an item accumulates in its own region once the scores are consumed;
//   two segments per item (NSEG = 2): the regions hold the two score tiles of one item, the softmax takes one maximum
//                         over both, and O accumulates over both P V products in region 0.
// Shared memory: two stages of (Q replicas | K | V) of one segment, so the whole next segment streams in behind the
// current one; P goes to shared memory as a K-major A operand whose rows 0..31 are the real query rows (the MMA also
// reads 96 rows of whatever follows - they only produce the unused accumulator rows 32..127).
//   warp 0  TMA producer     warp 1  tcgen05.mma issuer     warps 2..5  softmax (+ epilogue in the lane-quarter-0 warp)
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "ops.h"
#include "ptx.cuh"

namespace sprc {

int make_tmap_bf16(CUtensorMap* tm, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
                   uint64_t stride2, uint32_t b0, uint32_t b1, uint32_t b2, int rank);

namespace {

constexpr int XC_LK = 272;               // K / V rows staged per segment (row 256 = the odd key)
constexpr int XC_HALF = XC_LK / 2;       // TMA box rows
constexpr int XC_QBYTES = 128 * 128;     // four replicas of the 32-row query tile
constexpr int XC_KBYTES = XC_LK * 128;
constexpr int XC_STAGE = XC_QBYTES + 2 * XC_KBYTES;
constexpr int XC_OBYTES = 32 * 128;
constexpr int XC_THREADS = 6 * 32;
constexpr int XC_REGION = 256;

struct CrossParams {
  int B, H;
  int q_batch_rows, kv_batch_rows;
  float scale_log2;
  int fp16;
  int rev;
  const int32_t* kv_idx0;   // NSEG = 2: image index of segment 0 / 1 of every sample
  const int32_t* kv_idx1;
};

__device__ __forceinline__ float2 xc_unpack2(uint32_t w, int fp16) {
  if (fp16) {
    const __half2 h = *reinterpret_cast<const __half2*>(&w);
    return __half22float2(h);
  }
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
}
__device__ __forceinline__ float xc_dot8(const uint4& a, const uint4& b, int fp16) {
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 x = xc_unpack2(aw[i], fp16), y = xc_unpack2(bw[i], fp16);
    s = fmaf(x.x, y.x, s);
    s = fmaf(x.y, y.y, s);
  }
  return s;
}

template <int NSEG>
__global__ void __launch_bounds__(XC_THREADS, 1)
qf_cross_attention_v2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                             const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                             const CrossParams p) {
  constexpr int PBYTES = NSEG * 4 * 4096 + 12288;   // 64-key blocks 4 KB apart + the rows 32..127 the MMA also reads
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sP = smem;
  uint8_t* stages = smem + PBYTES;
  uint8_t* sO = stages + 2 * XC_STAGE;
  uint8_t* sV256 = sO + XC_OBYTES;                  // [NSEG][128 B] value rows 256
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV256 + 256);
  // the Q + K half of a stage is free once its scores are out and the softmax warps have read their key-256 operands,
  // the V half only after P V: separate barriers, so the next segment's Q + K stream in under this item's softmax
  uint64_t* full = bars;         // [2] Q + K of the stage loaded
  uint64_t* empty = bars + 2;    // [2] Q + K half free
  uint64_t* s_full = bars + 4;   // [2] scores of the segment in region r
  uint64_t* rfree = bars + 6;    // [2] region r may take new scores
  uint64_t* p_full = bars + 8;   // P of the item is in shared memory
  uint64_t* o_full = bars + 9;   // O of the item is complete
  uint64_t* fullv = bars + 10;   // [2] V of the stage loaded
  uint64_t* emptyv = bars + 12;  // [2] V half free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  const uint32_t xch = smem_u32(bars + 16);   // max[NSEG][4][32], part[NSEG][4][32], sum[4][32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.B * p.H;
  const int n_my = n_items > static_cast<int>(blockIdx.x)
                       ? (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                       : 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&fullv[s], 1);
      mbar_init(&emptyv[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&rfree[s], 1);
    }
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
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
      for (int it = 0; it < n_my; ++it) {
        const int item = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
        const int itm = p.rev ? n_items - 1 - item : item;
        const int b = itm / p.H, h = itm % p.H;
#pragma unroll
        for (int half = 0; half < 2; ++half) {   // Q + K of every segment of the item first, then the V halves
#pragma unroll
          for (int sg = 0; sg < NSEG; ++sg) {
            const int u = it * NSEG + sg;
            const int st = u & 1;
            const uint32_t par = ((u >> 1) & 1) ^ 1;
            uint8_t* sb = stages + st * XC_STAGE;
            const int img = NSEG == 1 ? b : (sg == 0 ? __ldg(p.kv_idx0 + b) : __ldg(p.kv_idx1 + b));
            const int kr = img * p.kv_batch_rows;
            // K/V maps are (d, row, head): heads are column slices of wide rows or contiguous [rows, 64] blocks
            if (half == 0) {
              mbar_wait(&empty[st], par);
              mbar_expect_tx(&full[st], XC_QBYTES + XC_KBYTES);
#pragma unroll
              for (int rep = 0; rep < 4; ++rep)
                tma_load_2d(&tmQ, &full[st], sb + rep * 4096, h * 64, b * p.q_batch_rows, kEvictNormal);
              tma_load_3d(&tmK, &full[st], sb + XC_QBYTES, 0, kr, h, kEvictFirst);
              tma_load_3d(&tmK, &full[st], sb + XC_QBYTES + XC_HALF * 128, 0, kr + XC_HALF, h, kEvictFirst);
            } else {
              mbar_wait(&emptyv[st], par);
              mbar_expect_tx(&fullv[st], XC_KBYTES);
              tma_load_3d(&tmV, &fullv[st], sb + XC_QBYTES + XC_KBYTES, 0, kr, h, kEvictFirst);
              tma_load_3d(&tmV, &fullv[st], sb + XC_QBYTES + XC_KBYTES + XC_HALF * 128, 0, kr + XC_HALF, h, kEvictFirst);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_s = umma_idesc_16(128, 256, p.fp16);
    const uint32_t idesc_pv = umma_idesc_16(128, 64, p.fp16) | (1u << 16);  // B operand MN-major
    auto issue_scores = [&](int it) {
#pragma unroll
      for (int sg = 0; sg < NSEG; ++sg) {
        const int u = it * NSEG + sg;
        const int st = u & 1;
        const uint32_t n = (u >> 1) & 1;
        uint8_t* sb = stages + st * XC_STAGE;
        mbar_wait(&full[st], n);
        mbar_wait(&rfree[st], n ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = umma_desc_k_sw128(smem_u32(sb));
          const uint64_t db = umma_desc_k_sw128(smem_u32(sb + XC_QBYTES));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tmem_base + st * XC_REGION, da + 2 * kk, db + 2 * kk, idesc_s, kk != 0 ? 1u : 0u);
          umma_commit(&s_full[st]);
        }
        __syncwarp();
      }
    };
    if (n_my > 0) issue_scores(0);
    for (int it = 0; it < n_my; ++it) {
      if (NSEG == 1 && it + 1 < n_my) issue_scores(it + 1);   // the other region: runs under this item's softmax
      mbar_wait(p_full, it & 1);   // P of this item is in shared memory, its scores are consumed
#pragma unroll
      for (int sg = 0; sg < NSEG; ++sg) mbar_wait(&fullv[(it * NSEG + sg) & 1], ((it * NSEG + sg) >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t o_addr = tmem_base + (NSEG == 1 ? (it & 1) * XC_REGION : 0);
#pragma unroll
        for (int sg = 0; sg < NSEG; ++sg) {
          const int st = (it * NSEG + sg) & 1;
          const uint32_t sv = smem_u32(stages + st * XC_STAGE + XC_QBYTES + XC_KBYTES);
#pragma unroll
          for (int ks = 0; ks < 16; ++ks) {
            const uint64_t da = umma_desc_k_sw128(smem_u32(sP + (sg * 4 + (ks >> 2)) * 4096)) + 2 * (ks & 3);
            const uint64_t db = umma_desc_mn_sw128(sv + ks * 16 * 128, XC_KBYTES);
            umma_bf16(o_addr, da, db, idesc_pv, (sg | ks) != 0 ? 1u : 0u);
          }
        }
        umma_commit(o_full);
#pragma unroll
        for (int sg = 0; sg < NSEG; ++sg) umma_commit(&emptyv[(it * NSEG + sg) & 1]);
      }
      __syncwarp();
      if (NSEG == 2 && it + 1 < n_my) issue_scores(it + 1);
    }
  } else {
    // ===================== softmax + epilogue (warps 2..5) =====================
    const int q = warp & 3;   // TMEM lane quarter = 64-key block of every segment this warp owns
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t prow = smem_u32(sP) + lane * 128;
    const uint32_t x_max = xch, x_part = xch + NSEG * 128 * 4, x_sum = xch + 2 * NSEG * 128 * 4;
    for (int it = 0; it < n_my; ++it) {
      const int item = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
      const int itm = p.rev ? n_items - 1 - item : item;
      const int b = itm / p.H, h = itm % p.H;
      uint32_t sr[NSEG][64];
      float mxl = -INFINITY;
#pragma unroll
      for (int sg = 0; sg < NSEG; ++sg) {
        const int u = it * NSEG + sg;
        const int st = u & 1;
        const uint32_t sb = smem_u32(stages + st * XC_STAGE);
        mbar_wait(&s_full[st], (u >> 1) & 1);   // the stage of this segment is loaded too (the MMA has read it)
        tc_fence_after();
        {
          uint32_t(&a0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sr[sg][0]);
          uint32_t(&a1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sr[sg][32]);
          const uint32_t s_addr = tmem_base + lane_addr + st * XC_REGION + q * 64;
          tmem_ld32(s_addr, a0);
          tmem_ld32(s_addr + 32, a1);
        }
        // this warp's 16 dims of the key-256 score of row `lane` (replica 0 of Q, K row 256: chunks in place)
        float part = 0.f;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = 2 * q + cc;
          part += xc_dot8(lds128(sb + lane * 128 + ((c ^ (lane & 7)) << 4)),
                          lds128(sb + XC_QBYTES + 256 * 128 + (c << 4)), p.fp16);
        }
        tmem_ld_wait();
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 64; ++j) mx = fmaxf(mx, __uint_as_float(sr[sg][j]));
        mxl = fmaxf(mxl, mx);
        sts32f(x_part + ((sg * 4 + q) * 32 + lane) * 4, part);
      }
      sts32f(x_max + (q * 32 + lane) * 4, mxl);
      tc_fence_before();
      asm volatile("bar.sync 1, 128;" ::: "memory");   // every score of the item has been read
      tc_fence_after();
      if (NSEG == 2 && q == 0 && lane == 0) mbar_arrive(&rfree[1]);   // region 1 only ever holds scores
      if (q == 0 && lane == 0) {   // scores are out, every warp has read Q / K row 256: the Q + K halves are free
#pragma unroll
        for (int sg = 0; sg < NSEG; ++sg) mbar_arrive(&empty[(it * NSEG + sg) & 1]);
      }
      float s256[NSEG];
      float mx = fmaxf(fmaxf(lds32f(x_max + lane * 4), lds32f(x_max + (32 + lane) * 4)),
                       fmaxf(lds32f(x_max + (64 + lane) * 4), lds32f(x_max + (96 + lane) * 4)));
#pragma unroll
      for (int sg = 0; sg < NSEG; ++sg) {
        s256[sg] = (lds32f(x_part + ((sg * 4 + 0) * 32 + lane) * 4) + lds32f(x_part + ((sg * 4 + 1) * 32 + lane) * 4)) +
                   (lds32f(x_part + ((sg * 4 + 2) * 32 + lane) * 4) + lds32f(x_part + ((sg * 4 + 3) * 32 + lane) * 4));
        mx = fmaxf(mx, s256[sg]);
      }
      const float moff = mx * p.scale_log2;
      if (it > 0) mbar_wait(o_full, (it - 1) & 1);   // the previous item's P V has read the P buffer
      float sum = 0.f;
      float p256[NSEG];
#pragma unroll
      for (int sg = 0; sg < NSEG; ++sg) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            const float e0 = ex2_approx(fmaf(__uint_as_float(sr[sg][c * 8 + j]), p.scale_log2, -moff));
            const float e1 = ex2_approx(fmaf(__uint_as_float(sr[sg][c * 8 + j + 1]), p.scale_log2, -moff));
            sum += e0 + e1;
            pk[j / 2] = pack_act(e0, e1, p.fp16);
          }
          // K-major A operand: key block kb at sP + kb*4096, row `lane`, 16-byte chunk c (8 keys), 128B swizzle
          sts128(prow + (sg * 4 + q) * 4096 + ((c ^ (lane & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
        }
        p256[sg] = ex2_approx(fmaf(s256[sg], p.scale_log2, -moff));
        if (q == 0) sum += p256[sg];
      }
      sts32f(x_sum + (q * 32 + lane) * 4, sum);
      if (q == 0) {
        // value rows 256 of the item's segments -> side buffer (the V halves are recycled before the epilogue runs)
#pragma unroll
        for (int sg = 0; sg < NSEG; ++sg) {
          const int u = it * NSEG + sg;
          mbar_wait(&fullv[u & 1], (u >> 1) & 1);
          if (lane < 8) {
            const uint32_t sb = smem_u32(stages + (u & 1) * XC_STAGE);
            const uint4 v = lds128(sb + XC_QBYTES + XC_KBYTES + 256 * 128 + (lane << 4));
            sts128(smem_u32(sV256) + sg * 128 + (lane << 4), v.x, v.y, v.z, v.w);
          }
        }
      }
      fence_proxy_async();   // P (generic-proxy writes) -> tcgen05.mma (async proxy)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      tc_fence_before();
      if (lane == 0) mbar_arrive(p_full);
      if (q == 0) {
        // ---- epilogue (lanes 0..31 hold the real rows): (O + p256 V256) / l -> 16 bit -> staging -> TMA store ----
        const uint32_t o_addr = tmem_base + (NSEG == 1 ? (it & 1) * XC_REGION : 0);
        mbar_wait(o_full, it & 1);
        tc_fence_after();
        uint32_t r[64];
        {
          uint32_t(&a0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
          uint32_t(&a1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[32]);
          tmem_ld32(o_addr, a0);
          tmem_ld32(o_addr + 32, a1);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&rfree[NSEG == 1 ? (it & 1) : 0]);
          bulk_wait_read0();  // the previous item's store has read the staging tile
        }
        __syncwarp();
        const float inv = 1.0f / ((lds32f(x_sum + lane * 4) + lds32f(x_sum + (32 + lane) * 4)) +
                                  (lds32f(x_sum + (64 + lane) * 4) + lds32f(x_sum + (96 + lane) * 4)));
        const uint32_t stg = smem_u32(sO) + lane * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = __uint_as_float(r[8 * c + j]);
#pragma unroll
          for (int sg = 0; sg < NSEG; ++sg) {
            const uint4 v = lds128(smem_u32(sV256) + sg * 128 + (c << 4));
            const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 v2 = xc_unpack2(vw[j], p.fp16);
              o[2 * j] = fmaf(p256[sg], v2.x, o[2 * j]);
              o[2 * j + 1] = fmaf(p256[sg], v2.y, o[2 * j + 1]);
            }
          }
          sts128(stg + ((c ^ (lane & 7)) << 4), pack_act(o[0] * inv, o[1] * inv, p.fp16),
                 pack_act(o[2] * inv, o[3] * inv, p.fp16), pack_act(o[4] * inv, o[5] * inv, p.fp16),
                 pack_act(o[6] * inv, o[7] * inv, p.fp16));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmO, smem_u32(sO), h * 64, b * p.q_batch_rows);
          bulk_commit();
        }
      }
      // x_sum / sV256 of this item are read by the quarter-0 warp before it joins the next item's first bar.sync
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

// ------------------------------------------------------------------------------------------------
// One segment (257 keys), TWO softmax groups.  ncu on the kernel above (profiles/r02n_ncu_cross_attention_summary.txt):
// 14.6 % issue activity, 52 % of DRAM throughput, 4.6 k cycles per item - an item is a serial chain
// (load -> S MMA -> softmax -> P V MMA -> epilogue) and one group of four softmax warps can only walk one chain at a
// time, so the time of a launch follows the SM clock.  Here warps 2..5 own the even items of the CTA (stage 0, TMEM
// region 0) and warps 6..9 the odd ones (stage 1, region 1): two chains per SM in antiphase, each with its own P
// buffer, exchange scratch, staging tile and named barrier.  The MMA issuer is ONE polling thread that serves whichever
// chain is ready (scores of a loaded stage, or P V of a finished softmax) instead of sleeping on one barrier.
// ------------------------------------------------------------------------------------------------
constexpr int XG_THREADS = 10 * 32;
constexpr int XG_PGROUP = 4 * 4096;                 // P of one group: four 64-key blocks 4 KB apart
constexpr int XG_PBYTES = 2 * XG_PGROUP + 12288;    // + the rows 32..127 the last block's MMA also reads
constexpr int XG_XCH = 3 * 128 * 4;                 // per group: max[4][32], part[4][32], sum[4][32]
constexpr int XG_SMEM = XG_PBYTES + 2 * XC_STAGE + 2 * XC_OBYTES + 256 + 160 + 2 * XG_XCH + 1024;

__global__ void __launch_bounds__(XG_THREADS, 1)
qf_cross_attention_g2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                             const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                             const CrossParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sP = smem;
  uint8_t* stages = smem + XG_PBYTES;
  uint8_t* sO = stages + 2 * XC_STAGE;              // [2][XC_OBYTES]
  uint8_t* sV256 = sO + 2 * XC_OBYTES;              // [2][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV256 + 256);
  // The K half of a stage (Q replicas + K) is free again as soon as the scores are out and the softmax warps have read
  // their key-256 operands; the V half only after P V.  Separate barriers let K of the group's NEXT item stream in under
  // this item's softmax and P V (with one barrier per stage the load sat in series with the chain).
  uint64_t* fullk = bars;        // [2] Q + K of group g's item loaded
  uint64_t* emptyk = bars + 2;   // [2] Q + K half of stage g free
  uint64_t* s_full = bars + 4;   // [2] scores of group g's item are in region g
  uint64_t* rfree = bars + 6;    // [2] region g may take new scores
  uint64_t* p_full = bars + 8;   // [2] P of group g's item is in shared memory
  uint64_t* o_full = bars + 10;  // [2] O of group g's item is complete
  uint64_t* fullv = bars + 12;   // [2] V of group g's item loaded
  uint64_t* emptyv = bars + 14;  // [2] V half of stage g free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t xch0 = smem_u32(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.B * p.H;
  const int n_my = n_items > static_cast<int>(blockIdx.x)
                       ? (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                       : 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&fullk[s], 1);
      mbar_init(&emptyk[s], 1);
      mbar_init(&fullv[s], 1);
      mbar_init(&emptyv[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&rfree[s], 1);
      mbar_init(&p_full[s], 4);
      mbar_init(&o_full[s], 1);
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

  if (warp == 0) {
    // ===================== TMA producer: item `it` -> stage it & 1, K half and V half on their own barriers =====
    if (elect_one()) {
      int next_k[2] = {0, 1}, next_v[2] = {0, 1};
      while (next_v[0] < n_my || next_v[1] < n_my) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint8_t* sb = stages + g * XC_STAGE;
          int it = next_k[g];
          if (it < n_my && mbar_test(&emptyk[g], ((it >> 1) & 1) ^ 1)) {
            const int item = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
            const int itm = p.rev ? n_items - 1 - item : item;
            const int b = itm / p.H, h = itm % p.H;
            const int kr = b * p.kv_batch_rows;
            mbar_expect_tx(&fullk[g], XC_QBYTES + XC_KBYTES);
#pragma unroll
            for (int rep = 0; rep < 4; ++rep)
              tma_load_2d(&tmQ, &fullk[g], sb + rep * 4096, h * 64, b * p.q_batch_rows, kEvictNormal);
            tma_load_3d(&tmK, &fullk[g], sb + XC_QBYTES, 0, kr, h, kEvictFirst);
            tma_load_3d(&tmK, &fullk[g], sb + XC_QBYTES + XC_HALF * 128, 0, kr + XC_HALF, h, kEvictFirst);
            next_k[g] = it + 2;
          }
          it = next_v[g];
          if (it < n_my && it < next_k[g] && mbar_test(&emptyv[g], ((it >> 1) & 1) ^ 1)) {
            const int item = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
            const int itm = p.rev ? n_items - 1 - item : item;
            const int b = itm / p.H, h = itm % p.H;
            const int kr = b * p.kv_batch_rows;
            mbar_expect_tx(&fullv[g], XC_KBYTES);
            tma_load_3d(&tmV, &fullv[g], sb + XC_QBYTES + XC_KBYTES, 0, kr, h, kEvictFirst);
            tma_load_3d(&tmV, &fullv[g], sb + XC_QBYTES + XC_KBYTES + XC_HALF * 128, 0, kr + XC_HALF, h, kEvictFirst);
            next_v[g] = it + 2;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: one polling thread, two chains =====================
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_16(128, 256, p.fp16);
      const uint32_t idesc_pv = umma_idesc_16(128, 64, p.fp16) | (1u << 16);  // B operand MN-major
      int s_next[2] = {0, 1}, pv_next[2] = {0, 1};   // next item of each group whose scores / P V are to be issued
      while (pv_next[0] < n_my || pv_next[1] < n_my) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint8_t* sb = stages + g * XC_STAGE;
          int it = s_next[g];
          if (it < n_my) {
            const uint32_t n = (it >> 1) & 1;
            if (mbar_test(&fullk[g], n) && mbar_test(&rfree[g], n ^ 1)) {
              tc_fence_after();
              const uint64_t da = umma_desc_k_sw128(smem_u32(sb));
              const uint64_t db = umma_desc_k_sw128(smem_u32(sb + XC_QBYTES));
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_bf16(tmem_base + g * XC_REGION, da + 2 * kk, db + 2 * kk, idesc_s, kk != 0 ? 1u : 0u);
              umma_commit(&s_full[g]);
              s_next[g] = it + 2;
            }
          }
          it = pv_next[g];
          if (it < n_my && it < s_next[g] && mbar_test(&p_full[g], (it >> 1) & 1) &&
              mbar_test(&fullv[g], (it >> 1) & 1)) {
            tc_fence_after();
            const uint32_t sv = smem_u32(sb + XC_QBYTES + XC_KBYTES);
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) {
              const uint64_t da = umma_desc_k_sw128(smem_u32(sP + g * XG_PGROUP + (ks >> 2) * 4096)) + 2 * (ks & 3);
              const uint64_t db = umma_desc_mn_sw128(sv + ks * 16 * 128, XC_KBYTES);
              umma_bf16(tmem_base + g * XC_REGION, da, db, idesc_pv, ks != 0 ? 1u : 0u);
            }
            umma_commit(&o_full[g]);
            umma_commit(&emptyv[g]);
            pv_next[g] = it + 2;
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax + epilogue: group g = warps 2 + 4 g .. 5 + 4 g =====================
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;   // TMEM lane quarter = 64-key block this warp owns
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t prow = smem_u32(sP) + g * XG_PGROUP + lane * 128;
    const uint32_t x_max = xch0 + g * XG_XCH, x_part = x_max + 128 * 4, x_sum = x_max + 2 * 128 * 4;
    const uint32_t sb = smem_u32(stages + g * XC_STAGE);
    const uint32_t sv256 = smem_u32(sV256) + g * 128;
    const uint32_t so = smem_u32(sO) + g * XC_OBYTES;
    const uint32_t region = tmem_base + g * XC_REGION;
    for (int it = g; it < n_my; it += 2) {
      const uint32_t par = (it >> 1) & 1;
      const int item = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
      const int itm = p.rev ? n_items - 1 - item : item;
      const int b = itm / p.H, h = itm % p.H;
      uint32_t sr[64];
      mbar_wait(&s_full[g], par);   // Q + K are loaded too (the MMA has read them)
      tc_fence_after();
      {
        uint32_t(&a0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sr[0]);
        uint32_t(&a1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sr[32]);
        const uint32_t s_addr = region + lane_addr + q * 64;
        tmem_ld32(s_addr, a0);
        tmem_ld32(s_addr + 32, a1);
      }
      // this warp's 16 dims of the key-256 score of row `lane` (replica 0 of Q, K row 256: chunks in place)
      float part = 0.f;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = 2 * q + cc;
        part += xc_dot8(lds128(sb + lane * 128 + ((c ^ (lane & 7)) << 4)),
                        lds128(sb + XC_QBYTES + 256 * 128 + (c << 4)), p.fp16);
      }
      tmem_ld_wait();
      float mxl = -INFINITY;
#pragma unroll
      for (int j = 0; j < 64; ++j) mxl = fmaxf(mxl, __uint_as_float(sr[j]));
      sts32f(x_part + (q * 32 + lane) * 4, part);
      sts32f(x_max + (q * 32 + lane) * 4, mxl);
      tc_fence_before();
      if (g == 0)
        asm volatile("bar.sync 1, 128;" ::: "memory");   // every score of the item has been read
      else
        asm volatile("bar.sync 2, 128;" ::: "memory");
      tc_fence_after();
      if (q == 0 && lane == 0) mbar_arrive(&emptyk[g]);   // scores are out, every warp has read Q / K row 256
      float mx = fmaxf(fmaxf(lds32f(x_max + lane * 4), lds32f(x_max + (32 + lane) * 4)),
                       fmaxf(lds32f(x_max + (64 + lane) * 4), lds32f(x_max + (96 + lane) * 4)));
      const float s256 = (lds32f(x_part + lane * 4) + lds32f(x_part + (32 + lane) * 4)) +
                         (lds32f(x_part + (64 + lane) * 4) + lds32f(x_part + (96 + lane) * 4));
      mx = fmaxf(mx, s256);
      const float moff = mx * p.scale_log2;
      if (it >= 2) mbar_wait(&o_full[g], par ^ 1);   // the group's previous P V has read the P buffer
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const float e0 = ex2_approx(fmaf(__uint_as_float(sr[c * 8 + j]), p.scale_log2, -moff));
          const float e1 = ex2_approx(fmaf(__uint_as_float(sr[c * 8 + j + 1]), p.scale_log2, -moff));
          sum += e0 + e1;
          pk[j / 2] = pack_act(e0, e1, p.fp16);
        }
        // K-major A operand: key block kb at +kb*4096, row `lane`, 16-byte chunk c (8 keys), 128B swizzle
        sts128(prow + q * 4096 + ((c ^ (lane & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
      }
      const float p256 = ex2_approx(fmaf(s256, p.scale_log2, -moff));
      if (q == 0) sum += p256;
      sts32f(x_sum + (q * 32 + lane) * 4, sum);
      if (q == 0) {
        // value row 256 -> side buffer (the V half is recycled before the epilogue runs)
        mbar_wait(&fullv[g], par);
        if (lane < 8) {
          const uint4 v = lds128(sb + XC_QBYTES + XC_KBYTES + 256 * 128 + (lane << 4));
          sts128(sv256 + (lane << 4), v.x, v.y, v.z, v.w);
        }
      }
      fence_proxy_async();   // P (generic-proxy writes) -> tcgen05.mma (async proxy)
      if (g == 0)
        asm volatile("bar.sync 1, 128;" ::: "memory");
      else
        asm volatile("bar.sync 2, 128;" ::: "memory");
      tc_fence_before();
      if (lane == 0) mbar_arrive(&p_full[g]);
      if (q == 0) {
        // ---- epilogue (lanes 0..31 hold the real rows): (O + p256 V256) / l -> 16 bit -> staging -> TMA store ----
        mbar_wait(&o_full[g], par);
        tc_fence_after();
        uint32_t r[64];
        {
          uint32_t(&a0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
          uint32_t(&a1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[32]);
          tmem_ld32(region, a0);
          tmem_ld32(region + 32, a1);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&rfree[g]);
          bulk_wait_read0();  // the group's previous store has read the staging tile
        }
        __syncwarp();
        const float inv = 1.0f / ((lds32f(x_sum + lane * 4) + lds32f(x_sum + (32 + lane) * 4)) +
                                  (lds32f(x_sum + (64 + lane) * 4) + lds32f(x_sum + (96 + lane) * 4)));
        const uint32_t stg = so + lane * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = __uint_as_float(r[8 * c + j]);
          const uint4 v = lds128(sv256 + (c << 4));
          const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 v2 = xc_unpack2(vw[j], p.fp16);
            o[2 * j] = fmaf(p256, v2.x, o[2 * j]);
            o[2 * j + 1] = fmaf(p256, v2.y, o[2 * j + 1]);
          }
          sts128(stg + ((c ^ (lane & 7)) << 4), pack_act(o[0] * inv, o[1] * inv, p.fp16),
                 pack_act(o[2] * inv, o[3] * inv, p.fp16), pack_act(o[4] * inv, o[5] * inv, p.fp16),
                 pack_act(o[6] * inv, o[7] * inv, p.fp16));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmO, so, h * 64, b * p.q_batch_rows);
          bulk_commit();
        }
      }
      // x_sum / sV256 of this item are read by the quarter-0 warp before it joins the group's next first bar.sync
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

int launch_cross_g2(const AttnDesc& a, cudaStream_t st) {
  CUtensorMap tmQ, tmK, tmV, tmO;
  const uint64_t w = (uint64_t)a.H * 64;
  const uint64_t qrows = (uint64_t)(a.B - 1) * a.q_batch_rows + a.Lq;
  const uint64_t krows = a.kv_rows_total > 0 ? (uint64_t)a.kv_rows_total : (uint64_t)(a.B - 1) * a.kv_batch_rows + 257;
  SPRC_TRY(make_tmap_bf16(&tmQ, a.Q, w, qrows, 1, a.ldq, 0, 64, 32, 1, 2));
  const uint64_t hstride = a.kv_head_stride > 0 ? (uint64_t)a.kv_head_stride : 64;
  SPRC_TRY(make_tmap_bf16(&tmK, a.K, 64, krows, a.H, a.ldk, hstride, 64, XC_HALF, 1, 3));
  SPRC_TRY(make_tmap_bf16(&tmV, a.V, 64, krows, a.H, a.ldv, hstride, 64, XC_HALF, 1, 3));
  SPRC_TRY(make_tmap_bf16(&tmO, a.O, w, qrows, 1, a.ldo, 0, 64, 32, 1, 2));
  CrossParams p;
  p.B = a.B;
  p.H = a.H;
  p.q_batch_rows = a.q_batch_rows;
  p.kv_batch_rows = a.kv_batch_rows;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.fp16 = act_fp16();
  p.rev = next_sweep_reverse();
  p.kv_idx0 = nullptr;
  p.kv_idx1 = nullptr;
  static bool attr_set = false;
  if (!attr_set) {
    SPRC_CUDA(cudaFuncSetAttribute(qf_cross_attention_g2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XG_SMEM));
    attr_set = true;
  }
  const int items = a.B * a.H;
  const int grid = items < device_sm_count() ? items : device_sm_count();
  prof_begin(st);
  SPRC_CUDA(launch_pdl(qf_cross_attention_g2_kernel, dim3(grid), dim3(XG_THREADS), XG_SMEM, st, tmQ, tmK, tmV, tmO, p));
  if (prof_enabled()) {
    char tag[56];
    snprintf(tag, sizeof(tag), "qf-cross-g2 B%d H%d Lq%d Lk%d", a.B, a.H, a.Lq, a.Lk);
    prof_end(PROF_ATTN, 4.0 * a.B * a.H * (double)a.Lq * a.Lk * 64, 2.0 * a.B * a.H * 64 * (2.0 * a.Lq + 2.0 * a.Lk), st,
             tag);
  }
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

template <int NSEG>
int launch_cross2(const AttnDesc& a, cudaStream_t st) {
  constexpr int PBYTES = NSEG * 4 * 4096 + 12288;
  const size_t smem = PBYTES + 2 * XC_STAGE + XC_OBYTES + 256 + 16 * 8 + (2 * NSEG + 1) * 128 * 4 + 1024;
  CUtensorMap tmQ, tmK, tmV, tmO;
  const uint64_t w = (uint64_t)a.H * 64;
  const uint64_t qrows = (uint64_t)(a.B - 1) * a.q_batch_rows + a.Lq;
  const uint64_t krows = a.kv_rows_total > 0 ? (uint64_t)a.kv_rows_total : (uint64_t)(a.B - 1) * a.kv_batch_rows + 257;
  SPRC_TRY(make_tmap_bf16(&tmQ, a.Q, w, qrows, 1, a.ldq, 0, 64, 32, 1, 2));
  const uint64_t hstride = a.kv_head_stride > 0 ? (uint64_t)a.kv_head_stride : 64;
  SPRC_TRY(make_tmap_bf16(&tmK, a.K, 64, krows, a.H, a.ldk, hstride, 64, XC_HALF, 1, 3));
  SPRC_TRY(make_tmap_bf16(&tmV, a.V, 64, krows, a.H, a.ldv, hstride, 64, XC_HALF, 1, 3));
  SPRC_TRY(make_tmap_bf16(&tmO, a.O, w, qrows, 1, a.ldo, 0, 64, 32, 1, 2));
  CrossParams p;
  p.B = a.B;
  p.H = a.H;
  p.q_batch_rows = a.q_batch_rows;
  p.kv_batch_rows = a.kv_batch_rows;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.fp16 = act_fp16();
  p.rev = next_sweep_reverse();
  p.kv_idx0 = a.kv_idx0;
  p.kv_idx1 = a.kv_idx1;
  static bool attr_set = false;
  if (!attr_set) {
    SPRC_CUDA(cudaFuncSetAttribute(qf_cross_attention_v2_kernel<NSEG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    attr_set = true;
  }
  const int items = a.B * a.H;
  const int grid = items < device_sm_count() ? items : device_sm_count();
  prof_begin(st);
  SPRC_CUDA(launch_pdl(qf_cross_attention_v2_kernel<NSEG>, dim3(grid), dim3(XC_THREADS), smem, st, tmQ, tmK, tmV, tmO,
                       p));
  if (prof_enabled()) {
    char tag[56];
    snprintf(tag, sizeof(tag), "qf-cross2 B%d H%d Lq%d Lk%d", a.B, a.H, a.Lq, a.Lk);
    prof_end(PROF_ATTN, 4.0 * a.B * a.H * (double)a.Lq * a.Lk * 64, 2.0 * a.B * a.H * 64 * (2.0 * a.Lq + 2.0 * a.Lk), st,
             tag);
  }
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

// Eligible: the Q-Former's cross-attention over whole images - 32 query rows, 257 keys per segment, one segment
// (rows of sample b) or two (image index tables kv_idx0 / kv_idx1 over a K/V table of kv_rows_total rows).
bool attention_cross2_eligible(const AttnDesc& a) {
  if (a.dh != 64 || a.Lq != 32 || a.key_mask || a.q_batch_rows < 32 || a.kv_batch_rows != 257) return false;
  if (a.ldq % 8 != 0 || a.ldo % 8 != 0) return false;
  if (a.kv_idx0 && a.kv_idx1) return a.Lk == 514 && a.Lk1 == 257 && a.kv_rows_total > 0;
  return !a.kv_idx0 && !a.kv_idx1 && a.Lk == 257;
}

int attention_cross2(const AttnDesc& a, cudaStream_t st) {
  if (a.kv_idx0) return launch_cross2<2>(a, st);
  static const bool one_group = getenv("SPRC_CROSS_ATTN_1G") != nullptr;   // A/B switch: one softmax group
  return one_group ? launch_cross2<1>(a, st) : launch_cross_g2(a, st);
}

}  // namespace sprc
