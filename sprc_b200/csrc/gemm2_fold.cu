// LayerNorm folded into the neighbouring GEMMs (GemmFold, common.h): the two epilogues that replace LayerNorm kernels,
// as a SEPARATE instantiation of the CTA-pair GEMM of gemm2.cu (the default kernel is not touched; main loop, roles and
// barriers are the same, see there).
//
//   CONSUMER (MODE 1, st_in set):  out16 = act(rstd * (A_raw Wf^T - mean * c) + d)
//       A_raw = the raw 16-bit copy of the pre-LN sums, Wf = round16(W diag(gamma)), c = row sums of Wf,
//       d = W beta + b (arrives as the bias); (mean, rstd) of a row from its K / 64 (mean, M2) partials.
//   PRODUCER (MODE 2, st_out set): s' = acc + b + r,  r = resid (already normalised) or
//       (resid - mean) * rstd * g + beta; writes s' (fp32, may alias resid), its raw 16-bit copy and the statistics
//       partials of s' (one per 64-column slice = one epilogue thread).
//
// Second version of the producer epilogue.  The first one (commit 14bbc55, measured in r02a: GEMMs 43.7 -> 61.2 ms per
// step) fetched the residual with per-thread row-strided global loads on the critical path of the accumulator drain.
// Here the residual chunk (32 rows x 16 fp32) of every epilogue warp is PREFETCHED by TMA into the warp's staging
// buffer two chunks ahead - the first two chunks of a tile while its main loop still runs - the sum is formed in place
// and leaves by TMA store; the 16-bit copy (32 full bytes per row) and the statistics go out as plain vector stores.
// Staging: two 2 KB buffers per warp (64 KB), so the operand ring has 4 stages instead of 6.
#include <stdio.h>
#include <stdlib.h>

#include "common.h"
#include "ops.h"
#include "ptx.cuh"

namespace sprc {

int make_tmap_any(CUtensorMap* tm, const void* ptr, int esz, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
                  uint64_t stride2, uint32_t b0, uint32_t b1, uint32_t b2, int rank, int swizzle_bytes);

namespace {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int EPI_WARPS = 16;
constexpr int STATS_WARP = 2 + EPI_WARPS;      // warp 18: per-row (mean, rstd) of the tile's 128 rows, one tile ahead
constexpr int THREADS = (3 + EPI_WARPS) * 32;
constexpr int STATS_BYTES = BM * 8;            // 128 (mean, rstd) pairs (one buffer: the ring leaves no room for two)
constexpr int A_BYTES = BM * BK * 2;          // 16 KB
constexpr int B_BYTES = (BN / 2) * BK * 2;    // 16 KB
constexpr int CHUNK_BYTES = 32 * 64;          // 32 rows x 16 fp32 (or x 32 16-bit values)

template <int MODE>
struct Cfg {
  static constexpr int STAGES = MODE == 2 ? 4 : 6;
  static constexpr int EPI_PER_WARP = MODE == 2 ? 2 * CHUNK_BYTES : CHUNK_BYTES;
  static constexpr int RING_BYTES = STAGES * (A_BYTES + B_BYTES);
  static constexpr int EPI_BYTES = EPI_WARPS * EPI_PER_WARP;
  static constexpr int BAR_BYTES = (2 * STAGES + 4 + (MODE == 2 ? 2 * EPI_WARPS : 0) + 2) * 8 + 16;
  static_assert(RING_BYTES + EPI_BYTES + STATS_BYTES + BAR_BYTES + 1024 <= 227 * 1024, "shared memory budget");
  static constexpr int SMEM_TOTAL = RING_BYTES + EPI_BYTES + STATS_BYTES + BAR_BYTES + 1024;
};

struct FoldParams {
  int M, N, K;
  int num_m_pairs, num_n_blocks, num_k_blocks;
  const float* bias;
  const float* bias2;
  int act;
  int fp16;
  int rev;
  int m_split;
  GemmFold f;
};

// Statistics partials are stored PART-major, st[part * M + row]: the 32 lanes of an epilogue warp (= 32 consecutive rows)
// read and write 256 contiguous bytes per partial.  (Row-major partials, as in the first version, cost 32 cache lines
// per load instruction and made the consumer 0.6-0.8x the plain GEMM: profiles/r02v_fold_probe.log.)
// One pass, Chan's merge of (mean, M2) over equal-sized slices of 64 columns.
// Per-row (mean, rstd) from the (mean, M2) partials of the row's 64-column slices (Chan's merge, equal slice sizes).
// Partials are stored PART-major, st[part * M + row]: the 32 lanes of a warp (= 32 consecutive rows) read and write
// 256 contiguous bytes per partial.  ONE warp per CTA does this for the 128 rows of a tile, one tile ahead of the
// epilogue, and hands the result over through shared memory: when each of the 16 epilogue warps merged the partials of
// its own rows (r02v-r02x), the epilogue's instruction count doubled (ncu r02y: 140 M against 67.5 M warp instructions,
// tensor pipe 58 % against 90 %) and the epilogue, not the MMA, paced the kernel.
// Lane l owns rows l, l + 32, l + 64, l + 96 of the tile; four partials of all four rows are loaded per batch.
__device__ __forceinline__ void tile_row_stats(const float2* __restrict__ st_lo, const float2* __restrict__ st_hi,
                                               int split, int m0, int lane, int M, int stride, int parts,
                                               float eps, float2* __restrict__ out) {
  float m[4], m2[4];
  const float2* base[4];
  bool ok[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = m0 + r * 32 + lane;
    ok[r] = row < M;
    const float2* st = (split > 0 && m0 + r * 32 >= split) ? st_hi : st_lo;
    base[r] = st + (ok[r] ? row : 0);
    m[r] = 0.f;
    m2[r] = 0.f;
  }
  for (int i0 = 0; i0 < parts; i0 += 4) {
    float2 pt[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r)
        pt[j][r] = (i0 + j < parts) ? __ldg(base[r] + static_cast<size_t>(i0 + j) * stride) : make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + j;
      if (i < parts) {
        const float inv = __frcp_rn(static_cast<float>(i + 1));
        const float w = 64.0f * static_cast<float>(i) * inv;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float dlt = pt[j][r].x - m[r];
          m[r] += dlt * inv;
          m2[r] += pt[j][r].y + dlt * dlt * w;
        }
      }
    }
  }
  const float invn = 1.0f / (64.0f * static_cast<float>(parts));
#pragma unroll
  for (int r = 0; r < 4; ++r)
    out[r * 32 + lane] = ok[r] ? make_float2(m[r], rsqrtf(m2[r] * invn + eps)) : make_float2(0.f, 1.f);
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1)
gemm_fold_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmC,
                      const __grid_constant__ CUtensorMap tmR, const FoldParams p) {
  constexpr int STAGES = Cfg<MODE>::STAGES;
  constexpr int RING_BYTES = Cfg<MODE>::RING_BYTES;
  constexpr int EPI_BYTES = Cfg<MODE>::EPI_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* sEpi = smem + RING_BYTES;
  float2* sStats = reinterpret_cast<float2*>(smem + RING_BYTES + EPI_BYTES);          // [2][BM] (mean, rstd)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + RING_BYTES + EPI_BYTES + STATS_BYTES);   // leader only
  uint64_t* empty_bar = full_bar + STAGES;                                          // one per CTA
  uint64_t* tfull_bar = empty_bar + STAGES;                                         // one per CTA
  uint64_t* tempty_bar = tfull_bar + 2;                                             // used in the leader only
  uint64_t* res_bar = tempty_bar + 2;                                               // [EPI_WARPS][2] residual chunk landed
  uint64_t* sfull_bar = res_bar + (MODE == 2 ? 2 * EPI_WARPS : 0);                   // statistics of a tile are in sStats
  uint64_t* sempty_bar = sfull_bar + 1;                                              // all epilogue warps have read them
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sempty_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();   // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  const int num_tiles = p.num_m_pairs * p.num_n_blocks;
  const GemmFold& f = p.f;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.m_split > 0) tma_prefetch_desc(&tmB2);
    tma_prefetch_desc(&tmC);
    if (MODE == 2) tma_prefetch_desc(&tmR);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * EPI_WARPS);
    }
    if (MODE == 2)
      for (int s = 0; s < 2 * EPI_WARPS; ++s) mbar_init(&res_bar[s], 1);
    mbar_init(sfull_bar, 1);
    mbar_init(sempty_bar, EPI_WARPS);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();
  griddep_launch();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      const uint32_t leader_full = mapa_u32(smem_u32(full_bar), 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < num_tiles; t += npairs) {
        const int tile = p.rev ? num_tiles - 1 - t : t;
        const int m0 = (tile / p.num_n_blocks) * (2 * BM) + static_cast<int>(crank) * BM;
        const int n0 = (tile % p.num_n_blocks) * BN + static_cast<int>(crank) * (BN / 2);
        const CUtensorMap* tmW = (p.m_split > 0 && (tile / p.num_n_blocks) * (2 * BM) >= p.m_split) ? &tmB2 : &tmB;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * (A_BYTES + B_BYTES));
          const uint32_t bar = leader_full + stage * 8;
          tma_load_3d_2cta(&tmA, bar, sA + stage * A_BYTES, kb * BK, m0, 0, kEvictNormal);
          tma_load_2d_2cta(tmW, bar, sB + stage * B_BYTES, kb * BK, n0, kEvictLast);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (crank == 0) {
      const uint32_t idesc = umma_idesc_16(2 * BM, BN, p.fp16);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = pair; t < num_tiles; t += npairs) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = umma_desc_k_sw128(smem_u32(sA + stage * A_BYTES));
            const uint64_t db = umma_desc_k_sw128(smem_u32(sB + stage * B_BYTES));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_2cta(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2cta(&empty_bar[stage]);
            if (kb == p.num_k_blocks - 1) umma_commit_2cta(&tfull_bar[as]);
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else if (warp == STATS_WARP) {
    // ===================== row statistics of the tile's 128 rows, one tile ahead of the epilogue =====================
    const float2* st_lo = MODE == 1 ? f.st_in : f.st_res;
    const float2* st_hi = MODE == 1 ? f.st_in2 : f.st_res2;
    if (st_lo != nullptr) {
      const int parts = (MODE == 1 ? p.K : p.N) >> 6;
      uint32_t sph = 0;
      for (int t = pair; t < num_tiles; t += npairs) {
        const int tile = p.rev ? num_tiles - 1 - t : t;
        const int m0 = (tile / p.num_n_blocks) * (2 * BM) + static_cast<int>(crank) * BM;
        mbar_wait(sempty_bar, sph ^ 1);   // every epilogue warp has read the previous tile's statistics
        tile_row_stats(st_lo, st_hi ? st_hi : st_lo, f.split, m0, lane, p.M, f.st_stride, parts, f.eps, sStats);
        __syncwarp();
        if (lane == 0) mbar_arrive(sfull_bar);
        sph ^= 1;
      }
    }
  } else {
    // ===================== epilogue (warps 2..17): 32 rows x 64 columns per warp and tile =====================
    const int q = warp & 3;
    const int cpart = (warp - 2) >> 2;
    const int ew = warp - 2;
    const uint32_t stile = smem_u32(sEpi) + ew * Cfg<MODE>::EPI_PER_WARP;
    const uint32_t sw = (lane >> 1) & 3;
    const uint32_t leader_tempty = mapa_u32(smem_u32(tempty_bar), 0);
    int as = 0;
    uint32_t aphase = 0;

    uint32_t sph = 0;   // statistics phase (consumer: always; producer: when the residual is normalised)
    if constexpr (MODE == 1) {
      // ---------------- consumer: 16-bit output, two 32-column chunks ----------------
      const uint32_t srow = stile + lane * 64;
      for (int t = pair; t < num_tiles; t += npairs) {
        const int tile = p.rev ? num_tiles - 1 - t : t;
        const int m0 = (tile / p.num_n_blocks) * (2 * BM) + static_cast<int>(crank) * BM + q * 32;
        const int n0 = (tile % p.num_n_blocks) * BN + cpart * (BN / 4);
        const bool w2 = p.m_split > 0 && m0 >= p.m_split;
        const float* bias = w2 ? p.bias2 : p.bias;
        const float* f_c = w2 ? f.c2 : f.c;
        mbar_wait(sfull_bar, sph);
        const float2 ms = sStats[q * 32 + lane];
        __syncwarp();
        if (lane == 0) mbar_arrive(sempty_bar);
        sph ^= 1;
        const float f_mu = ms.x, f_rs = ms.y;
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>(as * BN + cpart * (BN / 4));
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int n = n0 + cc * 32;
          const bool live = n < p.N && m0 < p.M;
          uint32_t r[32], o[16];
          tmem_ld32(t_row + cc * 32, r);
          tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n) + j);
              const float4 c4 = __ldg(reinterpret_cast<const float4*>(f_c + n) + j);
              float v0 = f_rs * (__uint_as_float(r[4 * j]) - f_mu * c4.x) + b.x;
              float v1 = f_rs * (__uint_as_float(r[4 * j + 1]) - f_mu * c4.y) + b.y;
              float v2 = f_rs * (__uint_as_float(r[4 * j + 2]) - f_mu * c4.z) + b.z;
              float v3 = f_rs * (__uint_as_float(r[4 * j + 3]) - f_mu * c4.w) + b.w;
              if (p.act == ACT_GELU) {
                const float2 g0 = gelu_erf2(make_float2(v0, v1)), g1 = gelu_erf2(make_float2(v2, v3));
                v0 = g0.x, v1 = g0.y, v2 = g1.x, v3 = g1.y;
              } else if (p.act == ACT_QUICKGELU) {
                v0 = quick_gelu(v0), v1 = quick_gelu(v1), v2 = quick_gelu(v2), v3 = quick_gelu(v3);
              }
              o[2 * j] = pack_act(v0, v1, p.fp16);
              o[2 * j + 1] = pack_act(v2, v3, p.fp16);
            }
          }
          if (cc == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_addr(leader_tempty + as * 8);
          }
          if (live) {
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128(srow + ((j ^ sw) << 4), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&tmC, stile, n, m0, 0);
              bulk_commit();
            }
          }
        }
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
      if (lane == 0) bulk_wait0();
    } else {
      // ---------------- producer: fp32 sum in place in the staging buffer, four 16-column chunks ----------------
      // chunk sequence of this warp: tiles t = pair, pair + npairs, ...; four chunks each; chunk index g = 4 * i + cc
      // uses staging buffer g & 1 and its barrier; the load of chunk g + 2 is issued once chunk g's store has read it.
      uint64_t* my_bar = res_bar + 2 * ew;
      const int my_tiles = pair < num_tiles ? (num_tiles - pair + npairs - 1) / npairs : 0;
      const int n_chunks = 4 * my_tiles;
      auto chunk_coords = [&](int g, int& m0, int& n) {
        const int t = pair + (g >> 2) * npairs;
        const int tile = p.rev ? num_tiles - 1 - t : t;
        m0 = (tile / p.num_n_blocks) * (2 * BM) + static_cast<int>(crank) * BM + q * 32;
        n = (tile % p.num_n_blocks) * BN + cpart * (BN / 4) + (g & 3) * 16;
      };
      auto issue_load = [&](int g) {   // lane 0 only
        int m0, n;
        chunk_coords(g, m0, n);
        if (n < p.N && m0 < p.M) {
          mbar_expect_tx(&my_bar[g & 1], CHUNK_BYTES);
          tma_load_3d(&tmR, &my_bar[g & 1], reinterpret_cast<void*>(sEpi + ew * Cfg<MODE>::EPI_PER_WARP +
                                                                    (g & 1) * CHUNK_BYTES),
                      n, m0, 0, kEvictNormal);
        }
      };
      if (lane == 0) {
        if (n_chunks > 0) issue_load(0);
        if (n_chunks > 1) issue_load(1);
      }
      uint32_t use[2] = {0, 0};   // completed uses of each staging buffer (barrier parity)
      for (int i = 0; i < my_tiles; ++i) {
        int m0, n0;
        chunk_coords(4 * i, m0, n0);
        const float* bias = (p.m_split > 0 && m0 >= p.m_split) ? p.bias2 : p.bias;
        const bool hi = f.split > 0 && m0 >= f.split;
        const int row = m0 + lane;
        const bool rowok = row < p.M;
        const float2* sr = hi ? f.st_res2 : f.st_res;
        const bool norm = sr != nullptr;
        const float* rg = hi ? f.res_g2 : f.res_g;
        const float* rb = hi ? f.res_b2 : f.res_b;
        float rmu = 0.f, rrs = 1.f;
        const int oparts = p.N >> 6;
        if (f.st_res != nullptr) {   // uniform over the kernel: the statistics warp runs exactly then
          mbar_wait(sfull_bar, sph);
          const float2 ms = sStats[q * 32 + lane];
          __syncwarp();
          if (lane == 0) mbar_arrive(sempty_bar);
          sph ^= 1;
          rmu = ms.x;
          rrs = ms.y;
        }
        float s_mean = 0.f, s_m2 = 0.f;
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>(as * BN + cpart * (BN / 4));
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int g = 4 * i + cc;
          const int n = n0 + cc * 16;
          const bool live = n < p.N && m0 < p.M;
          uint32_t o[16];
          tmem_ld16(t_row + cc * 16, o);
          tmem_ld_wait();
          if (cc == 3) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_addr(leader_tempty + as * 8);
          }
          if (live) {
            const int bsel = g & 1;
            const uint32_t buf = stile + bsel * CHUNK_BYTES;
            const uint32_t srow = buf + lane * 64;
            mbar_wait(&my_bar[bsel], use[bsel] & 1);
            ++use[bsel];
            float v[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 xr = lds128(srow + ((j ^ sw) << 4));
              float4 x = make_float4(__uint_as_float(xr.x), __uint_as_float(xr.y), __uint_as_float(xr.z),
                                     __uint_as_float(xr.w));
              if (norm) {
                const float4 gm = __ldg(reinterpret_cast<const float4*>(rg + n) + j);
                const float4 bb = __ldg(reinterpret_cast<const float4*>(rb + n) + j);
                x.x = (x.x - rmu) * rrs * gm.x + bb.x;
                x.y = (x.y - rmu) * rrs * gm.y + bb.y;
                x.z = (x.z - rmu) * rrs * gm.z + bb.z;
                x.w = (x.w - rmu) * rrs * gm.w + bb.w;
              }
              float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
              if (bias) b = __ldg(reinterpret_cast<const float4*>(bias + n) + j);
              v[4 * j] = __uint_as_float(o[4 * j]) + b.x + x.x;
              v[4 * j + 1] = __uint_as_float(o[4 * j + 1]) + b.y + x.y;
              v[4 * j + 2] = __uint_as_float(o[4 * j + 2]) + b.z + x.z;
              v[4 * j + 3] = __uint_as_float(o[4 * j + 3]) + b.w + x.w;
            }
            // statistics of this thread's 64 columns, 16 at a time (Chan)
            float cs = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) cs += v[j];
            const float cm = cs * (1.0f / 16.0f);
            float cm2 = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float dl = v[j] - cm;
              cm2 += dl * dl;
            }
            if (cc == 0) {
              s_mean = cm;
              s_m2 = cm2;
            } else {
              const float na = 16.0f * cc, nt = na + 16.0f;
              const float dl = cm - s_mean;
              s_mean += dl * (16.0f / nt);
              s_m2 += cm2 + dl * dl * (na * 16.0f / nt);
            }
            // fp32 sum back into the staging chunk (same swizzled places), 16-bit copy + statistics straight to global
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128(srow + ((j ^ sw) << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                     __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
            if (rowok) {
              uint4* q16 = reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(f.out16) + (size_t)row * p.N + n);
              q16[0] = make_uint4(pack_act(v[0], v[1], p.fp16), pack_act(v[2], v[3], p.fp16),
                                  pack_act(v[4], v[5], p.fp16), pack_act(v[6], v[7], p.fp16));
              q16[1] = make_uint4(pack_act(v[8], v[9], p.fp16), pack_act(v[10], v[11], p.fp16),
                                  pack_act(v[12], v[13], p.fp16), pack_act(v[14], v[15], p.fp16));
              if (cc == 3) {
                float2* so = hi ? f.st_out2 : f.st_out;
                so[(size_t)(n0 >> 6) * f.st_stride + row] = make_float2(s_mean, s_m2);
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&tmC, buf, n, m0, 0);
              bulk_commit();
              bulk_wait_read0();                      // the store has read the buffer: it may take the next residual chunk
              if (g + 2 < n_chunks) issue_load(g + 2);
            }
            __syncwarp();
          } else if (lane == 0 && g + 2 < n_chunks) {
            issue_load(g + 2);   // dead chunk (ragged edge): nothing was loaded or stored, keep the prefetch chain going
          }
        }
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
      if (lane == 0) bulk_wait0();
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 2 * BN);
  }
}

}  // namespace

// LayerNorm fold (GemmFold): dense rows only; validated here, launched on the CTA-pair grid of gemm2.cu.
int launch_gemm_2cta_fold(const GemmDesc& d, cudaStream_t st) {
  SPRC_REQUIRE(d.fold != nullptr, "gemm fold: no descriptor");
  const GemmFold& f = *d.fold;
  const bool prod = f.st_out != nullptr, cons = f.st_in != nullptr;
  SPRC_REQUIRE(prod != cons, "gemm fold: exactly one of st_in (consumer) / st_out (producer) must be set");
  SPRC_REQUIRE(!prod || f.split == 0 || ((f.st_res != nullptr) == (f.st_res2 != nullptr)),
               "gemm fold producer: with a split either both row ranges carry a normalised residual or neither");
  SPRC_REQUIRE(d.grp_rows == 0 && !d.out_col_block && d.N % 64 == 0 && f.split % 32 == 0 && d.K % 64 == 0,
               "gemm fold: dense rows, N %% 64 == 0, K %% 64 == 0 and split %% 32 == 0 needed (N=%d K=%d split=%d)", d.N,
               d.K, f.split);
  SPRC_REQUIRE(!cons || (f.c && d.bias && d.out_bf16 && !d.out_f32 && !d.residual && (!d.W2 || (f.c2 && d.bias2)) &&
                         (f.split == 0 || f.st_in2)),
               "gemm fold consumer: 16-bit output, c/d vectors for every weight set, st_in2 with a split");
  SPRC_REQUIRE(!prod || (d.ldc == d.N && d.out_f32 && !d.out_bf16 && !d.residual && f.resid && f.out16 &&
                         d.act == ACT_NONE && (f.split == 0 || f.st_out2) &&
                         (!f.st_res || (f.res_g && f.res_b)) && (!f.st_res2 || (f.res_g2 && f.res_b2))),
               "gemm fold producer: ldc = N, fp32 output, no TMA residual, resid/out16/statistics set");
  CUtensorMap tmA, tmB, tmB2, tmC, tmR;
  if (prod) {
    SPRC_TRY(make_tmap_any(&tmC, d.out_f32, 4, d.N, d.M, 1, d.ldc, (uint64_t)d.M * d.ldc, 16, 32, 1, 3, 64));
    SPRC_TRY(make_tmap_any(&tmR, f.resid, 4, d.N, d.M, 1, d.N, (uint64_t)d.M * d.N, 16, 32, 1, 3, 64));
  } else {
    SPRC_TRY(make_tmap_any(&tmC, d.out_bf16, 2, d.N, d.M, 1, d.ldc, (uint64_t)d.M * d.ldc, 32, 32, 1, 3, 64));
    tmR = tmC;
  }
  SPRC_TRY(make_tmap_any(&tmA, d.A, 2, d.K, d.M, 1, d.lda, (uint64_t)d.M * d.lda, BK, BM, 1, 3, 128));
  SPRC_TRY(make_tmap_any(&tmB, d.W, 2, d.K, d.N, 1, d.ldw, 0, BK, BN / 2, 1, 2, 128));
  SPRC_TRY(make_tmap_any(&tmB2, d.W2 ? d.W2 : d.W, 2, d.K, d.N, 1, d.ldw, 0, BK, BN / 2, 1, 2, 128));

  FoldParams p;
  p.M = d.M;
  p.N = d.N;
  p.K = d.K;
  p.num_m_pairs = (d.M + 2 * BM - 1) / (2 * BM);
  p.num_n_blocks = (d.N + BN - 1) / BN;
  p.num_k_blocks = (d.K + BK - 1) / BK;
  p.bias = d.bias;
  p.bias2 = d.bias2;
  p.act = d.act;
  p.fp16 = act_fp16();
  p.rev = next_sweep_reverse();
  p.m_split = d.W2 ? d.m_split : 0;
  p.f = f;
  if (p.f.st_stride <= 0) p.f.st_stride = d.M;
  SPRC_REQUIRE(p.f.st_stride >= d.M, "gemm fold: statistics plane stride %d < M %d", p.f.st_stride, d.M);

  const int tiles = p.num_m_pairs * p.num_n_blocks;
  int npairs = device_sm_count() / 2;
  if (npairs > tiles) npairs = tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * npairs);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = prod ? Cfg<2>::SMEM_TOTAL : Cfg<1>::SMEM_TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  static bool attr_set = false;
  if (!attr_set) {
    SPRC_CUDA(cudaFuncSetAttribute(gemm_fold_2cta_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   Cfg<1>::SMEM_TOTAL));
    SPRC_CUDA(cudaFuncSetAttribute(gemm_fold_2cta_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   Cfg<2>::SMEM_TOTAL));
    attr_set = true;
  }
  prof_begin(st);
  if (prod)
    SPRC_CUDA(cudaLaunchKernelEx(&cfg, gemm_fold_2cta_kernel<2>, tmA, tmB, tmB2, tmC, tmR, p));
  else
    SPRC_CUDA(cudaLaunchKernelEx(&cfg, gemm_fold_2cta_kernel<1>, tmA, tmB, tmB2, tmC, tmR, p));
  if (prof_enabled()) {
    char tag[56];
    snprintf(tag, sizeof(tag), "M%d N%d K%d fold-%s a%d%s", d.M, d.N, d.K, prod ? "producer" : "consumer", d.act,
             d.W2 ? " w2" : "");
    const double out_b = prod ? (double)d.M * d.N * (4.0 + 4.0 + 2.0) : (double)d.M * d.N * 2.0;
    prof_end(PROF_GEMM, 2.0 * d.M * (double)d.N * d.K, 2.0 * ((double)d.M * d.K + (double)d.N * d.K) + out_b, st, tag);
  }
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sprc
