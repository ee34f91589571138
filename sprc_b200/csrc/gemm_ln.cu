// Fused  out = LayerNorm(A * W^T + bias + residual)  for the Q-Former's post-LN sublayers (N = 768):
//   BertSelfOutput / BertOutput (Qformer.py:291-295, 373-381): hidden = LayerNorm(dense(x) + input_tensor), eps 1e-12.
//
// Unfused, each sublayer moved 18 bytes per hidden element through HBM (GEMM: reduce-add into the fp32 residual
// stream = read 4 + write 4; LayerNorm kernel: read 4, write fp32 4 + 16-bit 2) and the 768-wide output GEMMs plus
// their LayerNorms were the HBM-bound part of the query step.  Here the GEMM epilogue owns complete rows, so the
// sum never leaves the SM before it is normalised: read residual 4, write fp32 4 + 16-bit 2 = 10 bytes per element.
//
// A 128-row x 768-column fp32 accumulator does not fit the 512 TMEM columns of one SM, so the three 256-column
// N blocks of an M tile run on the three CTAs of a thread-block CLUSTER and exchange per-row LayerNorm partials
// (mean, M2 of their 64-column slices) through distributed shared memory:
//   warp 0      TMA producer (A tile 128x64, W tile 256x64, 3-stage ring)          [as gemm.cu]
//   warp 1      tcgen05.mma issuer, accumulators double-buffered in TMEM            [as gemm.cu]
//   warps 2..17 epilogue, thread = (row, 64-column slice):
//     pass 1  tcgen05.ld -> v = acc + bias + residual -> tcgen05.st back to TMEM; slice (mean, M2) -> st.async into all
//             three CTAs' partial tables (each store completes 8 bytes on the destination CTA's transaction barrier)
//     pass 2  wait for the 12 x 128 partials -> Chan-merge the 12 partials of the row -> tcgen05.ld v -> normalise ->
//             fp32 (residual stream, in place) and 16-bit (next GEMM operand) through swizzled staging + TMA stores
#include <stdio.h>
#include <stdlib.h>

#include "common.h"
#include "ops.h"
#include "ptx.cuh"

namespace sprc {

int make_tmap_any(CUtensorMap* tm, const void* ptr, int esz, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
                  uint64_t stride2, uint32_t b0, uint32_t b1, uint32_t b2, int rank, int swizzle_bytes);

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int LN_N = 768;
constexpr int CL = LN_N / BN;            // cluster size = N blocks per row
constexpr int STAGES = 3;
constexpr int EPI_WARPS = 16;
constexpr int THREADS = (2 + EPI_WARPS) * 32;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
constexpr int RING_BYTES = STAGES * (A_BYTES + B_BYTES);
constexpr int EPI_STAGE_BYTES = 32 * 64;
constexpr int EPI_BYTES = EPI_WARPS * EPI_STAGE_BYTES;
constexpr int NPART = CL * 4;            // partials per row: 3 CTAs x 4 column slices
constexpr int STATS_BYTES = 2 * NPART * BM * 8;
constexpr int BAR_BYTES = 256;
constexpr int SMEM_TOTAL = RING_BYTES + EPI_BYTES + STATS_BYTES + BAR_BYTES + 1024;

struct LnParams {
  int M, K;
  int num_m_blocks, num_k_blocks;
  int grp_rows, grp_stride, grp_shift;
  const float* bias;
  const float* residual;   // fp32, pitch ldc, may alias the fp32 output
  const float* gamma;
  const float* beta;
  float eps;
  int ldc;
  int fp16;
  int rev;
};

__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Asynchronous store into a peer CTA's shared memory that completes 8 bytes of transaction count on the peer's
// mbarrier once the data has landed: data and "ready" signal travel together, so no fence is needed on either side
// (a release/acquire pair at cluster scope costs MEMBAR.ALL.GPU + an L1 invalidate per thread and tile).
__device__ __forceinline__ void st_async_f32x2(uint32_t cluster_addr, float a, float b, uint32_t cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(
                   cluster_addr),
               "f"(a), "f"(b), "r"(cluster_bar)
               : "memory");
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_ln768_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                          const __grid_constant__ CUtensorMap tmC32, const __grid_constant__ CUtensorMap tmC16,
                          const LnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* sEpi = smem + RING_BYTES;
  uint8_t* sStats = sEpi + EPI_BYTES;   // [2 buffers][NPART][BM] float2 (mean, M2)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sStats + STATS_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* stats_bar = tempty_bar + 2;   // [2]: transaction barriers of the two partial tables
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stats_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();      // N block of this CTA
  const int cid = static_cast<int>(cluster_id_x());
  const int ncl = static_cast<int>(num_clusters_x());

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC32);
    tma_prefetch_desc(&tmC16);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], EPI_WARPS);
      mbar_init(&stats_bar[s], 1);   // armed per tile with the byte count of the 12 x 128 partials
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // every CTA's stats barriers exist before any peer arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();
  griddep_launch();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cid; t < p.num_m_blocks; t += ncl) {
        const int mb = p.rev ? p.num_m_blocks - 1 - t : t;
        const int m0 = mb * BM;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
          if (p.grp_rows == 0)
            tma_load_3d(&tmA, &full_bar[stage], sA + stage * A_BYTES, kb * BK, m0, 0, kEvictNormal);
          else
            tma_load_3d(&tmA, &full_bar[stage], sA + stage * A_BYTES, kb * BK, 0, m0 / p.grp_rows, kEvictNormal);
          tma_load_2d(&tmB, &full_bar[stage], sB + stage * B_BYTES, kb * BK, static_cast<int>(crank) * BN, kEvictLast);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc_16(BM, BN, p.fp16);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = cid; t < p.num_m_blocks; t += ncl) {
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
          for (int k = 0; k < BK / 16; ++k) umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (kb == p.num_k_blocks - 1) umma_commit(&tfull_bar[as]);
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
  } else {
    // ===================== epilogue (warps 2..17) =====================
    const int q = warp & 3;              // TMEM lane quarter
    const int cpart = (warp - 2) >> 2;   // 64-column slice of the CTA's 256 columns
    const int row = q * 32 + lane;       // row inside the M tile = TMEM lane
    const int ncol0 = static_cast<int>(crank) * BN + cpart * 64;   // first global column of the slice
    const uint32_t stile = smem_u32(sEpi) + (warp - 2) * EPI_STAGE_BYTES;
    const uint32_t srow = stile + lane * 64;
    const uint32_t sw = (lane >> 1) & 3;
    const uint32_t stats_base = smem_u32(sStats);
    const uint32_t my_part = crank * 4 + cpart;
    uint32_t peer_stats[CL], peer_bar[CL];
#pragma unroll
    for (int c = 0; c < CL; ++c) {
      peer_stats[c] = mapa(stats_base, c);
      peer_bar[c] = mapa(smem_u32(stats_bar), c);
    }
    int as = 0;
    uint32_t aphase = 0;
    for (int t = cid; t < p.num_m_blocks; t += ncl) {
      const int mb = p.rev ? p.num_m_blocks - 1 - t : t;
      const int m0w = mb * BM + q * 32;   // first row of this warp's 32 rows
      const int m = m0w + lane;
      const bool row_ok = m < p.M;
      const bool live = m0w < p.M;        // warp-uniform
      long long prow = m;
      int c1 = m0w, c2 = 0;
      if (p.grp_rows > 0) {
        prow = static_cast<long long>(m / p.grp_rows) * p.grp_stride + (m % p.grp_rows);
        c2 = m0w >> p.grp_shift;
        c1 = p.grp_rows >= 32 ? (m0w & (p.grp_rows - 1)) : 0;
      }
      const float* res_row = p.residual + prow * p.ldc + ncol0;
      // residual slice of the first 32 columns: in flight while the accumulator is still being produced
      float4 r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        r[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_ok) r[j] = *(reinterpret_cast<const float4*>(res_row) + j);
      }
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN + cpart * 64);
      // ---- pass 1: v = acc + bias + residual (kept in TMEM), slice statistics ----
      float mean_c[2], m2_c[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t a[32];
        tmem_ld32(t_row + c * 32, a);
        tmem_ld_wait();
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + ncol0 + c * 32) + j);
          const float v0 = __uint_as_float(a[4 * j]) + b.x + r[j].x, v1 = __uint_as_float(a[4 * j + 1]) + b.y + r[j].y;
          const float v2 = __uint_as_float(a[4 * j + 2]) + b.z + r[j].z, v3 = __uint_as_float(a[4 * j + 3]) + b.w + r[j].w;
          a[4 * j] = __float_as_uint(v0), a[4 * j + 1] = __float_as_uint(v1);
          a[4 * j + 2] = __float_as_uint(v2), a[4 * j + 3] = __float_as_uint(v3);
          sum += (v0 + v1) + (v2 + v3);
        }
        tmem_st32(t_row + c * 32, a);
        if (c == 0) {
          // second residual slice: in flight during the statistics of the first
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (row_ok) r[j] = *(reinterpret_cast<const float4*>(res_row + 32) + j);
        }
        const float mu = sum * (1.f / 32.f);
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = __uint_as_float(a[j]) - mu;
          ss = fmaf(d, d, ss);
        }
        mean_c[c] = mu;
        m2_c[c] = ss;
      }
      tmem_st_wait();
      {
        // Chan merge of the two 32-column halves, then publish (mean, M2) of the 64-column slice to all CTAs
        const float dm = mean_c[0] - mean_c[1];
        const float mu64 = 0.5f * (mean_c[0] + mean_c[1]);
        const float m264 = m2_c[0] + m2_c[1] + 16.f * dm * dm;
        const uint32_t off = (static_cast<uint32_t>(as) * NPART + my_part) * (BM * 8) + row * 8;
        if (warp == 2 && lane == 0) mbar_expect_tx(&stats_bar[as], NPART * BM * 8);   // arm this tile's phase
#pragma unroll
        for (int c = 0; c < CL; ++c) st_async_f32x2(peer_stats[c] + off, mu64, m264, peer_bar[c] + as * 8);
      }
      // ---- row statistics from the 12 partials ----
      mbar_wait(&stats_bar[as], aphase);
      float mean, rstd;
      {
        float2 pt[NPART];
        float msum = 0.f;
#pragma unroll
        for (int i = 0; i < NPART; ++i) {
          pt[i] = lds_f32x2(stats_base + (static_cast<uint32_t>(as) * NPART + i) * (BM * 8) + row * 8);
          msum += pt[i].x;
        }
        mean = msum * (1.f / NPART);
        float m2 = 0.f;
#pragma unroll
        for (int i = 0; i < NPART; ++i) {
          const float d = pt[i].x - mean;
          m2 += pt[i].y + 64.f * d * d;
        }
        rstd = rsqrtf(m2 * (1.f / LN_N) + p.eps);
      }
      // ---- pass 2: normalise, write the fp32 residual stream and the 16-bit operand copy ----
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t a[32];
        tmem_ld32(t_row + c * 32, a);
        tmem_ld_wait();
        if (c == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[as]);   // the accumulator stage is free for tile t + 2
        }
        const int n = ncol0 + c * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + n) + j);
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.beta + n) + j);
          a[4 * j] = __float_as_uint((__uint_as_float(a[4 * j]) - mean) * rstd * g.x + b.x);
          a[4 * j + 1] = __float_as_uint((__uint_as_float(a[4 * j + 1]) - mean) * rstd * g.y + b.y);
          a[4 * j + 2] = __float_as_uint((__uint_as_float(a[4 * j + 2]) - mean) * rstd * g.z + b.z);
          a[4 * j + 3] = __float_as_uint((__uint_as_float(a[4 * j + 3]) - mean) * rstd * g.w + b.w);
        }
        if (live) {
          // two fp32 chunks of 16 columns, then one 16-bit chunk of 32 columns (64 bytes per row each)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128(srow + ((j ^ sw) << 4), a[16 * h + 4 * j], a[16 * h + 4 * j + 1], a[16 * h + 4 * j + 2],
                     a[16 * h + 4 * j + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&tmC32, stile, n + 16 * h, c1, c2);
              bulk_commit();
            }
          }
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(srow + ((j ^ sw) << 4),
                   pack_act(__uint_as_float(a[8 * j]), __uint_as_float(a[8 * j + 1]), p.fp16),
                   pack_act(__uint_as_float(a[8 * j + 2]), __uint_as_float(a[8 * j + 3]), p.fp16),
                   pack_act(__uint_as_float(a[8 * j + 4]), __uint_as_float(a[8 * j + 5]), p.fp16),
                   pack_act(__uint_as_float(a[8 * j + 6]), __uint_as_float(a[8 * j + 7]), p.fp16));
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&tmC16, stile, n, c1, c2);
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
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA retires while a peer may still write its partial table / arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

int max_clusters() {
  static int cached = 0;
  if (cached) return cached;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL * 64);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_TOTAL;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gemm_ln768_tcgen05_kernel, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = device_sm_count() / CL - 4;   // conservative: GPC sizes need not be multiples of the cluster size
  }
  cached = n;
  if (getenv("SPRC_DEBUG")) fprintf(stderr, "[sprc] gemm_ln: %d active clusters of %d CTAs\n", n, CL);
  return n;
}

}  // namespace

int gemm_ln_tcgen05(const GemmDesc& d, const float* gamma, const float* beta, float eps, bf16* out_ln16,
                    cudaStream_t st) {
  SPRC_REQUIRE(d.M > 0 && d.K > 0, "gemm_ln: empty problem %dx%dx%d", d.M, d.N, d.K);
  SPRC_REQUIRE(d.N == LN_N, "gemm_ln: N=%d (the fused LayerNorm epilogue is built for 768-wide rows)", d.N);
  SPRC_REQUIRE(d.K % 8 == 0 && d.lda % 8 == 0 && d.ldw % 8 == 0 && d.ldc % 8 == 0,
               "gemm_ln: K/lda/ldw/ldc must be multiples of 8 (K=%d lda=%d ldw=%d ldc=%d)", d.K, d.lda, d.ldw, d.ldc);
  SPRC_REQUIRE(d.out_f32 && out_ln16 && d.residual && d.bias && gamma && beta && d.act == ACT_NONE,
               "gemm_ln: needs bias, residual, gamma, beta, an fp32 and a 16-bit output, no activation");
  SPRC_REQUIRE((reinterpret_cast<uintptr_t>(d.residual) & 15) == 0, "gemm_ln: residual must be 16-byte aligned");
  SPRC_REQUIRE(d.grp_rows == 0 || (BM % d.grp_rows == 0 && d.M % d.grp_rows == 0 && d.grp_stride >= d.grp_rows),
               "gemm_ln: grp_rows=%d must divide 128 and M=%d", d.grp_rows, d.M);
  static bool attr_set = false;
  if (!attr_set) {
    SPRC_CUDA(cudaFuncSetAttribute(gemm_ln768_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    attr_set = true;
  }
  CUtensorMap tmA, tmB, tmC32, tmC16;
  if (d.grp_rows > 0) {
    const int groups = d.M / d.grp_rows;
    const uint32_t br = d.grp_rows < 32 ? d.grp_rows : 32;
    SPRC_TRY(make_tmap_any(&tmA, d.A, 2, d.K, d.grp_rows, groups, d.lda, (uint64_t)d.grp_stride * d.lda, BK,
                           d.grp_rows, BM / d.grp_rows, 3, 128));
    SPRC_TRY(make_tmap_any(&tmC32, d.out_f32, 4, d.N, d.grp_rows, groups, d.ldc, (uint64_t)d.grp_stride * d.ldc, 16, br,
                           32 / br, 3, 64));
    SPRC_TRY(make_tmap_any(&tmC16, out_ln16, 2, d.N, d.grp_rows, groups, d.ldc, (uint64_t)d.grp_stride * d.ldc, 32, br,
                           32 / br, 3, 64));
  } else {
    SPRC_TRY(make_tmap_any(&tmA, d.A, 2, d.K, d.M, 1, d.lda, (uint64_t)d.M * d.lda, BK, BM, 1, 3, 128));
    SPRC_TRY(make_tmap_any(&tmC32, d.out_f32, 4, d.N, d.M, 1, d.ldc, (uint64_t)d.M * d.ldc, 16, 32, 1, 3, 64));
    SPRC_TRY(make_tmap_any(&tmC16, out_ln16, 2, d.N, d.M, 1, d.ldc, (uint64_t)d.M * d.ldc, 32, 32, 1, 3, 64));
  }
  SPRC_TRY(make_tmap_any(&tmB, d.W, 2, d.K, d.N, 1, d.ldw, 0, BK, BN, 1, 2, 128));

  LnParams p;
  p.M = d.M;
  p.K = d.K;
  p.num_m_blocks = (d.M + BM - 1) / BM;
  p.num_k_blocks = (d.K + BK - 1) / BK;
  p.grp_rows = d.grp_rows;
  p.grp_stride = d.grp_stride;
  p.grp_shift = 0;
  while (d.grp_rows > 0 && (1 << p.grp_shift) < d.grp_rows) ++p.grp_shift;
  p.bias = d.bias;
  p.residual = d.residual;
  p.gamma = gamma;
  p.beta = beta;
  p.eps = eps;
  p.ldc = d.ldc;
  p.fp16 = act_fp16();
  p.rev = next_sweep_reverse();

  int ncl = max_clusters();
  if (ncl > p.num_m_blocks) ncl = p.num_m_blocks;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL * ncl);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  prof_begin(st);
  SPRC_CUDA(cudaLaunchKernelEx(&cfg, gemm_ln768_tcgen05_kernel, tmA, tmB, tmC32, tmC16, p));
  if (prof_enabled()) {
    char tag[56];
    snprintf(tag, sizeof(tag), "M%d N%d K%d g%d +LN", d.M, d.N, d.K, d.grp_rows);
    prof_end(PROF_GEMM, 2.0 * d.M * (double)d.N * d.K,
             2.0 * ((double)d.M * d.K + (double)d.N * d.K) + (double)d.M * d.N * 10.0, st, tag);
  }
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sprc
