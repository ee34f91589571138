// CTA-pair tcgen05 GEMM (cta_group::2):  C = epilogue(A[M,K] * W[N,K]^T), 256 x 256 output tiles per PAIR of CTAs.
//
// Why pairs (tools/microbench/mma_rate, profiles/r01c_*): a single-CTA 128x256x16 tcgen05.mma with both operands in
// shared memory takes 171 cycles against its 128-cycle floor, because the tensor core reads its operands from shared
// memory at ~78 B/clk and a 128x256 tile needs 12 KB per MMA; the same tile shape also costs 48 KB of TMA traffic per
// K block, which paces the K = 768 Q-Former GEMMs.  In a pair each CTA stages its own 128 rows of A but only HALF
// (128 rows) of the 256 W rows: 8 KB of operand reads per MMA and 32 KB of TMA traffic per K block and CTA for the
// same 128x256 accumulator per SM.
//
// Roles per CTA (576 threads, as gemm.cu):
//   warp 0      TMA producer of this CTA's A rows and W-row half; completion bytes of BOTH CTAs are counted on the
//               leader's (cluster rank 0) full barrier
//   warp 1      leader only: single-thread tcgen05.mma.cta_group::2 issuer (UMMA 256 x 256 x 16); tcgen05.commit
//               multicasts the "slot free" / "accumulator full" arrivals to the barriers of both CTAs
//   warps 2..17 epilogue of this CTA's 128 accumulator rows (identical to gemm.cu); "accumulator drained" arrivals
//               of both CTAs go to the leader's barrier
#include <stdio.h>
#include <stdlib.h>

#include "common.h"
#include "ops.h"
#include "ptx.cuh"

namespace sprc {

int make_tmap_any(CUtensorMap* tm, const void* ptr, int esz, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
                  uint64_t stride2, uint32_t b0, uint32_t b1, uint32_t b2, int rank, int swizzle_bytes);

namespace {

constexpr int BM = 128;           // rows per CTA (pair tile: 256)
constexpr int BN = 256;           // columns per pair tile; each CTA stages BN / 2 rows of W
constexpr int BK = 64;
constexpr int STAGES = 6;
constexpr int EPI_WARPS = 16;
constexpr int EPI_STAGE_BYTES = 32 * 64;
constexpr int THREADS = (2 + EPI_WARPS) * 32;
constexpr int A_BYTES = BM * BK * 2;          // 16 KB
constexpr int B_BYTES = (BN / 2) * BK * 2;    // 16 KB
constexpr int RING_BYTES = STAGES * (A_BYTES + B_BYTES);
constexpr int EPI_BYTES = EPI_WARPS * EPI_STAGE_BYTES;
constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
constexpr int SMEM_TOTAL = RING_BYTES + EPI_BYTES + BAR_BYTES + 1024;

struct Gemm2Params {
  int M, N, K;
  int num_m_pairs, num_n_blocks, num_k_blocks;
  int grp_rows, grp_stride, grp_shift;
  const float* bias;
  int act;
  int fp16;
  int out_is_f32;
  int accumulate;
  int rev;
  int col_block;   // 1: column-blocked 16-bit output (GemmDesc::out_col_block = 64)
  int m_split;     // > 0: pair tiles whose first row is >= m_split read tmB2 / bias2 (GemmDesc::W2)
  const float* bias2;
};

__global__ void __launch_bounds__(THREADS, 1)
gemm_bf16_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                              const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmC,
                              const Gemm2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* sEpi = smem + RING_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + RING_BYTES + EPI_BYTES);   // used in the leader only
  uint64_t* empty_bar = full_bar + STAGES;                                          // one per CTA
  uint64_t* tfull_bar = empty_bar + STAGES;                                         // one per CTA
  uint64_t* tempty_bar = tfull_bar + 2;                                             // used in the leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();   // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  const int num_tiles = p.num_m_pairs * p.num_n_blocks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.m_split > 0) tma_prefetch_desc(&tmB2);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);    // the leader's producer arms it; both CTAs' TMA bytes complete it
      mbar_init(&empty_bar[s], 1);   // one multicast commit per phase
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);               // one multicast commit per tile
      mbar_init(&tempty_bar[s], 2 * EPI_WARPS);  // epilogue warps of both CTAs
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised and its TMEM is allocated before anything is signalled
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
          if (p.grp_rows == 0)
            tma_load_3d_2cta(&tmA, bar, sA + stage * A_BYTES, kb * BK, m0, 0, kEvictNormal);
          else
            tma_load_3d_2cta(&tmA, bar, sA + stage * A_BYTES, kb * BK, 0, m0 / p.grp_rows, kEvictNormal);
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
        mbar_wait(&tempty_bar[as], aphase ^ 1);   // both CTAs' epilogues have drained this accumulator stage
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
            umma_commit_2cta(&empty_bar[stage]);   // frees the slot in both CTAs
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
  } else {
    // ===================== epilogue (warps 2..17; see gemm.cu for the staging scheme) =====================
    const int q = warp & 3;
    const int cpart = (warp - 2) >> 2;
    const uint32_t stile = smem_u32(sEpi) + (warp - 2) * EPI_STAGE_BYTES;
    const uint32_t srow = stile + lane * 64;
    const uint32_t sw = (lane >> 1) & 3;
    const bool out32 = p.out_is_f32 != 0;
    const int CH = out32 ? 16 : 32;
    const int nchunks = (BN / 4) / CH;
    const uint32_t leader_tempty = mapa_u32(smem_u32(tempty_bar), 0);
    int as = 0;
    uint32_t aphase = 0;
    for (int t = pair; t < num_tiles; t += npairs) {
      const int tile = p.rev ? num_tiles - 1 - t : t;
      const int m0 = (tile / p.num_n_blocks) * (2 * BM) + static_cast<int>(crank) * BM + q * 32;
      const int n0 = (tile % p.num_n_blocks) * BN + cpart * (BN / 4);
      const float* bias = (p.m_split > 0 && m0 >= p.m_split) ? p.bias2 : p.bias;
      int c1 = m0, c2 = 0;
      if (p.grp_rows > 0) {
        c2 = m0 >> p.grp_shift;
        c1 = p.grp_rows >= 32 ? (m0 & (p.grp_rows - 1)) : 0;
      }
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                             static_cast<uint32_t>(as * BN + cpart * (BN / 4));
#pragma unroll 1
      for (int cc = 0; cc < nchunks; ++cc) {
        const int n = n0 + cc * CH;
        const bool live = n < p.N && m0 < p.M;
        uint32_t o[16];
        if (out32) {
          tmem_ld16(t_row + cc * 16, o);
          tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
              if (bias) b = __ldg(reinterpret_cast<const float4*>(bias + n) + j);
              float v0 = __uint_as_float(o[4 * j]) + b.x, v1 = __uint_as_float(o[4 * j + 1]) + b.y;
              float v2 = __uint_as_float(o[4 * j + 2]) + b.z, v3 = __uint_as_float(o[4 * j + 3]) + b.w;
              if (p.act == ACT_GELU) {
                v0 = gelu_erf(v0), v1 = gelu_erf(v1), v2 = gelu_erf(v2), v3 = gelu_erf(v3);
              } else if (p.act == ACT_QUICKGELU) {
                v0 = quick_gelu(v0), v1 = quick_gelu(v1), v2 = quick_gelu(v2), v3 = quick_gelu(v3);
              }
              o[4 * j] = __float_as_uint(v0), o[4 * j + 1] = __float_as_uint(v1);
              o[4 * j + 2] = __float_as_uint(v2), o[4 * j + 3] = __float_as_uint(v3);
            }
          }
        } else {
          uint32_t r[32];
          tmem_ld32(t_row + cc * 32, r);
          tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
              if (bias) b = __ldg(reinterpret_cast<const float4*>(bias + n) + j);
              float v0 = __uint_as_float(r[4 * j]) + b.x, v1 = __uint_as_float(r[4 * j + 1]) + b.y;
              float v2 = __uint_as_float(r[4 * j + 2]) + b.z, v3 = __uint_as_float(r[4 * j + 3]) + b.w;
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
        }
        if (cc == nchunks - 1) {
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
            if (p.accumulate)
              tma_reduce_add_3d(&tmC, stile, n, c1, c2);
            else if (p.col_block)
              tma_store_3d(&tmC, stile, n & 63, c1, n >> 6);
            else
              tma_store_3d(&tmC, stile, n, c1, c2);
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
  cluster_sync_all();   // neither CTA retires (or frees TMEM) while the pair's MMAs / commits may still touch it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 2 * BN);
  }
}

}  // namespace

bool gemm_2cta_enabled() {
  static const bool on = [] {
    const char* e = getenv("SPRC_GEMM_2CTA");
    return !(e && e[0] == '0');
  }();
  return on;
}

// Caller (gemm_bf16_tcgen05) has validated the descriptor (N % 32 == 0); the last N block may be ragged.
int launch_gemm_2cta(const GemmDesc& d, cudaStream_t st) {
  CUtensorMap tmA, tmB, tmB2, tmC;
  {
    const bool f32 = d.out_f32 != nullptr;
    const void* out = f32 ? static_cast<const void*>(d.out_f32) : static_cast<const void*>(d.out_bf16);
    const int esz = f32 ? 4 : 2;
    const uint32_t ch = f32 ? 16 : 32;
    if (d.grp_rows > 0) {
      const uint32_t br = d.grp_rows < 32 ? d.grp_rows : 32;
      SPRC_TRY(make_tmap_any(&tmC, out, esz, d.N, d.grp_rows, d.M / d.grp_rows, d.ldc, (uint64_t)d.grp_stride * d.ldc,
                             ch, br, 32 / br, 3, 64));
    } else if (d.out_col_block) {
      SPRC_TRY(make_tmap_any(&tmC, out, esz, 64, d.M, d.N / 64, 64, (uint64_t)d.M * 64, ch, 32, 1, 3, 64));
    } else {
      SPRC_TRY(make_tmap_any(&tmC, out, esz, d.N, d.M, 1, d.ldc, (uint64_t)d.M * d.ldc, ch, 32, 1, 3, 64));
    }
  }
  if (d.grp_rows > 0) {
    const int groups = d.M / d.grp_rows;
    SPRC_TRY(make_tmap_any(&tmA, d.A, 2, d.K, d.grp_rows, groups, d.lda, (uint64_t)d.grp_stride * d.lda, BK, d.grp_rows,
                           BM / d.grp_rows, 3, 128));
  } else {
    SPRC_TRY(make_tmap_any(&tmA, d.A, 2, d.K, d.M, 1, d.lda, (uint64_t)d.M * d.lda, BK, BM, 1, 3, 128));
  }
  SPRC_TRY(make_tmap_any(&tmB, d.W, 2, d.K, d.N, 1, d.ldw, 0, BK, BN / 2, 1, 2, 128));
  SPRC_TRY(make_tmap_any(&tmB2, d.W2 ? d.W2 : d.W, 2, d.K, d.N, 1, d.ldw, 0, BK, BN / 2, 1, 2, 128));

  Gemm2Params p;
  p.M = d.M;
  p.N = d.N;
  p.K = d.K;
  p.num_m_pairs = (d.M + 2 * BM - 1) / (2 * BM);
  p.num_n_blocks = (d.N + BN - 1) / BN;
  p.num_k_blocks = (d.K + BK - 1) / BK;
  p.grp_rows = d.grp_rows;
  p.grp_stride = d.grp_stride;
  p.grp_shift = 0;
  while (d.grp_rows > 0 && (1 << p.grp_shift) < d.grp_rows) ++p.grp_shift;
  p.bias = d.bias;
  p.act = d.act;
  p.fp16 = act_fp16();
  p.out_is_f32 = d.out_f32 ? 1 : 0;
  p.accumulate = d.residual ? 1 : 0;
  p.rev = next_sweep_reverse();
  p.col_block = d.out_col_block ? 1 : 0;
  p.m_split = d.W2 ? d.m_split : 0;
  p.bias2 = d.bias2;

  const int tiles = p.num_m_pairs * p.num_n_blocks;
  int npairs = device_sm_count() / 2;
  if (npairs > tiles) npairs = tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * npairs);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_TOTAL;
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
    SPRC_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   SMEM_TOTAL));
    attr_set = true;
  }
  prof_begin(st);
  SPRC_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_2cta_kernel, tmA, tmB, tmB2, tmC, p));
  if (prof_enabled()) {
    char tag[56];
    snprintf(tag, sizeof(tag), "M%d N%d K%d g%d a%d r%d f%d 2cta%s", d.M, d.N, d.K, d.grp_rows, d.act,
             d.residual ? 1 : 0, d.out_f32 ? 1 : 0, d.W2 ? " w2" : "");
    prof_end(PROF_GEMM, 2.0 * d.M * (double)d.N * d.K,
             2.0 * ((double)d.M * d.K + (double)d.N * d.K) + (double)d.M * d.N * (d.out_f32 ? 4.0 : 2.0), st, tag);
  }
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sprc
