// tcgen05 GEMM for sm_100a:  C = epilogue(A[M,K] * W[N,K]^T), bf16 operands, fp32 accumulate in TMEM.
//
// This one kernel carries every dense contraction of the hot path (SURVEY.md §2.3 K1,K3,K5,K6,K7,
// K10-K13): ViT patch-embed / QKV / proj / fc1 / fc2 (eva_vit.py:55-59,122-146, clip_vit.py:118-122),
// the Q-Former's linear layers (Qformer.py:133-139,287,358,373) and the ITC heads
// (blip2_qformer_cir_align_prompt.py:80-81).
//
// Structure (persistent, one CTA per SM, 192 threads):
//   warp 0      TMA producer: A tile 128x64 and W tile BNx64 (bf16, 128B swizzle) into a STAGES-deep ring
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16), accumulators
//               double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps tile i+1
//   warps 2..17 epilogue (four warps per TMEM lane quarter, each draining a quarter of the columns): tcgen05.ld
//               (thread = row) -> bias / GELU / QuickGELU -> 64-byte row pieces into a swizzled 2 KB staging tile ->
//               one TMA store per 32-row x 64-byte chunk (bf16 activations / fp32), or a TMA reduce-add into the
//               fp32 residual stream (the residual is never loaded by the SM)
#include <stdio.h>

#include "common.h"
#include "ops.h"
#include "ptx.cuh"

namespace sprc {

static constexpr int BM = 128;
static constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
static constexpr int EPI_WARPS = 16;                     // four warps per TMEM lane quarter (column quarters)
static constexpr int EPI_STAGE_BYTES = 32 * 64;          // per-warp transpose tile: 32 rows x 64 B, XOR-swizzled
static constexpr int GEMM_THREADS = (2 + EPI_WARPS) * 32;

struct GemmKernelParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  int a_grp_rows;  // 0: dense (coords k, m0, 0); else (k, 0, m0 / a_grp_rows)
  int grp_rows, grp_stride, grp_shift;
  const float* bias;
  const float* residual;
  float* out_f32;
  bf16* out_bf16;
  int ldc;
  int act;
  int fp16;  // operand / bf16-output format: 0 bf16, 1 fp16
  int out_is_f32;  // output element type of tmC
  int accumulate;  // 1: out_f32 += result (TMA reduce-add; the residual already lives in the output buffer)
  int rev;         // 1: sweep the tiles from the last M block down (see next_sweep_reverse)
  int col_block;   // 1: column-blocked 16-bit output (GemmDesc::out_col_block = 64)
};

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int RING_BYTES = STAGES * (A_BYTES + B_BYTES);
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int EPI_BYTES = EPI_WARPS * EPI_STAGE_BYTES;
  static constexpr int TOTAL = RING_BYTES + EPI_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmC, const GemmKernelParams p) {
  using L = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * L::A_BYTES;
  uint8_t* sEpi = smem + L::RING_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::RING_BYTES + L::EPI_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();    // operands / residual are written by the preceding kernels on the stream
  griddep_launch();  // the next kernel may set itself up on this SM as soon as this CTA retires

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int tile = p.rev ? num_tiles - 1 - t : t;
        const int m0 = (tile / p.num_n_blocks) * BM;
        const int n0 = (tile % p.num_n_blocks) * BN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], L::A_BYTES + L::B_BYTES);
          if (p.a_grp_rows == 0)
            tma_load_3d(&tmA, &full_bar[stage], sA + stage * L::A_BYTES, kb * BK, m0, 0, kEvictNormal);
          else
            tma_load_3d(&tmA, &full_bar[stage], sA + stage * L::A_BYTES, kb * BK, 0, m0 / p.a_grp_rows,
                        kEvictNormal);
          tma_load_2d(&tmB, &full_bar[stage], sB + stage * L::B_BYTES, kb * BK, n0, kEvictLast);
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
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[as], aphase ^ 1);  // epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = umma_desc_k_sw128(smem_u32(sA + stage * L::A_BYTES));
          const uint64_t db = umma_desc_k_sw128(smem_u32(sB + stage * L::B_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
            umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
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
    // tcgen05.ld (32x32b) gives each thread ONE ROW x CH consecutive columns.  Bias and activation are applied
    // there, the row's 64 output bytes (16 fp32 / 32 bf16 columns) go to this warp's staging tile with the
    // SWIZZLE_64B pattern (16-byte piece j of row r at r*64 + ((j ^ (r >> 1 & 3)) << 4): conflict-free STS.128)
    // and lane 0 hands the 32-row x 64-byte tile to the TMA: a plain store, or a reduce-add into the fp32
    // residual stream.  Shared-memory traffic per output byte is one write + one TMA read; the earlier
    // STS/LDS/STG transpose cost more of the shared-memory pipe than the UMMA operand reads at K = 768 and
    // paced those GEMMs (profiles/r01_*: 7-8 us per 128x256 tile against 5.8 us of MMA).
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int cpart = (warp - 2) >> 2;   // which quarter of the tile's columns this warp drains
    const uint32_t stile = smem_u32(sEpi) + (warp - 2) * EPI_STAGE_BYTES;
    const uint32_t srow = stile + lane * 64;
    const uint32_t sw = (lane >> 1) & 3;
    const bool out32 = p.out_is_f32 != 0;
    const int CH = out32 ? 16 : 32;                  // columns per chunk (64 bytes of output per row)
    const int nchunks = (BN / 4) / CH;               // chunks per warp per tile
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int tile = p.rev ? num_tiles - 1 - t : t;
      const int m0 = (tile / p.num_n_blocks) * BM + q * 32;   // first row of this warp's 32 rows
      const int n0 = (tile % p.num_n_blocks) * BN + cpart * (BN / 4);
      // TMA coordinates of the warp's rows (dense: row m0; grouped: see GemmDesc)
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
        const int n = n0 + cc * CH;                  // first column of the chunk (warp-uniform)
        const bool live = n < p.N && m0 < p.M;
        uint32_t o[16];                              // the row's 64 output bytes
        if (out32) {
          tmem_ld16(t_row + cc * 16, o);
          tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.bias) b = __ldg(reinterpret_cast<const float4*>(p.bias + n) + j);
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
              if (p.bias) b = __ldg(reinterpret_cast<const float4*>(p.bias + n) + j);
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
          // the accumulator stage is in registers: hand it back to the MMA warp before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
        if (live) {
          if (lane == 0) bulk_wait_read0();  // the previous chunk's TMA store has read the staging tile
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
    if (lane == 0) bulk_wait0();  // stores performed before the CTA (and its shared memory) retires
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// Tensor [d2][d1][d0] with d0 contiguous, element size esz (2: bf16/fp16, 4: fp32); strides in elements;
// box {b0,b1,b2}; swizzle_bytes 128 (operand tiles) or 64 (epilogue staging tiles).
static int make_tmap(CUtensorMap* tm, const void* ptr, int esz, uint64_t d0, uint64_t d1, uint64_t d2,
                     uint64_t stride1, uint64_t stride2, uint32_t b0, uint32_t b1, uint32_t b2, int rank,
                     int swizzle_bytes) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return set_error(-38, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1 * esz, stride2 * esz};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (strides[0] & 15) || (rank == 3 && (strides[1] & 15)))
    return set_error(-22, "TMA operand must be 16-byte aligned (ptr %p, pitches %llu/%llu B)", ptr,
                     (unsigned long long)strides[0], (unsigned long long)strides[1]);
  CUresult r = enc(tm, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(-22, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

int make_tmap_any(CUtensorMap* tm, const void* ptr, int esz, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
                  uint64_t stride2, uint32_t b0, uint32_t b1, uint32_t b2, int rank, int swizzle_bytes) {
  return make_tmap(tm, ptr, esz, d0, d1, d2, stride1, stride2, b0, b1, b2, rank, swizzle_bytes);
}

int make_tmap_bf16(CUtensorMap* tm, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
                   uint64_t stride2, uint32_t b0, uint32_t b1, uint32_t b2, int rank) {
  return make_tmap(tm, ptr, 2, d0, d1, d2, stride1, stride2, b0, b1, b2, rank, 128);
}

// out[row, 0:N] = src[row, 0:N] over the (possibly grouped) rows of a GemmDesc: used only when the residual
// does not already live in the output buffer (the model always accumulates in place).
__global__ void __launch_bounds__(256)
copy_rows_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int M, int n4, int ld, int grp_rows,
                     int grp_stride) {
  const long long total = static_cast<long long>(M) * n4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(i / n4), c = static_cast<int>(i % n4);
    long long r = m;
    if (grp_rows > 0) r = static_cast<long long>(m / grp_rows) * grp_stride + (m % grp_rows);
    reinterpret_cast<float4*>(dst + r * ld)[c] = reinterpret_cast<const float4*>(src + r * ld)[c];
  }
}

template <int BN, int STAGES>
static int launch_gemm(const GemmDesc& d, cudaStream_t st) {
  using L = GemmSmem<BN, STAGES>;
  CUtensorMap tmA, tmB, tmC;
  {
    // output map: 64-byte row pieces (16 fp32 / 32 bf16 columns) x the 32 rows one epilogue warp owns
    const bool f32 = d.out_f32 != nullptr;
    const void* out = f32 ? static_cast<const void*>(d.out_f32) : static_cast<const void*>(d.out_bf16);
    const int esz = f32 ? 4 : 2;
    const uint32_t ch = f32 ? 16 : 32;
    if (d.grp_rows > 0) {
      const uint32_t br = d.grp_rows < 32 ? d.grp_rows : 32;
      SPRC_TRY(make_tmap(&tmC, out, esz, d.N, d.grp_rows, d.M / d.grp_rows, d.ldc, (uint64_t)d.grp_stride * d.ldc,
                         ch, br, 32 / br, 3, 64));
    } else if (d.out_col_block) {
      SPRC_TRY(make_tmap(&tmC, out, esz, 64, d.M, d.N / 64, 64, (uint64_t)d.M * 64, ch, 32, 1, 3, 64));
    } else {
      SPRC_TRY(make_tmap(&tmC, out, esz, d.N, d.M, 1, d.ldc, (uint64_t)d.M * d.ldc, ch, 32, 1, 3, 64));
    }
  }
  if (d.grp_rows > 0) {
    const int groups = d.M / d.grp_rows;
    SPRC_TRY(make_tmap_bf16(&tmA, d.A, d.K, d.grp_rows, groups, d.lda, (uint64_t)d.grp_stride * d.lda, BK,
                            d.grp_rows, BM / d.grp_rows, 3));
  } else {
    SPRC_TRY(make_tmap_bf16(&tmA, d.A, d.K, d.M, 1, d.lda, (uint64_t)d.M * d.lda, BK, BM, 1, 3));
  }
  SPRC_TRY(make_tmap_bf16(&tmB, d.W, d.K, d.N, 1, d.ldw, 0, BK, BN, 1, 2));

  GemmKernelParams p;
  p.M = d.M;
  p.N = d.N;
  p.K = d.K;
  p.num_m_blocks = (d.M + BM - 1) / BM;
  p.num_n_blocks = (d.N + BN - 1) / BN;
  p.num_k_blocks = (d.K + BK - 1) / BK;
  p.a_grp_rows = d.grp_rows;
  p.grp_rows = d.grp_rows;
  p.grp_stride = d.grp_stride;
  p.grp_shift = 0;
  while (d.grp_rows > 0 && (1 << p.grp_shift) < d.grp_rows) ++p.grp_shift;
  p.bias = d.bias;
  p.residual = d.residual;
  p.out_f32 = d.out_f32;
  p.out_bf16 = d.out_bf16;
  p.ldc = d.ldc;
  p.act = d.act;
  p.fp16 = act_fp16();
  p.out_is_f32 = d.out_f32 ? 1 : 0;
  p.accumulate = d.residual ? 1 : 0;
  p.rev = next_sweep_reverse();
  p.col_block = d.out_col_block ? 1 : 0;
  if (d.residual && d.residual != d.out_f32) {
    const long long total = (long long)d.M * (d.N / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    copy_rows_f32_kernel<<<(int)blocks, 256, 0, st>>>(d.residual, d.out_f32, d.M, d.N / 4, d.ldc, d.grp_rows,
                                                      d.grp_stride);
    count_launch();
  }

  static bool attr_set = false;
  if (!attr_set) {
    SPRC_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, STAGES>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
  prof_begin(st);
  SPRC_CUDA(launch_pdl(gemm_bf16_tcgen05_kernel<BN, STAGES>, dim3(grid), dim3(GEMM_THREADS), L::TOTAL, st, tmA, tmB,
                       tmC, p));
  if (prof_enabled()) {
    char tag[56];
    snprintf(tag, sizeof(tag), "M%d N%d K%d g%d a%d r%d f%d bn%d", d.M, d.N, d.K, d.grp_rows, d.act,
             d.residual ? 1 : 0, d.out_f32 ? 1 : 0, BN);
    prof_end(PROF_GEMM, 2.0 * d.M * (double)d.N * d.K,
             2.0 * ((double)d.M * d.K + (double)d.N * d.K) + (double)d.M * d.N * (d.out_f32 ? 4.0 : 2.0), st, tag);
  }
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

bool gemm_2cta_enabled();
int launch_gemm_2cta(const GemmDesc& d, cudaStream_t st);   // gemm2.cu

int gemm_bf16_tcgen05(const GemmDesc& d, cudaStream_t st) {
  SPRC_REQUIRE(d.M > 0 && d.N > 0 && d.K > 0, "gemm: empty problem %dx%dx%d", d.M, d.N, d.K);
  SPRC_REQUIRE(d.N % 32 == 0, "gemm: N=%d must be a multiple of 32", d.N);
  SPRC_REQUIRE(d.K % 8 == 0 && d.lda % 8 == 0 && d.ldw % 8 == 0 && d.ldc % 8 == 0,
               "gemm: K/lda/ldw/ldc must be multiples of 8 (K=%d lda=%d ldw=%d ldc=%d)", d.K, d.lda, d.ldw, d.ldc);
  SPRC_REQUIRE((d.out_f32 != nullptr) != (d.out_bf16 != nullptr), "gemm: exactly one output pointer");
  SPRC_REQUIRE(!d.residual || d.out_f32, "gemm: a residual needs the fp32 output");
  SPRC_REQUIRE(!d.residual || d.act == ACT_NONE, "gemm: activation with a residual is not supported");
  SPRC_REQUIRE(d.grp_rows == 0 || (BM % d.grp_rows == 0 && d.M % d.grp_rows == 0 && d.grp_stride >= d.grp_rows),
               "gemm: grp_rows=%d must divide 128 and M=%d", d.grp_rows, d.M);
  SPRC_REQUIRE(d.out_col_block == 0 || (d.out_col_block == 64 && d.out_bf16 && d.grp_rows == 0 && d.N % 64 == 0),
               "gemm: out_col_block needs 64, a 16-bit output, dense rows and N %% 64 == 0");
  if (d.W2) {
    SPRC_REQUIRE(d.m_split > 0 && d.m_split < d.M && d.m_split % 256 == 0 && d.grp_rows == 0 && !d.out_col_block,
                 "gemm: two weight sets need dense rows and 0 < m_split (%d) < M (%d), m_split %% 256 == 0", d.m_split,
                 d.M);
  }
  const long long tiles256 = (long long)((d.M + BM - 1) / BM) * ((d.N + 255) / 256);
  // CTA pairs (256 x 256 tiles) whenever every pair gets work; the single-CTA kernels cover ragged N and small problems
  // (a ragged last N block - ViT-g's 1408 = 5.5 x 256, 4224 = 16.5 x 256 - runs as a half-empty 256-column tile: its
  // missing W rows are zero-filled by the TMA and its stores are clipped, 3-9 % padded MMA work against the ~25 %
  // slower single-CTA kernel)
  if (gemm_2cta_enabled() && d.N >= 256 && tiles256 >= device_sm_count() && (!d.residual || d.residual == d.out_f32))
    return launch_gemm_2cta(d, st);
  if (d.W2) {   // the single-CTA kernels take one weight set: two launches over the two row ranges
    GemmDesc a = d, b = d;
    a.W2 = b.W2 = nullptr;
    a.bias2 = b.bias2 = nullptr;
    a.m_split = b.m_split = 0;
    a.M = d.m_split;
    b.M = d.M - d.m_split;
    b.A = d.A + (size_t)d.m_split * d.lda;
    b.W = d.W2;
    b.bias = d.bias2;
    if (d.residual) b.residual = d.residual + (size_t)d.m_split * d.ldc;
    if (d.out_f32) b.out_f32 = d.out_f32 + (size_t)d.m_split * d.ldc;
    if (d.out_bf16) b.out_bf16 = d.out_bf16 + (size_t)d.m_split * d.ldc;
    SPRC_TRY(gemm_bf16_tcgen05(a, st));
    return gemm_bf16_tcgen05(b, st);
  }
  if (d.N % 256 == 0 && tiles256 >= device_sm_count()) return launch_gemm<256, 4>(d, st);
  return launch_gemm<128, 6>(d, st);
}

// ------------------------------------------------------------------------------------------------
// CUDA-core checker (tests only): one thread per output element, same epilogue semantics.
// ------------------------------------------------------------------------------------------------
__global__ void gemm_simt_kernel(const bf16* __restrict__ A, const bf16* __restrict__ W, GemmKernelParams p,
                                 int lda, int ldw) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= p.N || m >= p.M) return;
  long long prow = m;
  if (p.grp_rows > 0) prow = (long long)(m / p.grp_rows) * p.grp_stride + (m % p.grp_rows);
  const bf16* a = A + prow * lda;
  const bf16* w = W + (size_t)n * ldw;
  float acc = 0.f;
  const unsigned short* au = reinterpret_cast<const unsigned short*>(a);
  const unsigned short* wu = reinterpret_cast<const unsigned short*>(w);
  for (int k = 0; k < p.K; ++k) acc += from_act(au[k], p.fp16) * from_act(wu[k], p.fp16);
  if (p.bias) acc += p.bias[n];
  if (p.act == ACT_GELU) acc = gelu_erf(acc);
  if (p.act == ACT_QUICKGELU) acc = quick_gelu(acc);
  size_t off = (size_t)prow * p.ldc + n;
  if (p.col_block) off = (size_t)(n >> 6) * ((size_t)p.M * 64) + (size_t)m * 64 + (n & 63);
  if (p.residual) acc += p.residual[off];
  if (p.out_f32)
    p.out_f32[off] = acc;
  else
    reinterpret_cast<unsigned short*>(p.out_bf16)[off] = to_act(acc, p.fp16);
}

int gemm_bf16_simt(const GemmDesc& d, cudaStream_t st) {
  GemmKernelParams p = {};
  p.M = d.M;
  p.N = d.N;
  p.K = d.K;
  p.grp_rows = d.grp_rows;
  p.grp_stride = d.grp_stride;
  p.bias = d.bias;
  p.residual = d.residual;
  p.out_f32 = d.out_f32;
  p.out_bf16 = d.out_bf16;
  p.ldc = d.ldc;
  p.act = d.act;
  p.fp16 = act_fp16();
  p.col_block = d.out_col_block ? 1 : 0;
  dim3 grid((d.N + 127) / 128, d.M);
  gemm_simt_kernel<<<grid, 128, 0, st>>>(d.A, d.W, p, d.lda, d.ldw);
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sprc
