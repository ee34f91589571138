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
//   warps 2..17 epilogue (four warps per TMEM lane quarter, each draining a quarter of the columns): tcgen05.ld 32 lanes x 32 columns -> bias / GELU / QuickGELU / residual ->
//               128-bit global stores (fp32 residual stream or bf16 activations)
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
                         const GemmKernelParams p) {
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

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
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
    // Two phases per 16-column chunk.  Phase 1: tcgen05.ld gives each thread ONE ROW x 16 columns; it is
    // written to this warp's swizzled smem tile.  Phase 2: the warp re-reads the tile so that 4 consecutive
    // lanes hold one row's 64 contiguous bytes (8 rows per instruction) and applies bias / activation /
    // residual there - every global load and store then covers whole 32-byte sectors of consecutive
    // addresses.  (The first version stored row-per-thread, 32 different lines per instruction, and was
    // epilogue-bound at K = 768.)  Four warps per SMSP keep the GELU math off the critical path.
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int cpart = (warp - 2) >> 2;   // which quarter of the tile's columns this warp drains
    const uint32_t stile = smem_u32(sEpi) + (warp - 2) * EPI_STAGE_BYTES;
    const int prow = lane >> 2;          // phase 2: row within an 8-row group
    const int ppiece = lane & 3;         // phase 2: 16-byte piece (4 fp32 columns) of the 64-byte row
    constexpr int CPW = BN / 64;         // 16-column chunks per warp
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / p.num_n_blocks) * BM;
      const int n0 = (tile % p.num_n_blocks) * BN;
      // physical output rows of the 4 rows this lane serves in phase 2 (-1: out of range)
      long long orow[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + q * 32 + i * 8 + prow;
        long long r = m;
        if (p.grp_rows > 0) r = static_cast<long long>(m >> p.grp_shift) * p.grp_stride + (m & (p.grp_rows - 1));
        orow[i] = m < p.M ? r : -1;
      }
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN);
#pragma unroll 1
      for (int cc = 0; cc < CPW; ++cc) {
        const int c = cpart * CPW + cc;
        const int n = n0 + c * 16;                 // first column of the chunk (warp-uniform)
        const bool col_ok = n < p.N;
        const int ncol = n + ppiece * 4;           // this lane's 4 columns in phase 2
        // residual and bias do not depend on the accumulator: issue their loads first
        float4 res[4];
        float4 bia = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col_ok) {
          if (p.bias) bia = __ldg(reinterpret_cast<const float4*>(p.bias + ncol));
          if (p.residual) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (orow[i] >= 0)
                res[i] = *reinterpret_cast<const float4*>(p.residual + static_cast<size_t>(orow[i]) * p.ldc + ncol);
          }
        }
        uint32_t r[16];
        tmem_ld16(t_row + c * 16, r);
        tmem_ld_wait();
        // phase 1: row `lane`, piece j -> byte offset lane*64 + ((j ^ ((lane >> 1) & 3)) * 16)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts128(stile + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        __syncwarp();
        if (col_ok) {
          uint4 raws[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = i * 8 + prow;
            raws[i] = lds128(stile + row * 64 + ((ppiece ^ ((row >> 1) & 3)) << 4));
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 raw = raws[i];
            float v0 = __uint_as_float(raw.x) + bia.x, v1 = __uint_as_float(raw.y) + bia.y;
            float v2 = __uint_as_float(raw.z) + bia.z, v3 = __uint_as_float(raw.w) + bia.w;
            if (p.act == ACT_GELU) {
              v0 = gelu_erf(v0), v1 = gelu_erf(v1), v2 = gelu_erf(v2), v3 = gelu_erf(v3);
            } else if (p.act == ACT_QUICKGELU) {
              v0 = quick_gelu(v0), v1 = quick_gelu(v1), v2 = quick_gelu(v2), v3 = quick_gelu(v3);
            }
            if (orow[i] >= 0) {
              const size_t off = static_cast<size_t>(orow[i]) * p.ldc + ncol;
              if (p.residual) v0 += res[i].x, v1 += res[i].y, v2 += res[i].z, v3 += res[i].w;
              if (p.out_f32)
                *reinterpret_cast<float4*>(p.out_f32 + off) = make_float4(v0, v1, v2, v3);
              else
                *reinterpret_cast<uint2*>(p.out_bf16 + off) =
                    make_uint2(pack_act(v0, v1, p.fp16), pack_act(v2, v3, p.fp16));
            }
          }
        }
        __syncwarp();  // the tile is rewritten by the next chunk's phase 1
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
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

// bf16 tensor [d2][d1][d0] with d0 contiguous; strides in elements; box {b0,b1,b2}; 128B swizzle.
int make_tmap_bf16(CUtensorMap* tm, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
                   uint64_t stride2, uint32_t b0, uint32_t b1, uint32_t b2, int rank) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return set_error(-38, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1 * 2, stride2 * 2};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (strides[0] & 15) || (rank == 3 && (strides[1] & 15)))
    return set_error(-22, "TMA operand must be 16-byte aligned (ptr %p, pitches %llu/%llu B)", ptr,
                     (unsigned long long)strides[0], (unsigned long long)strides[1]);
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(-22, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

template <int BN, int STAGES>
static int launch_gemm(const GemmDesc& d, cudaStream_t st) {
  using L = GemmSmem<BN, STAGES>;
  CUtensorMap tmA, tmB;
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

  static bool attr_set = false;
  if (!attr_set) {
    SPRC_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, STAGES>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
  prof_begin(st);
  gemm_bf16_tcgen05_kernel<BN, STAGES><<<grid, GEMM_THREADS, L::TOTAL, st>>>(tmA, tmB, p);
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

int gemm_bf16_tcgen05(const GemmDesc& d, cudaStream_t st) {
  SPRC_REQUIRE(d.M > 0 && d.N > 0 && d.K > 0, "gemm: empty problem %dx%dx%d", d.M, d.N, d.K);
  SPRC_REQUIRE(d.N % 32 == 0, "gemm: N=%d must be a multiple of 32", d.N);
  SPRC_REQUIRE(d.K % 8 == 0 && d.lda % 8 == 0 && d.ldw % 8 == 0 && d.ldc % 8 == 0,
               "gemm: K/lda/ldw/ldc must be multiples of 8 (K=%d lda=%d ldw=%d ldc=%d)", d.K, d.lda, d.ldw, d.ldc);
  SPRC_REQUIRE((d.out_f32 != nullptr) != (d.out_bf16 != nullptr), "gemm: exactly one output pointer");
  SPRC_REQUIRE(d.grp_rows == 0 || (BM % d.grp_rows == 0 && d.M % d.grp_rows == 0 && d.grp_stride >= d.grp_rows),
               "gemm: grp_rows=%d must divide 128 and M=%d", d.grp_rows, d.M);
  const long long tiles256 = (long long)((d.M + BM - 1) / BM) * ((d.N + 255) / 256);
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
  const size_t off = (size_t)prow * p.ldc + n;
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
  dim3 grid((d.N + 127) / 128, d.M);
  gemm_simt_kernel<<<grid, 128, 0, st>>>(d.A, d.W, p, d.lda, d.ldw);
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sprc
