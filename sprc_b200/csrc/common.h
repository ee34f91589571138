// Host-side plumbing shared by all translation units of libsprc_b200: error reporting
// (thread-local message + negative errno-style codes, see include/sprc_b200.h), CUDA
// error checking and the descriptors the kernels' host launchers take.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace sprc {

typedef __nv_bfloat16 bf16;

int set_error(int code, const char* fmt, ...);
const char* last_error();

#define SPRC_CUDA(expr)                                                                               \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess)                                                                            \
      return ::sprc::set_error(-5, "%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                 \
                               cudaGetErrorString(_e));                                               \
  } while (0)

#define SPRC_TRY(expr)        \
  do {                        \
    int _rc = (expr);         \
    if (_rc != 0) return _rc; \
  } while (0)

#define SPRC_REQUIRE(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) return ::sprc::set_error(-22, __VA_ARGS__);      \
  } while (0)

int device_sm_count();

// Programmatic dependent launch (PDL): the kernel may be scheduled while its predecessor on the stream is
// still draining; every kernel launched this way executes griddepcontrol.wait (ptx.cuh: griddep_wait) before
// it touches global memory, so only its prologue (barrier init, TMEM alloc, descriptor prefetch, index math)
// overlaps the predecessor's tail.  SPRC_PDL=0 in the environment falls back to plain launches.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
// 16-bit activation / operand format of the whole library: 0 = bf16 (default), 1 = fp16
int act_fp16();
void set_act_fp16(int on);

// ------------------------------------------------------------------------------------------------
// GEMM:  C[M,N] = epilogue( A[M,K] (bf16, row pitch lda) * W[N,K]^T (bf16, row pitch ldw) )
// ------------------------------------------------------------------------------------------------
enum Act { ACT_NONE = 0, ACT_GELU = 1, ACT_QUICKGELU = 2 };

struct GemmFold;

struct GemmDesc {
  const bf16* A = nullptr;
  const bf16* W = nullptr;
  int M = 0, N = 0, K = 0;
  int lda = 0, ldw = 0;
  // Row grouping: logical row m lives at physical row (m / grp_rows) * grp_stride + (m % grp_rows)
  // of A (relative to the A pointer) and of the output/residual (relative to their pointers).
  // grp_rows == 0 means "no grouping" (dense rows).  Used for the Q-Former's "first 32 rows of
  // every 64-row sample" operands (Qformer.py:436,466 row slices) without any gather copies.
  int grp_rows = 0, grp_stride = 0;
  const float* bias = nullptr;      // [N] fp32
  const float* residual = nullptr;  // fp32, pitch ldc (may alias out_f32)
  float* out_f32 = nullptr;         // exactly one of out_f32 / out_bf16
  bf16* out_bf16 = nullptr;
  int ldc = 0;
  int act = ACT_NONE;
  // out_col_block = 64: the 16-bit output is written column-blocked, element (m, n) at
  // out + (n / 64) * (M * 64) + m * 64 + (n % 64)  (ldc ignored): every 64-column block - one attention head of a
  // packed K/V projection - becomes a contiguous [M, 64] matrix.  Dense rows only.
  int out_col_block = 0;
  // Two weight sets in one launch: rows [0, m_split) use (W, bias), rows [m_split, M) use (W2, bias2) — the fusion
  // pass's query rows -> *_query FFN, text rows -> text FFN (Qformer.py:455-468) as ONE grid instead of a full-size
  // launch plus a half-empty one.  m_split must be a multiple of 256 (one pair tile); dense rows; same N, K, ldw.
  const bf16* W2 = nullptr;
  const float* bias2 = nullptr;
  int m_split = 0;
  // LayerNorm folded into the neighbouring GEMMs (gemm2_fold.cu; see GemmFold below)
  const GemmFold* fold = nullptr;
};

// Post-LN sublayer  y = LN(s), s = dense(a) + x  (Qformer.py:291-295, 373-381) without a LayerNorm kernel: the
// residual stream holds the PRE-LN sums s (fp32 + a raw 16-bit copy) and per-row statistics; LN is applied where
// its output is consumed.  Row statistics are width / 64 partials (mean, M2) per row (12 for the Q-Former's 768) = one
// per 64-column slice an epilogue thread owns, stored PART-major ([part][M rows of the GEMM], so a warp's 32 rows are
// contiguous) and merged (Chan) by whoever reads them - deterministic, no atomics.
// The ViT's pre-LN blocks (eva_vit.py:173-176, clip_vit.py:132-139) use the same two forms with a raw residual.
//   CONSUMER GEMM (st_in != null): A = raw 16-bit s, W = W * diag(gamma) (16-bit), GemmDesc::bias = d = W beta + b,
//     c = row sums of the rounded folded weight:  out = act(rstd * (acc - mean * c) + d).
//   PRODUCER GEMM (st_out != null; ldc == N, fp32 out, GemmDesc::residual must be null): s' = acc + bias + r with
//     r = resid (already normalised, st_res == null) or (resid - mean) * rstd * res_g + res_b; writes s' (fp32,
//     may alias resid), its raw 16-bit copy out16 and the statistics of s'.
// Rows >= split (a multiple of 32; 0 = one range) take the *2 members: the fusion pass's query rows and text rows
// went through different LayerNorms (output_query / output) and may sit in different statistics buffers.
struct GemmFold {
  int split = 0;
  int st_stride = 0;   // rows per statistics plane (st[part * st_stride + row]); 0 = the M of the launch
  const float2* st_in = nullptr;
  const float2* st_in2 = nullptr;
  const float* c = nullptr;
  const float* c2 = nullptr;
  const float* resid = nullptr;
  const float2* st_res = nullptr;
  const float2* st_res2 = nullptr;
  const float* res_g = nullptr;
  const float* res_b = nullptr;
  const float* res_g2 = nullptr;
  const float* res_b2 = nullptr;
  float2* st_out = nullptr;
  float2* st_out2 = nullptr;
  bf16* out16 = nullptr;
  float eps = 1e-12f;
};
constexpr int kFoldParts = 12;   // statistics partials per Q-Former row of 768 (64 columns each)

int gemm_bf16_tcgen05(const GemmDesc& d, cudaStream_t st);  // the product path (UTCHMMA + TMA)
int gemm_bf16_simt(const GemmDesc& d, cudaStream_t st);     // CUDA-core checker used by tests only

}  // namespace sprc
