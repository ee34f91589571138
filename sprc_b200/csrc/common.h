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
};


int gemm_bf16_tcgen05(const GemmDesc& d, cudaStream_t st);  // the product path (UTCHMMA + TMA)
int gemm_bf16_simt(const GemmDesc& d, cudaStream_t st);     // CUDA-core checker used by tests only

}  // namespace sprc
