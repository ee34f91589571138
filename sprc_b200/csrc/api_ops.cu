// C-ABI: library-level queries and the single-op entry points declared at the bottom of
// include/sprc_b200.h (used by tests/ and micro-benchmarks to check each kernel in isolation).
#include "../../include/sprc_b200.h"
#include "common.h"
#include "ops.h"

using namespace sprc;

extern "C" {

int sprc_abi_version(void) { return SPRC_ABI_VERSION; }
const char* sprc_last_error(void) { return last_error(); }
int64_t sprc_launch_count(void) { return launch_count(); }
int sprc_set_act_dtype(int fp16) {
  if (fp16 != 0 && fp16 != 1) return set_error(-22, "sprc_set_act_dtype: 0 (bf16) or 1 (fp16)");
  set_act_fp16(fp16);
  return 0;
}
int sprc_profile(int enable) {
  prof_set(enable != 0);
  return 0;
}
int sprc_profile_dump(const char* path) {
  if (!path) return set_error(-22, "sprc_profile_dump: null path");
  return prof_dump(path);
}
int sprc_profile_read(double* out, int ncat) {
  if (!out || ncat <= 0) return set_error(-22, "sprc_profile_read: bad arguments");
  return prof_read(out, ncat);
}

int sprc_op_gemm(const void* A, const void* W, int M, int N, int K, int lda, int ldw, int grp_rows, int grp_stride,
                 const float* bias, const float* residual, float* out_f32, void* out_bf16, int ldc, int act,
                 int impl, void* stream) {
  GemmDesc d;
  d.A = static_cast<const bf16*>(A);
  d.W = static_cast<const bf16*>(W);
  d.M = M;
  d.N = N;
  d.K = K;
  d.lda = lda;
  d.ldw = ldw;
  d.grp_rows = grp_rows;
  d.grp_stride = grp_stride;
  d.bias = bias;
  d.residual = residual;
  d.out_f32 = out_f32;
  d.out_bf16 = static_cast<bf16*>(out_bf16);
  d.ldc = ldc;
  d.act = act;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return impl == 0 ? gemm_bf16_tcgen05(d, st) : gemm_bf16_simt(d, st);
}

int sprc_op_gemm2w(const void* A, const void* W, const void* W2, int M, int m_split, int N, int K, const float* bias,
                   const float* bias2, const float* residual, float* out_f32, void* out_bf16, int act, void* stream) {
  GemmDesc d;
  d.A = static_cast<const bf16*>(A);
  d.W = static_cast<const bf16*>(W);
  d.W2 = static_cast<const bf16*>(W2);
  d.M = M;
  d.m_split = m_split;
  d.N = N;
  d.K = K;
  d.lda = K;
  d.ldw = K;
  d.bias = bias;
  d.bias2 = bias2;
  d.residual = residual;
  d.out_f32 = out_f32;
  d.out_bf16 = static_cast<bf16*>(out_bf16);
  d.ldc = N;
  d.act = act;
  return gemm_bf16_tcgen05(d, static_cast<cudaStream_t>(stream));
}

int sprc_op_attention_pairs(const void* Q, const void* K, const void* V, void* O, int B, int H, int ldq, int ldk, int ldv,
                            int ldo, int q_batch_rows, const int32_t* kv_idx0, const int32_t* kv_idx1,
                            int64_t kv_rows_total, int64_t kv_head_stride, float scale, void* stream) {
  if (!Q || !K || !V || !O || !kv_idx0 || !kv_idx1) return set_error(-22, "sprc_op_attention_pairs: null argument");
  AttnDesc a;
  a.Q = static_cast<const bf16*>(Q);
  a.K = static_cast<const bf16*>(K);
  a.V = static_cast<const bf16*>(V);
  a.O = static_cast<bf16*>(O);
  a.B = B;
  a.H = H;
  a.dh = 64;
  a.Lq = 32;
  a.Lk = 514;
  a.Lk1 = 257;
  a.ldq = ldq;
  a.ldk = ldk;
  a.ldv = ldv;
  a.ldo = ldo;
  a.q_batch_rows = q_batch_rows;
  a.kv_batch_rows = 257;
  a.kv_idx0 = kv_idx0;
  a.kv_idx1 = kv_idx1;
  a.kv_rows_total = kv_rows_total;
  a.kv_head_stride = kv_head_stride;
  a.scale = scale;
  return attention(a, static_cast<cudaStream_t>(stream));
}

int sprc_preprocess_targetpad(const uint8_t* pixels, const int64_t* desc, const int32_t* tables, int n, int dim,
                              int max_rows, uint8_t* tmp, const float* mean3, const float* std3, float* out,
                              void* stream) {
  return preprocess_targetpad(pixels, reinterpret_cast<const long long*>(desc), tables, n, dim, max_rows, tmp, mean3,
                              std3, out, static_cast<cudaStream_t>(stream));
}

int sprc_op_attention_ragged(const void* qkv, int ldqkv, void* out, int ldo, int B, int rows_total,
                             const int32_t* pairs_dev, float scale, void* stream) {
  return attention_qf_ragged(static_cast<const bf16*>(qkv), ldqkv, static_cast<bf16*>(out), ldo, B, rows_total,
                             reinterpret_cast<const int4*>(pairs_dev), scale, static_cast<cudaStream_t>(stream));
}

int sprc_op_layernorm(const float* x, int rows, int width, const float* gamma, const float* beta, float eps,
                      int grp_rows, int grp_stride, float* out_f32, void* out_bf16, void* stream) {
  return layernorm(x, rows, width, gamma, beta, eps, grp_rows, grp_stride, out_f32, static_cast<bf16*>(out_bf16),
                   static_cast<cudaStream_t>(stream));
}

int sprc_op_attention(const void* Q, const void* K, const void* V, void* O, int B, int H, int dh, int Lq, int Lk,
                      int ldq, int ldk, int ldv, int ldo, int q_batch_rows, int kv_batch_rows,
                      const float* key_mask, float scale, void* stream) {
  AttnDesc a;
  a.Q = static_cast<const bf16*>(Q);
  a.K = static_cast<const bf16*>(K);
  a.V = static_cast<const bf16*>(V);
  a.O = static_cast<bf16*>(O);
  a.B = B;
  a.H = H;
  a.dh = dh;
  a.Lq = Lq;
  a.Lk = Lk;
  a.ldq = ldq;
  a.ldk = ldk;
  a.ldv = ldv;
  a.ldo = ldo;
  a.q_batch_rows = q_batch_rows;
  a.kv_batch_rows = kv_batch_rows;
  a.key_mask = key_mask;
  a.scale = scale;
  return attention(a, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
