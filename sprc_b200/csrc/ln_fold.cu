// LayerNorm folded into the neighbouring GEMMs (GemmFold, common.h; epilogues in gemm2_fold.cu): folded weights.
//
// Reference arithmetic (lavis/models/blip2_models/Qformer.py:291-295 BertSelfOutput, :373-381 BertOutput):
//     y = LayerNorm(dense(a) + x) ;  the next sublayer reads y twice: as the A operand of its GEMMs and as its residual.
//     producer GEMM   s' = acc + b + LN(s)        writes s' fp32 + raw 16-bit copy + (mean, M2) partials per row
//     consumer GEMM   act(rstd (s16 (W diag(g))^T - mean c) + d),  c = rowsum(W diag(g)),  d = W beta + b
// i.e. LN(s) W^T + b with the per-row scalars pulled out of the contraction.  c is summed over the ROUNDED 16-bit folded
// weight, so `acc - mean c` is exactly sum_k (s16_k - mean) Wf_nk in the tensor core's own operands.
//
// Schedule (SPRC_LN_FOLD=1): layers 0 .. L-2 of the ragged composed-query passes, of the rerank pairs and of the gallery
// pass run without LayerNorm kernels; the last layer runs the default schedule on a materialised stream (its
// row-restricted outputs and the [CLS] gather stay as they are).  The ViT's pre-LN blocks use the same two epilogues
// with a raw residual (Model::vit_blocks_fold).
#include <math.h>
#include <stdlib.h>

#include "model.h"
#include "ptx.cuh"

namespace sprc {

bool ln_fold_enabled() {
  static const bool on = [] {
    const char* e = getenv("SPRC_LN_FOLD");
    return e && e[0] == '1';
  }();
  return on;
}

// One warp per output row n:  Wf[n,k] = round16(W[n,k] * gamma[k]);  c[n] = sum_k Wf[n,k];
// d[n] = sum_k W[n,k] * beta[k] + bias[n].
__global__ void __launch_bounds__(256)
fold_weight_kernel(const unsigned short* __restrict__ W, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ bias, int N, int K,
                   unsigned short* __restrict__ Wf, float* __restrict__ c, float* __restrict__ d, int fp16) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  float cs = 0.f, ds = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = from_act(W[(size_t)n * K + k], fp16);
    const unsigned short wf = to_act(w * gamma[k], fp16);
    Wf[(size_t)n * K + k] = wf;
    cs += from_act(wf, fp16);
    ds += w * beta[k];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cs += __shfl_xor_sync(0xffffffffu, cs, o);
    ds += __shfl_xor_sync(0xffffffffu, ds, o);
  }
  if (lane == 0) {
    c[n] = cs;
    d[n] = ds + (bias ? bias[n] : 0.f);
  }
}

int fold_weight(const bf16* W, const float* gamma, const float* beta, const float* bias, int N, int K, bf16* Wf,
                float* c, float* d, cudaStream_t st) {
  SPRC_REQUIRE(W && gamma && beta && Wf && c && d && N > 0 && K > 0, "fold_weight: bad arguments");
  fold_weight_kernel<<<(N + 7) / 8, 256, 0, st>>>(reinterpret_cast<const unsigned short*>(W), gamma, beta, bias, N, K,
                                                  reinterpret_cast<unsigned short*>(Wf), c, d, act_fp16());
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

static int plain_linear(const bf16* A, int M, int K, const bf16* W, int N, const float* bias, int act, bf16* out,
                        cudaStream_t st) {
  GemmDesc d;
  d.A = A;
  d.W = W;
  d.M = M;
  d.N = d.ldc = N;
  d.K = d.lda = d.ldw = K;
  d.bias = bias;
  d.act = act;
  d.out_bf16 = out;
  return gemm_bf16_tcgen05(d, st);
}

int Model::fold_one(FoldedLinear* f, const bf16* W, const float* bias, const float* gamma, const float* beta, int N,
                    cudaStream_t st, int K) {
  if (!f->w) {
    SPRC_TRY(alloc_t(&f->w, (size_t)N * K));
    SPRC_TRY(alloc_t(&f->c, N));
    SPRC_TRY(alloc_t(&f->d, N));
  }
  return fold_weight(W, gamma, beta, bias, N, K, f->w, f->c, f->d, st);
}

// Folded weights of every GEMM that reads a LayerNorm output in layers 0 .. L-2 of the ragged passes (re-derived after
// every sprc_load_weights), plus the two statistics buffers.
int Model::prepare_fold(cudaStream_t st) {
  if (fold_ready) return 0;
  if (!fold_st[0]) {
    SPRC_TRY(alloc_t(&fold_st[0], (size_t)qf_rows * kFoldParts));   // [12 parts][qf_rows] (mean, M2), part-major
    SPRC_TRY(alloc_t(&fold_st[1], (size_t)qf_rows * kFoldParts));
  }
  folds.resize(qf_layers);
  for (int l = 0; l < qf_layers; ++l) {
    const QfLayer& L = layers[l];
    QfFold& F = folds[l];
    if (l > 0) {   // self-attention Q/K/V read the previous layer's FFN LayerNorms (output_query / output)
      const QfLayer& P = layers[l - 1];
      SPRC_TRY(fold_one(&F.qkv_q, L.qkv_w, L.qkv_b, P.qo_g, P.qo_beta, 2304, st));
      SPRC_TRY(fold_one(&F.qkv_t, L.qkv_w, L.qkv_b, P.to_g, P.to_beta, 2304, st));
    }
    if (L.has_cross) {
      SPRC_TRY(fold_one(&F.cq, L.cq_w, L.cq_b, L.so_g, L.so_beta, 768, st));
      SPRC_TRY(fold_one(&F.qi, L.qi_w, L.qi_b, L.co_g, L.co_beta, 3072, st));
    } else {
      SPRC_TRY(fold_one(&F.qi, L.qi_w, L.qi_b, L.so_g, L.so_beta, 3072, st));
    }
    SPRC_TRY(fold_one(&F.ti, L.ti_w, L.ti_b, L.so_g, L.so_beta, 3072, st));
  }
  fold_ready = true;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// ViT (pre-LN blocks, eva_vit.py:173-176 / clip_vit.py:132-139):  x += proj(attn(LN1(x)));  x += fc2(act(fc1(LN2(x))))
// The residual stream is never normalised, so the producers (proj, fc2) add the RAW residual and the consumers (qkv of
// the next block, fc1) fold norm1 / norm2.  Block 0's norm1 runs as a kernel (its input has no statistics yet) and so
// does ln_vision (its output is the product).  Statistics: Dv / 64 partials per token (16 ViT-L, 22 ViT-g).
// ------------------------------------------------------------------------------------------------
int Model::prepare_vit_fold(cudaStream_t st) {
  if (vit_fold_ready) return 0;
  if (!vit_st[0]) {
    SPRC_TRY(alloc_t(&vit_st[0], (size_t)vit_cap * 257 * (Dv / 64)));
    SPRC_TRY(alloc_t(&vit_st[1], (size_t)vit_cap * 257 * (Dv / 64)));
  }
  vit_folds.resize(depth);
  for (int i = 0; i < depth; ++i) {
    const VitBlock& b = blocks[i];
    if (i > 0) SPRC_TRY(fold_one(&vit_folds[i].qkv, b.qkv_w, b.qkv_b, b.ln1_g, b.ln1_b, 3 * Dv, st, Dv));
    SPRC_TRY(fold_one(&vit_folds[i].fc1, b.fc1_w, b.fc1_b, b.ln2_g, b.ln2_b, mlp, st, Dv));
  }
  vit_fold_ready = true;
  return 0;
}

bool Model::vit_fold_usable() const { return ln_fold_enabled() && Dv % 64 == 0; }

int Model::vit_blocks_fold(int B, cudaStream_t st) {
  SPRC_TRY(prepare_vit_fold(st));
  const int T = B * 257;
  const float scale = 1.0f / sqrtf((float)dh);
  int cur = 0;
  auto producer = [&](const bf16* A, int K, const bf16* W, const float* bias) -> int {
    GemmFold f;
    f.resid = x;     // raw residual stream (st_res stays null: nothing to normalise)
    f.out16 = xn;    // raw 16-bit copy = A operand of the next consumer
    f.st_out = vit_st[cur ^ 1];
    f.st_stride = vit_cap * 257;
    f.eps = vit_eps;
    GemmDesc d;
    d.A = A;
    d.M = T;
    d.K = d.lda = d.ldw = K;
    d.N = d.ldc = Dv;
    d.W = W, d.bias = bias;
    d.out_f32 = x;
    d.fold = &f;
    SPRC_TRY(gemm_bf16_tcgen05(d, st));
    cur ^= 1;
    return 0;
  };
  auto consumer = [&](const FoldedLinear& w, int N, int act, bf16* out) -> int {
    GemmFold f;
    f.st_in = vit_st[cur];
    f.st_stride = vit_cap * 257;
    f.c = w.c;
    f.eps = vit_eps;
    GemmDesc d;
    d.A = xn;
    d.M = T;
    d.K = d.lda = d.ldw = Dv;
    d.N = d.ldc = N;
    d.W = w.w, d.bias = w.d;
    d.act = act;
    d.out_bf16 = out;
    d.fold = &f;
    return gemm_bf16_tcgen05(d, st);
  };
  for (int i = 0; i < depth; ++i) {
    const VitBlock& b = blocks[i];
    if (i == 0) {
      SPRC_TRY(layernorm(x, T, Dv, b.ln1_g, b.ln1_b, vit_eps, 0, 0, nullptr, xn, st));
      SPRC_TRY(plain_linear(xn, T, Dv, b.qkv_w, 3 * Dv, b.qkv_b, ACT_NONE, qkv, st));
    } else {
      SPRC_TRY(consumer(vit_folds[i].qkv, 3 * Dv, ACT_NONE, qkv));
    }
    AttnDesc a;
    a.Q = qkv;
    a.K = qkv + Dv;
    a.V = qkv + 2 * Dv;
    a.O = att;
    a.B = B;
    a.H = heads;
    a.dh = dh;
    a.Lq = a.Lk = 257;
    a.ldq = a.ldk = a.ldv = 3 * Dv;
    a.ldo = Dv;
    a.q_batch_rows = a.kv_batch_rows = 257;
    a.scale = scale;
    SPRC_TRY(attention(a, st));
    SPRC_TRY(producer(att, Dv, b.proj_w, b.proj_b));
    SPRC_TRY(consumer(vit_folds[i].fc1, mlp, vit_act, h1));
    SPRC_TRY(producer(h1, mlp, b.fc2_w, b.fc2_b));
  }
  return 0;
}

// T8 == 0: the gallery pass (32 query rows per image, dense rows, no text rows; Model::qformer_layers with S = 32)
bool Model::fold_usable(int B, int T8) const { return ln_fold_enabled() && qf_layers >= 2 && T8 >= 0 && B > 0; }

// Layers 0 .. L-2 of one ragged Q-Former pass in the folded schedule, then the materialising LayerNorms; the caller
// (qformer_layers_ragged) runs the last layer in the default schedule.  Row ranges: query rows [0, 32 B) and text rows
// [32 B, 32 B + T8) owe DIFFERENT LayerNorms after the fusion pass's FFNs and may sit in different statistics buffers
// (the cross-attention sublayer touches the query rows only), hence the per-range state (cur*, g*, b*).
int Model::qformer_layers_ragged_fold(int B, int T8, bool with_enc, int Lk, const int32_t* kv_idx0,
                                      const int32_t* kv_idx1, cudaStream_t st) {
  SPRC_TRY(prepare_fold(st));
  const int qrows = 32 * B, rows_all = qrows + T8;
  SPRC_REQUIRE(rows_all <= qf_rows, "qformer: %d rows exceed workspace (%d)", rows_all, qf_rows);
  const size_t to = (size_t)qrows;
  bool raw = false;        // the stream (qh fp32, qhb 16-bit) holds pre-LN sums that still owe a LayerNorm
  int curQ = 0, curT = 0;  // statistics buffer of the query rows / text rows
  const float *gQ = nullptr, *bQ = nullptr, *gT = nullptr, *bT = nullptr;

  // Two weight sets in one launch need the row split on a pair-tile boundary (GemmDesc::m_split % 256 == 0); other
  // batch sizes run the query-row and text-row halves of those GEMMs as two launches, so the arithmetic of a row never
  // depends on the batch it sits in (tests/test_parity_gpu.py::test_composed_query_is_batch_invariant).
  const bool one_launch = qrows % 256 == 0;

  // Producer over rows [row0, row0 + M): s' = A W^T + b + LN(s) in place (qh, qhb), statistics into the other buffer
  // of each row range.  A is given for row0.  The caller flips cur* once the whole sublayer has been issued.
  auto producer = [&](int row0, int M, const bf16* A, int K, const bf16* W, const float* b, const bf16* W2,
                      const float* b2) -> int {
    const bool text_only = row0 >= qrows;
    const int c0 = text_only ? curT : curQ;
    GemmFold f;
    f.split = (!text_only && row0 + M > qrows) ? qrows - row0 : 0;
    f.st_stride = qf_rows;   // statistics planes are [part][qf_rows]; a launch over rows [row0, ..) starts at + row0
    f.resid = qh + (size_t)row0 * 768;
    f.out16 = qhb + (size_t)row0 * 768;
    if (raw) {
      f.st_res = fold_st[c0] + row0;
      f.res_g = text_only ? gT : gQ, f.res_b = text_only ? bT : bQ;
      f.st_res2 = fold_st[curT] + row0, f.res_g2 = gT, f.res_b2 = bT;
    }
    f.st_out = fold_st[c0 ^ 1] + row0;
    f.st_out2 = fold_st[curT ^ 1] + row0;
    GemmDesc d;
    d.A = A;
    d.M = M;
    d.K = d.lda = d.ldw = K;
    d.N = d.ldc = 768;
    d.W = W, d.bias = b;
    if (W2) d.W2 = W2, d.bias2 = b2, d.m_split = qrows;
    d.out_f32 = qh + (size_t)row0 * 768;
    d.fold = &f;
    return gemm_bf16_tcgen05(d, st);
  };
  // Consumer over rows [row0, row0 + M) of the raw stream: out rows [row0, ...) = act(LN(s) W^T + b), N wide
  auto consumer = [&](int row0, int M, const FoldedLinear& w, const FoldedLinear* w2, int N, int act,
                      bf16* out) -> int {
    const bool text_only = row0 >= qrows;
    GemmFold f;
    f.split = (!text_only && row0 + M > qrows) ? qrows - row0 : 0;
    f.st_stride = qf_rows;
    f.st_in = fold_st[text_only ? curT : curQ] + row0;
    f.st_in2 = fold_st[curT] + row0;
    f.c = w.c;
    GemmDesc d;
    d.A = qhb + (size_t)row0 * 768;
    d.M = M;
    d.K = d.lda = d.ldw = 768;
    d.N = d.ldc = N;
    d.W = w.w, d.bias = w.d;
    if (w2) d.W2 = w2->w, d.bias2 = w2->d, d.m_split = qrows, f.c2 = w2->c;
    d.act = act;
    d.out_bf16 = out + (size_t)row0 * N;
    d.fold = &f;
    return gemm_bf16_tcgen05(d, st);
  };
  // GEMMs whose query rows and text rows use different weights: one launch (GemmDesc::W2) or one per row range
  auto consumer2 = [&](const FoldedLinear& wq, const FoldedLinear& wt, int N, int act, bf16* out) -> int {
    if (one_launch) return consumer(0, rows_all, wq, &wt, N, act, out);
    SPRC_TRY(consumer(0, qrows, wq, nullptr, N, act, out));
    return consumer(qrows, T8, wt, nullptr, N, act, out);
  };
  auto producer2 = [&](const bf16* A, int K, const bf16* Wq, const float* bq, const bf16* Wt,
                       const float* bt) -> int {
    if (one_launch) return producer(0, rows_all, A, K, Wq, bq, Wt, bt);
    SPRC_TRY(producer(0, qrows, A, K, Wq, bq, nullptr, nullptr));
    return producer(qrows, T8, A + (size_t)qrows * K, K, Wt, bt, nullptr, nullptr);
  };

  for (int l = 0; l < qf_layers - 1; ++l) {
    const QfLayer& L = layers[l];
    const QfFold& F = folds[l];
    // ---- self-attention over all rows (Qformer.py:175-281) ----
    if (!raw)
      SPRC_TRY(plain_linear(qhb, rows_all, 768, L.qkv_w, 2304, L.qkv_b, ACT_NONE, qqkv, st));
    else if (with_enc && T8 > 0)
      SPRC_TRY(consumer2(F.qkv_q, F.qkv_t, 2304, ACT_NONE, qqkv));
    else
      SPRC_TRY(consumer(0, rows_all, with_enc ? F.qkv_q : F.qkv_t, nullptr, 2304, ACT_NONE, qqkv));
    if (T8 > 0) {
      SPRC_TRY(attention_qf_ragged(qqkv, 2304, qctx, 768, B, rows_all, static_cast<const int4*>(m_pairs), 0.125f, st));
    } else {   // gallery pass: 32 query rows per image, no mask (Model::qformer_layers, S = 32)
      AttnDesc a;
      a.Q = qqkv;
      a.K = qqkv + 768;
      a.V = qqkv + 1536;
      a.O = qctx;
      a.B = B;
      a.H = 12;
      a.dh = 64;
      a.Lq = a.Lk = 32;
      a.ldq = a.ldk = a.ldv = 2304;
      a.ldo = 768;
      a.q_batch_rows = a.kv_batch_rows = 32;
      a.scale = 0.125f;
      SPRC_TRY(attention(a, st));
    }
    SPRC_TRY(producer(0, rows_all, qctx, 768, L.so_w, L.so_b, nullptr, nullptr));
    curQ ^= 1, curT ^= 1, raw = true;
    gQ = gT = L.so_g, bQ = bT = L.so_beta;
    if (with_enc) {
      if (L.has_cross) {   // query rows only (Qformer.py:436-452)
        const int ci = l / 2;
        const long long kv_rows = (long long)B * 257;
        SPRC_TRY(consumer(0, qrows, F.cq, nullptr, 768, ACT_NONE, qcq));
        AttnDesc c;
        c.Q = qcq;
        if (kv_idx0) {   // rerank: plain K/V rows, keys = cat(reference image, candidate image)
          c.K = kv + (size_t)ci * 1536;
          c.V = kv + (size_t)ci * 1536 + 768;
          c.ldk = c.ldv = n_cross * 1536;
          c.kv_idx0 = kv_idx0;
          c.kv_idx1 = kv_idx1;
        } else {         // head-major blocks (cross_kv)
          c.K = kv + (size_t)ci * 24 * kv_rows * 64;
          c.V = kv + ((size_t)ci * 24 + 12) * kv_rows * 64;
          c.kv_head_stride = kv_rows * 64;
          c.ldk = c.ldv = 64;
        }
        c.O = qctx;
        c.B = B;
        c.H = 12;
        c.dh = 64;
        c.Lq = 32;
        c.Lk = Lk;
        c.ldq = 768;
        c.ldo = 768;
        c.q_batch_rows = 32;
        c.kv_batch_rows = 257;
        c.scale = 0.125f;
        c.Lk1 = 257;
        SPRC_TRY(attention(c, st));
        SPRC_TRY(producer(0, qrows, qctx, 768, L.co_w, L.co_b, nullptr, nullptr));
        curQ ^= 1;
        gQ = L.co_g, bQ = L.co_beta;
      }
      // query rows -> *_query FFN, text rows -> text FFN (Qformer.py:455-468)
      if (T8 > 0) {
        SPRC_TRY(consumer2(F.qi, F.ti, 3072, ACT_GELU, qffn));
        SPRC_TRY(producer2(qffn, 3072, L.qo_w, L.qo_b, L.to_w, L.to_b));
      } else {
        SPRC_TRY(consumer(0, qrows, F.qi, nullptr, 3072, ACT_GELU, qffn));
        SPRC_TRY(producer(0, qrows, qffn, 3072, L.qo_w, L.qo_b, nullptr, nullptr));
      }
      curQ ^= 1, curT ^= 1;
      gQ = L.qo_g, bQ = L.qo_beta, gT = L.to_g, bT = L.to_beta;
    } else {   // no encoder states: every row takes the text FFN (Qformer.py:434-435, 469-475)
      SPRC_TRY(consumer(0, rows_all, F.ti, nullptr, 3072, ACT_GELU, qffn));
      SPRC_TRY(producer(0, rows_all, qffn, 3072, L.to_w, L.to_b, nullptr, nullptr));
      curQ ^= 1, curT ^= 1;
      gQ = gT = L.to_g, bQ = bT = L.to_beta;
    }
  }
  // materialise LN(s) for the last layer (fp32 stream + 16-bit operand copy)
  SPRC_TRY(layernorm(qh, qrows, 768, gQ, bQ, 1e-12f, 0, 0, qh, qhb, st));
  if (T8 > 0) SPRC_TRY(layernorm(qh + to * 768, T8, 768, gT, bT, 1e-12f, 0, 0, qh + to * 768, qhb + to * 768, st));
  return 0;
}

}  // namespace sprc
