// Host launchers of the non-GEMM kernels (elementwise.cu, attention.cu, scan.cu).
#pragma once
#include "common.h"

namespace sprc {

int64_t launch_count();
void count_launch(int n = 1);

// per-launch CUDA-event timing by category (bench.py roofline; off by default)
enum ProfCat { PROF_GEMM = 0, PROF_ATTN = 1, PROF_ELEMWISE = 2, PROF_SCAN = 3, PROF_MERGE = 4, PROF_NCAT = 5 };
bool prof_enabled();
// Row-sweeping kernels (GEMM tiles, LayerNorm rows, attention items) alternate their sweep direction from launch
// to launch, so each kernel starts on the rows its producer touched LAST, which are the ones still in the 126 MB L2
// (activations of a 592-query step are 116-290 MB per tensor).  Results do not depend on the direction.
int next_sweep_reverse();
void prof_begin(cudaStream_t st);
void prof_end(int cat, double flops, double bytes, cudaStream_t st, const char* tag = nullptr);
int prof_dump(const char* path);
void prof_set(bool on);
int prof_read(double* out, int ncat);

// ---- elementwise.cu --------------------------------------------------------------------------
// Row-wise LayerNorm over `width` (fp32 statistics, two-pass like ATen): eva_vit.py:175-176
// (eps 1e-6), clip_vit.py:100-107 (eps 1e-5), blip2.py:193-199 ln_vision (eps 1e-5),
// Qformer.py:65,288,374 (eps 1e-12).  Rows may be grouped like GemmDesc rows.
int layernorm(const float* x, int rows, int width, const float* gamma, const float* beta, float eps, int grp_rows,
              int grp_stride, float* out_f32, bf16* out_bf16, cudaStream_t st);

// images fp32 [B,3,224,224] -> bf16 patch rows [B*256, ldp] in conv-weight order (c, ky, kx)
// (eva_vit.py:196,203 / clip_vit.py:160,173 express the same contraction as a stride-14 conv).
int im2col_patches(const float* images, int B, bf16* patches, int ldp, cudaStream_t st);

// x[b,0,:] = cls + pos[0];  x[b,1+p,:] = patch_out[b*256+p,:] + pos[1+p]   (eva_vit.py:327-331,
// clip_vit.py:176-177)
int vit_assemble_tokens(const float* patch_out, const float* cls, const float* pos, int B, int width, float* x,
                        cudaStream_t st);

// Q-Former embedding rows before LayerNorm (Qformer.py:98-110): rows [0,32) of each sample = query
// embeds (broadcast [32,768] when q_batch_rows == 0, else sample b starts at row b*q_batch_rows),
// rows [32,64) = word[ids] + pos[0..31].  ids == nullptr -> only the 32 query rows (S = 32).
// Sample b reads token ids row b / ids_div (rerank repeats each caption T times, rerank.py:418-419).
int qformer_embed_rows(const float* query_embeds, int q_batch_rows, const int64_t* ids, int ids_div,
                       const float* word_emb, const float* pos_emb, int vocab, int B, float* out, cudaStream_t st);

// Ragged Q-Former layout (attention_qfr.cu): pre-LayerNorm embedding rows and row gathers
int qformer_embed_ragged(const float* query_embeds, int q_is_batched, const int64_t* ids, int ids_div,
                         const int* slot_sample, const int* toff, const int* len, const float* word_emb,
                         const float* pos_emb, int vocab, int B, int rows_total, float* out, cudaStream_t st);
int gather_rows768(const float* src32, const bf16* src16, const int* rows, int base, int n, float* dst32, bf16* dst16,
                   cudaStream_t st);
int attention_qf_ragged(const bf16* qkv, int ldqkv, bf16* out, int ldo, int B, int rows_total, const int4* pairs_dev,
                        float scale, cudaStream_t st);

// additive self-attention mask (Qformer.py:807): out[b, j] = 0 for j < 32, (1 - mask[b, j-32]) * -10000 after
// (sample b reads attention_mask row b / div)
int qformer_key_mask(const int64_t* attention_mask, int div, int B, float* out, cudaStream_t st);

int convert_f32_to_bf16(const float* in, bf16* out, size_t n, cudaStream_t st);
int convert_f16_to_bf16(const void* in, bf16* out, size_t n, cudaStream_t st);
int convert_f16_to_f32(const void* in, float* out, size_t n, cudaStream_t st);
// out[i, :] = table[rows[i], :] for rows of `row_elems` elements (bf16 -> bf16, or fp32 -> bf16)
int gather_rows_bf16(const void* table, int table_dtype, const int32_t* rows, int n_rows, size_t row_elems,
                     bf16* out, cudaStream_t st);

// F.normalize(x, dim=-1) with eps 1e-12 over rows of 256 (blip2_qformer_cir_align_prompt.py:348,385).
// Input rows are `in_row_stride` floats apart (lets the text [CLS] row 32 of each sample be picked).
int l2norm_rows256(const float* in, size_t in_row_stride, int rows, float* out_f32, bf16* out_bf16,
                   cudaStream_t st);

// itm_head + mean over 32 query rows + softmax[:, 1] (blip2_qformer_cir_rerank.py:440-445).
// h is fp32 [pairs, rows_per_pair, 768]; only the first 32 rows of each pair are used.
int itm_head_prob(const float* h, int rows_per_pair, int pairs, const float* w, const float* b, float* p,
                  cudaStream_t st);

// ---- preprocess.cu ---------------------------------------------------------------------------
// TargetPad + bicubic Resize + CenterCrop + ToTensor + Normalize (data_utils.py:52-72, 91-105) on decoded RGB uint8
// images, bit-exact with PIL/torchvision; descriptors and coefficient tables come from sprc_b200/preprocess.py.
int preprocess_targetpad(const uint8_t* pixels, const long long* desc, const int* tables, int n, int dim, int max_rows,
                         uint8_t* tmp, const float* mean, const float* stdv, float* out, cudaStream_t st);

// ---- attention.cu ----------------------------------------------------------------------------
// softmax(Q K^T * scale + key_mask) V for B samples x H heads, bf16 in/out, fp32 softmax:
// ViT MHSA (eva_vit.py:128-145, clip_vit.py:134), Q-Former self- and cross-attention
// (Qformer.py:211-270).  Rows of sample b start at row b*q_batch_rows (Q,O) / b*kv_batch_rows (K,V);
// head h occupies columns [h*dh, (h+1)*dh).  Optional two-segment keys (rerank, KV = cat(ref,target)):
// keys [0,Lk1) come from sample kv_idx0[b], keys [Lk1,Lk) from sample kv_idx1[b].
struct AttnDesc {
  const bf16* Q = nullptr;
  const bf16* K = nullptr;
  const bf16* V = nullptr;
  bf16* O = nullptr;
  int B = 0, H = 0, dh = 0, Lq = 0, Lk = 0;
  int ldq = 0, ldk = 0, ldv = 0, ldo = 0;
  int q_batch_rows = 0, kv_batch_rows = 0;
  const float* key_mask = nullptr;  // additive [B, Lk]
  float scale = 1.f;
  const int32_t* kv_idx0 = nullptr;
  const int32_t* kv_idx1 = nullptr;
  int Lk1 = 0;
  // > 0: K and V are head-major (GemmDesc::out_col_block layout): head h starts kv_head_stride elements after head
  // h - 1 and its rows are ldk = ldv = 64 elements apart.  0: heads are 64-column slices of ldk/ldv-pitch rows.
  long long kv_head_stride = 0;
  // rows of the K/V table the sample index tables kv_idx0 / kv_idx1 point into (bounds of the TMA tensor map)
  long long kv_rows_total = 0;
};
int attention(const AttnDesc& a, cudaStream_t st);

// ---- scan.cu ---------------------------------------------------------------------------------
int sim_topk(const bf16* queries, int Q, const bf16* gallery, int64_t N, int64_t row_offset, int k,
             float* out_score, int32_t* out_idx, float* out_full, void* workspace, size_t workspace_bytes,
             cudaStream_t st, int out_group_rows = 0, size_t out_group_stride = 0);
size_t sim_topk_workspace_bytes(int Q, int64_t N, int k, bool caller_has_full);
int topk_merge(const float* cand_score, const int32_t* cand_idx, int P, int Q, int k, float* out_score,
               int32_t* out_idx, cudaStream_t st, size_t pstride = 0);   // pstride: elements between lists (0 = Q*k)
int gather_scores(const bf16* queries, int Q, const bf16* gallery, int64_t N, const int32_t* rows, int m,
                  float* out, cudaStream_t st);

}  // namespace sprc
