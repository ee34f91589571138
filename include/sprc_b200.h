/* libsprc_b200 — C ABI of the B200-native composed-image-retrieval inference path.
 *
 * The reference (chunmeifeng/SPRC) has no FFI layer: its boundary is the Python module surface
 * `lavis.models.load_model_and_preprocess` -> `Blip2QformerCirAlignPrompt.{extract_target_features,
 * inference, inference_rerank}` plus `validate_blip.compute_*`.  This header is the
 * C boundary underneath our Python mirror of that surface (sprc_b200/model.py binds it with ctypes);
 * every entry point names the reference function it replaces (paths relative to /root/reference/src).
 *
 * Conventions
 *   - return 0 on success, a negative errno-style code on failure; `sprc_last_error()` returns a
 *     thread-local message for the last failure on the calling thread;
 *   - all `const void*` / `void*` tensor arguments are DEVICE pointers unless the name ends in `_host`;
 *     the caller owns every input and output buffer, the handle owns weights and workspace;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden synchronisation
 *     except in the `*_host` convenience calls, which synchronise the stream before returning;
 *   - one handle per GPU / process, not re-entrant.
 */
#ifndef SPRC_B200_H_
#define SPRC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPRC_ABI_VERSION 1

#define SPRC_NUM_QUERY_TOKENS 32 /* blip2_qformer_cir_align_prompt.py:52  num_query_token */
#define SPRC_MAX_TXT_LEN 32      /* blip2_qformer_cir_align_prompt.py:55  max_txt_len */
#define SPRC_EMBED_DIM 256       /* blip2_qformer_cir_align_prompt.py:54  embed_dim */
#define SPRC_VIT_TOKENS 257      /* (224/14)^2 + cls; eva_vit.py:324-331, clip_vit.py:171-178 */
#define SPRC_QF_HIDDEN 768       /* BertConfig bert-base-uncased; blip2.py:47-61 */

typedef struct sprc_handle sprc_handle;

enum sprc_vit_kind {
  SPRC_VIT_EVA_G = 0, /* eva_vit.py:428-441   create_eva_vit_g : 1408 x 39 blocks, 16 heads x 88, mlp 6144 */
  SPRC_VIT_CLIP_L = 1 /* clip_vit.py:242-250  create_clip_vit_L: 1024 x 23 blocks, 16 heads x 64, mlp 4096 */
};

enum sprc_dtype { SPRC_F32 = 0, SPRC_F16 = 1, SPRC_BF16 = 2, SPRC_I64 = 3, SPRC_I32 = 4 };

typedef struct sprc_config {
  int vit_kind;    /* enum sprc_vit_kind */
  int vit_depth;   /* 0 = the reference's depth for vit_kind; smaller values build truncated test models */
  int qf_layers;   /* 0 = 12 (BERT-base); smaller values for tests */
  int max_images;  /* largest B accepted by sprc_encode_gallery (workspace is sized for it) */
  int max_queries; /* largest Bq accepted by sprc_encode_query */
  int max_pairs;   /* largest R*T accepted by sprc_rerank (0 = rerank workspace not allocated) */
  int device;      /* CUDA device ordinal */
  int act_dtype;   /* 16-bit operand/activation format of every kernel: 0 = bf16 (default), 1 = fp16 (the
                      reference's own autocast precision, blip2.py:36-44).  Process-wide; tensors named
                      *_bf16 in this header hold whichever format is active. */
} sprc_config;

/* One tensor of the reference checkpoint (`ckpt["Blip2QformerCirAlignPrompt"]`, utils.py:208-222),
 * under its reference state-dict key, fp32 or fp16, host or device memory. */
typedef struct sprc_tensor_desc {
  const char* name;
  int dtype; /* enum sprc_dtype */
  int ndim;
  int64_t shape[4];
  const void* data;
} sprc_tensor_desc;

int sprc_abi_version(void);
const char* sprc_last_error(void);

/* Blip2QformerCirAlignPrompt.__init__ / from_config (blip2_qformer_cir_align_prompt.py:44-92,502-529). */
int sprc_create(const sprc_config* cfg, sprc_handle** out);
void sprc_destroy(sprc_handle* h);

/* model.load_state_dict(ckpt[...], strict=False) (blip_validate.py:107-109).  Unknown keys are ignored
 * (the LM head, temp, prompt_tokens); `*n_missing` receives how many required tensors are still unset. */
int sprc_load_weights(sprc_handle* h, const sprc_tensor_desc* tensors, int n, int* n_missing);
/* Name of the i-th still-missing required tensor, or NULL. */
const char* sprc_missing_weight(sprc_handle* h, int i);

/* extract_target_features (blip2_qformer_cir_align_prompt.py:364-386): images fp32 [B,3,224,224] ->
 * feats [B,32,256] (unit-norm rows) and raws = ln_vision(ViT(images)) [B,257,Dv].  Any output may be NULL. */
int sprc_encode_gallery(sprc_handle* h, const float* images, int B, float* feats_f32, void* feats_bf16,
                        float* raws_f32, void* raws_bf16, void* stream);

/* The fusion half of `inference` (blip2_qformer_cir_align_prompt.py:312-350): reference embeds
 * [Bq,257,Dv] (fp32 or bf16, `ref_dtype`) + token ids / attention mask [Bq,32] int64 (tokenisation stays
 * on the host, :323-329) -> fusion_feats [Bq,256] fp32 unit-norm.  `ref_rows` (optional, int32 [Bq])
 * gathers the reference rows out of a resident raw-embed table instead of a packed [Bq,...] tensor. */
int sprc_encode_query(sprc_handle* h, const void* ref_raws, int ref_dtype, const int32_t* ref_rows,
                      const int64_t* input_ids, const int64_t* attention_mask, int Bq, float* fusion_f32,
                      void* fusion_bf16, void* stream);

/* Same result as sprc_encode_query, computed over the live text rows only: the reference pads every caption to 32
 * tokens and runs the padded rows through both Q-Former passes although nothing ever reads them (padded keys have
 * softmax weight exp(-10000) = 0, align_prompt.py:343,348 keep the query rows / row 32).  text_len_host[b] (HOST
 * memory, int32 [Bq]) = number of live tokens of caption b = sum of its attention-mask row, which must be a prefix
 * mask (tokens, then padding: what BertTokenizer(padding="max_length") produces).  The tokenizer runs on the host, so
 * the lengths are known there without a device round trip. */
int sprc_encode_query_lens(sprc_handle* h, const void* ref_raws, int ref_dtype, const int32_t* ref_rows,
                           const int64_t* input_ids, const int32_t* text_len_host, int Bq, float* fusion_f32,
                           void* fusion_bf16, void* stream);

/* The similarity half of `inference` (:353-358) fused with the ranking of validate_blip.py:44-46,253-255:
 * sim[q,n] = max_t <query[q], gallery[n,t]>, top-k by (sim desc, row asc).  gallery is bf16 [N,32,256],
 * queries bf16 [Q,256].  out_full (optional) receives the whole fp32 [Q,N] matrix (what `inference`
 * returns); out_score/out_idx (optional, [Q,k]) the ranking, idx = row_offset + local row. */
int sprc_sim_topk(sprc_handle* h, const void* queries_bf16, int Q, const void* gallery_bf16, int64_t N,
                  int64_t row_offset, int k, float* out_score, int32_t* out_idx, float* out_full, void* stream);

/* The same scan with a GROUPED output, for the multi-GPU step: Q = world * group_rows queries (all
 * ranks' batches after the all-gather of query vectors) against this rank's shard in ONE launch; the [group_rows, k]
 * block of query group g is written at out_score / out_idx + g * group_stride (elements), i.e. straight into the
 * exchange buffer [dest rank][scores | rows][Bq][k] with group_stride = 2 * Bq * k and out_idx = out_score + Bq * k. */
int sprc_sim_topk_grouped(sprc_handle* h, const void* queries_bf16, int Q, const void* gallery_bf16, int64_t N,
                          int64_t row_offset, int k, float* out_score, int32_t* out_idx, int group_rows,
                          int64_t group_stride, void* stream);

/* Merge P candidate lists per query (the per-shard top-k after the NCCL all-gather):
 * cand_* are [P,Q,k]; output [Q,k] sorted by (score desc, idx asc). */
int sprc_topk_merge(sprc_handle* h, const float* cand_score, const int32_t* cand_idx, int P, int Q, int k,
                    float* out_score, int32_t* out_idx, void* stream);
/* The same merge straight off the exchange buffer of the multi-GPU step: cand is int32
 * [P][2][Q][k] - for every source rank p the fp32 scores (bit pattern) of this rank's Q queries, then their global
 * rows - exactly what ONE all-to-all of per-shard candidates delivers, so each rank merges only its own queries. */
int sprc_topk_merge_packed(sprc_handle* h, const int32_t* cand, int P, int Q, int k, float* out_score,
                           int32_t* out_idx, void* stream);

/* sim of selected (query, row) pairs — CIRR subset members (validate_blip.py:268-271): rows int32 [Q,m]
 * (negative = skip, score -inf) -> out [Q,m]. */
int sprc_gather_scores(sprc_handle* h, const void* queries_bf16, int Q, const void* gallery_bf16, int64_t N,
                       const int32_t* rows, int m, float* out, void* stream);

/* inference_rerank (blip2_qformer_cir_rerank.py:399-445): for each of R queries, T candidates;
 * ref rows int32 [R], candidate rows int32 [R*T] index a resident bf16 raw-embed table [*,257,Dv];
 * p[R*T] = softmax(mean_q itm_head(h))[:, 1]. */
int sprc_rerank(sprc_handle* h, const void* raws_bf16, const int32_t* ref_rows, const int32_t* cand_rows,
                const int64_t* input_ids, const int64_t* attention_mask, int R, int T, float* p, void* stream);

/* sprc_rerank over the live text rows only (see sprc_encode_query_lens): text_len_host int32 [R] in HOST memory. */
int sprc_rerank_lens(sprc_handle* h, const void* raws_bf16, const int32_t* ref_rows, const int32_t* cand_rows,
                     const int64_t* input_ids, const int32_t* text_len_host, int R, int T, float* p, void* stream);

/* End-to-end query step with HOST buffers (what generate_*_val_predictions + compute_* do per batch,
 * validate_blip.py:386-408,253-255): H2D of ids/mask/ref_rows, fusion, scan, top-k, D2H of [Bq,k]; returns when the
 * results are in out_*_host (= sprc_query_topk_host_submit + sprc_query_topk_host_wait). */
int sprc_query_topk_host(sprc_handle* h, const void* raws_bf16, const void* gallery_bf16, int64_t N,
                         const int32_t* ref_rows_host, const int64_t* input_ids_host,
                         const int64_t* attention_mask_host, int Bq, int k, float* out_score_host,
                         int32_t* out_idx_host, void* stream);
/* The same step split in two so that a query loop can keep the GPU busy while the host prepares the next batch
 * (the reference's loop, validate_blip.py:386-408, is strictly serial): _submit enqueues the H2D copies, the kernels
 * and the D2H copies of ONE batch on `stream` and returns; _wait blocks until the OLDEST submitted, not yet awaited
 * batch has its results in the out_*_host buffers it was submitted with.  Up to 4 batches may be in flight; the host
 * buffers of a batch (pinned memory) must stay untouched until its _wait returns.  Same stream for every call. */
int sprc_query_topk_host_submit(sprc_handle* h, const void* raws_bf16, const void* gallery_bf16, int64_t N,
                                const int32_t* ref_rows_host, const int64_t* input_ids_host,
                                const int64_t* attention_mask_host, int Bq, int k, float* out_score_host,
                                int32_t* out_idx_host, void* stream);
int sprc_query_topk_host_wait(sprc_handle* h);
/* The same step from caption STRINGS, as `inference` receives them (align_prompt.py:312-329: text is list[str],
 * tokenised inside the call): the batch is tokenised by `tok` (sprc_tokenize_host, `threads` workers) straight into
 * pinned staging owned by the handle, then submitted like sprc_query_topk_host_submit; pair with
 * sprc_query_topk_host_wait.  While the GPU works on batch i the host tokenises batch i+1.  texts / offsets as in
 * sprc_tokenize_host; ref_rows_host is copied before returning.  Returns -84 if a caption needs the string-level
 * host path (nothing is submitted then). */
typedef struct sprc_tokenizer sprc_tokenizer;
int sprc_query_topk_strings_submit(sprc_handle* h, const sprc_tokenizer* tok, const void* raws_bf16,
                                   const void* gallery_bf16, int64_t N, const int32_t* ref_rows_host, const char* texts,
                                   const int64_t* offsets, int Bq, int k, int threads, float* out_score_host,
                                   int32_t* out_idx_host, void* stream);

/* Number of kernels this library has launched on behalf of the calling process (bench `gpu_launches`). */
int64_t sprc_launch_count(void);

/* Sets the process-wide 16-bit format for the single-op entry points (sprc_create sets it from the config). */
int sprc_set_act_dtype(int fp16);

/* Optional per-launch timing with CUDA events on the launching stream, by kernel category
 * (0 gemm, 1 attention, 2 layernorm, 3 scan, 4 merge).  sprc_profile(1) clears and enables,
 * sprc_profile_read fills out[cat*4 + {0,1,2,3}] = {ms, algorithmic flops, algorithmic bytes, launches}. */
#define SPRC_PROF_NCAT 5
int sprc_profile(int enable);
int sprc_profile_read(double* out, int ncat);
/* CSV of per-(category, shape tag) aggregates of the recorded launches: cat,tag,launches,total_ms,flops,bytes */
int sprc_profile_dump(const char* path);

/* Gallery-side image preprocessing on the GPU (replaces data_utils.py:52-72,91-105 `targetpad_transform` = TargetPad,
 * Resize(dim, BICUBIC), CenterCrop(dim), ToTensor, Normalize on PIL images; bit-exact with Pillow's 8-bit resampler).
 * pixels: packed decoded RGB uint8 images (device); desc: [n][16] int64 per-image descriptors and tables: int32
 * fixed-point coefficient tables (device), both produced by sprc_b200/preprocess.py; tmp: uint8 workspace of
 * sum(rows_i) * dim * 3 bytes; mean3/std3: host pointers; out: [n,3,dim,dim] fp32 (device). */
int sprc_preprocess_targetpad(const uint8_t* pixels, const int64_t* desc, const int32_t* tables, int n, int dim,
                              int max_rows, uint8_t* tmp, const float* mean3, const float* std3, float* out,
                              void* stream);

/* ---- single-op entry points (tests and micro-benchmarks) ------------------------------------- */
/* C = act(A[M,K] W[N,K]^T + bias) (+ residual); impl 0 = tcgen05 product kernel, 1 = CUDA-core checker. */
int sprc_op_gemm(const void* A_bf16, const void* W_bf16, int M, int N, int K, int lda, int ldw, int grp_rows,
                 int grp_stride, const float* bias, const float* residual, float* out_f32, void* out_bf16,
                 int ldc, int act, int impl, void* stream);
/* out_f32 = LayerNorm(A W^T + bias + residual) * gamma + beta over rows of N = 768, out_ln16 = the same in the 16-bit
 * operand format (fused Q-Former post-LN sublayer, Qformer.py:291-295,373-381); residual may alias out_f32. */
/* Same GEMM with TWO weight sets in one launch: rows [0, m_split) use (W, bias), rows [m_split, M) use (W2, bias2)
 * (the fusion pass's query rows -> *_query FFN, text rows -> text FFN, Qformer.py:455-468).  Dense rows,
 * m_split % 256 == 0. */
int sprc_op_gemm2w(const void* A_bf16, const void* W_bf16, const void* W2_bf16, int M, int m_split, int N, int K,
                   const float* bias, const float* bias2, const float* residual, float* out_f32, void* out_bf16,
                   int act, void* stream);
/* Q-Former self-attention over the ragged row layout (csrc/attention_qfr.cu): qkv [rows_total, 3*768] packed
 * Q|K|V, rows [0,32B) query rows, then per-sample text slots; pairs_dev int32 [ceil(B/2)][4] = {toff0, L0, toff1, L1}. */
int sprc_op_attention_ragged(const void* qkv, int ldqkv, void* out, int ldo, int B, int rows_total,
                             const int32_t* pairs_dev, float scale, void* stream);
int sprc_op_layernorm(const float* x, int rows, int width, const float* gamma, const float* beta, float eps,
                      int grp_rows, int grp_stride, float* out_f32, void* out_bf16, void* stream);
int sprc_op_attention(const void* Q, const void* K, const void* V, void* O, int B, int H, int dh, int Lq,
                      int Lk, int ldq, int ldk, int ldv, int ldo, int q_batch_rows, int kv_batch_rows,
                      const float* key_mask, float scale, void* stream);

/* Cross-attention of inference_rerank (blip2_qformer_cir_rerank.py:419-436): sample b's 32 query rows attend over
 * cat(image kv_idx0[b], image kv_idx1[b]) = 514 keys of a K/V table with kv_rows_total rows (257 per image); heads are
 * 64-column slices of ldk/ldv-pitch rows (kv_head_stride = 0) or contiguous [kv_rows_total, 64] blocks kv_head_stride
 * elements apart.  dh = 64. */
int sprc_op_attention_pairs(const void* Q, const void* K, const void* V, void* O, int B, int H, int ldq, int ldk, int ldv,
                            int ldo, int q_batch_rows, const int32_t* kv_idx0, const int32_t* kv_idx1,
                            int64_t kv_rows_total, int64_t kv_head_stride, float scale, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Host-side caption tokenizer (csrc/tokenizer.cpp; no CUDA).  Replaces the per-batch Python tokenizer call inside
 * `inference` (blip2_qformer_cir_align_prompt.py:323-329: `self.tokenizer(text, padding="max_length",
 * truncation=True, max_length=32)` with the tokenizer of blip2.py:30-34 = transformers 4.36 BertTokenizer + [DEC]):
 * BasicTokenizer + greedy WordPiece, [CLS] ... [SEP], truncated / zero-padded to max_len.
 * sprc_tokenizer_create: vocab_utf8 = contents of a BERT vocab.txt (one token per line, id = line number), or
 *   NULL / 0 for the synthetic hashed vocabulary used by string-driven synthetic runs (id = 1000 + FNV-1a % 29000).
 * sprc_tokenize_host: texts = n UTF-8 captions back to back, caption i = bytes [offsets[i], offsets[i+1]);
 *   ids / mask int64 [n, max_len] (HOST), lens int32 [n] live tokens (or NULL), complex_flags uint8 [n] (or NULL):
 *   1 = the caption holds a character whose normalisation depends on its neighbours (combining marks, final sigma;
 *   tools/gen_unicode_tables.py) - its row is zeroed, lens = -1, and the caller tokenises it with the exact
 *   string-level path (sprc_b200/tokenizer.py).  threads <= 0: one per core, at most 16. */
int sprc_tokenizer_create(const char* vocab_utf8, int64_t vocab_bytes, sprc_tokenizer** out);
void sprc_tokenizer_destroy(sprc_tokenizer* t);
int sprc_tokenize_host(const sprc_tokenizer* t, const char* texts, const int64_t* offsets, int n, int max_len,
                       int threads, int64_t* ids, int64_t* mask, int32_t* lens, uint8_t* complex_flags);

/* ---- indexing feed: PNG files -> packed RGB8 (host; csrc/png.cpp) --------------------------------------------
 * Replaces the decode half of the reference's index DataLoader (src/utils.py:54-64 `DataLoader(num_workers=2)` whose
 * workers run `PIL.Image.open(path)` + `convert("RGB")`, src/data_utils.py:91-105,167-186,253-270): `threads` C++
 * workers decode a batch of files straight into `out` (a pinned arena in practice; the GPU resize kernels of
 * sprc_preprocess_targetpad read it after ONE copy).  Pixel values are Pillow's `Image.open(p).convert("RGB")`,
 * bit for bit (tests/test_png.py), for non-interlaced PNGs with 8-bit gray / gray+alpha / RGB / RGBA / palette samples
 * and 1/2/4-bit gray / palette.
 * paths = n UTF-8 file names back to back, file i = bytes [path_offsets[i], path_offsets[i+1]);
 * pixel_offsets int64 [n+1]: image i occupies out[pixel_offsets[i] .. +3*w*h); wh int32 [n][3] = (width, height,
 * the mode Pillow opens the file in: 0 "RGB", 1 "L", 2 "1", 3 "P", 4 "LA", 5 "RGBA" - resize-then-convert, the
 * reference's transform order, equals convert-then-resize only for 0 and 1);
 * status int32 [n]: 0 decoded, 1 not taken by this decoder (16-bit samples, Adam7, not a PNG: the caller decodes the
 * file with Pillow), 2 corrupt, 3 unreadable (Pillow decides; the reference's datasets drop such images).
 * Returns -34 without decoding when the batch needs more than out_capacity bytes (pixel_offsets[n] = bytes needed).
 * threads <= 0: one per core, at most 32. */
int sprc_png_decode_files(const char* paths, const int64_t* path_offsets, int n, int threads, uint8_t* out,
                          int64_t out_capacity, int64_t* pixel_offsets, int32_t* wh, int32_t* status);
/* The decoder's own DEFLATE / zlib inflater (csrc/inflate.h) alone, for tests and micro-benchmarks: 0 when `in` is a
 * zlib stream that inflates to exactly out_bytes bytes with a matching Adler-32, a positive code otherwise (inside
 * sprc_png_decode_files such streams are retried with zlib itself). */
int sprc_op_inflate_zlib(const uint8_t* in, int64_t in_bytes, uint8_t* out, int64_t out_bytes);

#ifdef __cplusplus
}
#endif
#endif /* SPRC_B200_H_ */
