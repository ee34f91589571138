// Handle behind the C ABI: model dimensions, packed weights, workspace, and the forward passes
// (ViT -> ln_vision -> Q-Former -> ITC heads) that the sprc_* entry points enqueue.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "common.h"
#include "ops.h"

namespace sprc {

enum QfLiveOut { QF_OUT_ALL = 0, QF_OUT_QUERY_ROWS = 1, QF_OUT_TEXT_CLS = 2 };

struct WeightSlot {
  void* dst = nullptr;   // device destination (start of the packed tensor)
  int dst_dtype = 0;     // 0 = f32, 2 = bf16  (enum sprc_dtype values)
  int64_t rows = 0;      // expected source shape, flattened to [rows, cols]
  int64_t cols = 0;
  int64_t ld = 0;        // destination pitch in elements
  int64_t row_off = 0;   // destination row offset inside a packed tensor
  bool flexible_rows = false;  // source may have fewer rows (word embeddings)
  bool required = true;
  bool loaded = false;
};

struct VitBlock {
  float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  bf16 *qkv_w, *proj_w, *fc1_w, *fc2_w;
  float *qkv_b, *proj_b, *fc1_b, *fc2_b;
};

struct QfLayer {
  bf16* qkv_w;  // [2304,768] rows = [query; key; value]   (Qformer.py:133-139)
  float* qkv_b;
  bf16* so_w;   // attention.output.dense
  float *so_b, *so_g, *so_beta;
  bool has_cross;
  bf16 *cq_w, *co_w;  // crossattention.self.query / crossattention.output.dense
  float *cq_b, *co_b, *co_g, *co_beta;
  bf16 *ti_w, *to_w;  // intermediate / output (text FFN)
  float *ti_b, *to_b, *to_g, *to_beta;
  bf16 *qi_w, *qo_w;  // intermediate_query / output_query
  float *qi_b, *qo_b, *qo_g, *qo_beta;
};

struct Model {
  // ---- dimensions ----
  int vit_kind = 0, Dv = 0, depth = 0, heads = 16, dh = 0, mlp = 0;
  float vit_eps = 1e-6f;
  int vit_act = ACT_GELU;
  int qf_layers = 12, n_cross = 6;
  int max_images = 0, max_queries = 0, max_pairs = 0;
  int device = 0;
  int act_dtype_fp16 = 0;  // format the weights were packed in (library mode at creation)
  int vocab = 30523;
  static constexpr int KP = 592;  // patch K (588) padded to a 16-byte pitch

  // ---- weights ----
  std::map<std::string, WeightSlot> slots;
  std::vector<void*> allocs;
  std::vector<std::string> missing_cache;
  float *cls = nullptr, *pos = nullptr, *patch_b = nullptr, *ln_pre_g = nullptr, *ln_pre_b = nullptr;
  bf16* patch_w = nullptr;
  std::vector<VitBlock> blocks;
  float *lnv_g = nullptr, *lnv_b = nullptr;
  float *query_tokens = nullptr, *word_emb = nullptr, *pos_emb = nullptr, *emb_g = nullptr, *emb_b = nullptr;
  std::vector<QfLayer> layers;
  bf16* kv_w = nullptr;  // [n_cross*1536, Dv]: per cross layer rows = [key(768); value(768)]
  float* kv_b = nullptr;
  bf16 *vproj_w = nullptr, *tproj_w = nullptr;
  float *vproj_b = nullptr, *tproj_b = nullptr, *itm_w = nullptr, *itm_b = nullptr;
  void* staging = nullptr;
  size_t staging_bytes = 0;

  // ---- workspace ----
  int vit_cap = 0;   // images
  int enc_cap = 0;   // images whose raw embeds / cross K,V fit (max(images, queries, rerank images))
  int qf_rows = 0;   // Q-Former rows
  bf16 *patches = nullptr, *xn = nullptr, *qkv = nullptr, *att = nullptr, *h1 = nullptr;
  float *patch_out = nullptr, *x = nullptr;
  bf16 *raws = nullptr, *kv = nullptr;
  float *qh = nullptr, *qt = nullptr, *qproj = nullptr, *qmask = nullptr;
  bf16 *qhb = nullptr, *qqkv = nullptr, *qctx = nullptr, *qcq = nullptr, *qffn = nullptr;
  // host-call staging (sprc_query_topk_host) and scan workspace
  int64_t* d_ids = nullptr;
  int64_t* d_mask = nullptr;
  int32_t* d_rows = nullptr;
  int32_t* d_rows2 = nullptr;
  int32_t* d_meta = nullptr;     // ragged layout tables (see build_ragged_meta)
  size_t meta_cap = 0;
  std::vector<int32_t> h_meta;
  // pinned staging ring for the row tables: a pageable source would make cudaMemcpyAsync synchronise the stream
  // first (the host could never run ahead of the GPU); a slot is reused only after its own copy has executed
  static constexpr int kMetaRing = 8;
  int32_t* h_meta_pin[kMetaRing] = {};
  cudaEvent_t meta_ev[kMetaRing] = {};
  unsigned meta_seq = 0;
  const int32_t *m_toff = nullptr, *m_len = nullptr, *m_slot = nullptr, *m_cls = nullptr;
  const void* m_pairs = nullptr;
  bf16* d_fusion = nullptr;
  float* d_topk_score = nullptr;
  int32_t* d_topk_idx = nullptr;
  void* scan_ws = nullptr;
  size_t scan_ws_bytes = 0;

  ~Model();
  int init(int vit_kind, int vit_depth, int qf_layers, int max_images, int max_queries, int max_pairs, int device);
  int alloc(void** p, size_t bytes);
  template <typename T>
  int alloc_t(T** p, size_t n) {
    return alloc(reinterpret_cast<void**>(p), n * sizeof(T));
  }
  int ensure_scan_ws(size_t bytes);

  // ---- weights ----
  void add_slot(const std::string& name, void* dst, int dtype, int64_t rows, int64_t cols, int64_t ld = 0,
                int64_t row_off = 0, bool required = true, bool flexible = false);
  int load_tensor(const char* name, int dtype, int ndim, const int64_t* shape, const void* data);
  int count_missing();

  // ---- forward passes ----
  // ViT + ln_vision: images fp32 [B,3,224,224] -> raws (fp32 and/or bf16 [B*257, Dv])
  int vit_forward(const float* images, int B, float* raws_f32, bf16* raws_bf16, cudaStream_t st);
  // packed cross-attention K/V of all cross layers for n_img images' raw embeds
  int cross_kv(const bf16* raws_bf16, int n_img, bool head_major, cudaStream_t st);
  // 12 BertLayers over B samples of S rows (S = 32: query rows only; S = 64: 32 query + 32 text rows).
  // with_enc: cross-attention + dual FFN (Qformer.py:435-468); else text FFN on every row (:469-475).
  // live_out: which rows of the LAST layer's output the caller reads (dead rows are not computed there).
  int qformer_layers(int B, int S, bool with_enc, int Lk, const int32_t* kv_idx0, const int32_t* kv_idx1,
                     const float* key_mask, int live_out, long long kv_rows, cudaStream_t st);
  int encode_gallery(const float* images, int B, float* feats_f32, bf16* feats_bf16, float* raws_f32,
                     bf16* raws_bf16, cudaStream_t st);
  int encode_query(const void* ref_raws, int ref_dtype, const int32_t* ref_rows, const int64_t* ids,
                   const int64_t* mask, int Bq, float* fusion_f32, bf16* fusion_bf16, cudaStream_t st);
  // Same result over the RAGGED row layout (attention_qfr.cu): only the live text rows of every caption are
  // computed.  text_len_host[b] = number of live tokens of caption b (sum of its attention mask), host memory.
  long long kv_table_rows = 0;   // rows / layout of the cross-attention K/V table the last cross_kv call wrote
  bool kv_head_major = false;
  int encode_query_ragged(const void* ref_raws, int ref_dtype, const int32_t* ref_rows, const int64_t* ids,
                          const int32_t* text_len_host, int Bq, float* fusion_f32, bf16* fusion_bf16, cudaStream_t st);
  // kv_idx0 / kv_idx1 (rerank): two-segment keys cat(ref, target) through sample index tables over plain K/V rows
  int qformer_layers_ragged(int B, int T8, bool with_enc, int Lk, const int32_t* kv_idx0, const int32_t* kv_idx1,
                            cudaStream_t st);
  // row tables of the ragged layout for B samples; sample b has lens_host[b / repeat] live text tokens
  int build_ragged_meta(const int32_t* lens_host, int repeat, int B, int* T8_out, cudaStream_t st);
  // text_len_host (optional, host int32 [R]): caption lengths -> ragged rows (live text tokens only)
  int rerank(const bf16* raws_table, const int32_t* ref_rows, const int32_t* cand_rows, const int64_t* ids,
             const int64_t* mask, const int32_t* text_len_host, int R, int T, float* p, cudaStream_t st);
};

}  // namespace sprc
