// C-ABI: model-level entry points (handle, weights, gallery/query encoders, scan, rerank).
// Each function documents, in include/sprc_b200.h, the reference function it stands in for.
#include <vector>
#include <new>

#include "../../include/sprc_b200.h"
#include "common.h"
#include "model.h"
#include "ops.h"

using namespace sprc;

struct sprc_handle {
  Model m;
  // sprc_query_topk_host_submit / _wait: one event per batch in flight (ring, oldest first)
  static constexpr int kInflight = 4;
  cudaEvent_t done[kInflight] = {nullptr, nullptr, nullptr, nullptr};
  unsigned submitted = 0, awaited = 0;
  // sprc_query_topk_strings_submit: pinned ids / mask / reference-row staging, one slot per batch in flight
  int64_t* tok_stage[kInflight] = {nullptr, nullptr, nullptr, nullptr};
};

namespace {
// scan workspace for handle-less calls (scan-only tools); grows on demand, freed at process exit
void* g_ws = nullptr;
size_t g_ws_bytes = 0;
int ensure_global_ws(size_t bytes) {
  if (bytes <= g_ws_bytes) return 0;
  if (g_ws) cudaFree(g_ws);
  g_ws = nullptr;
  g_ws_bytes = 0;
  if (cudaMalloc(&g_ws, bytes) != cudaSuccess) return set_error(-12, "cudaMalloc of %zu bytes failed", bytes);
  g_ws_bytes = bytes;
  return 0;
}
inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }
}  // namespace

namespace sprc {
bool ragged_query_enabled();   // model.cu (SPRC_RAGGED=0 disables the ragged query path)
}

extern "C" {

int sprc_create(const sprc_config* cfg, sprc_handle** out) {
  if (!cfg || !out) return set_error(-22, "sprc_create: null argument");
  *out = nullptr;
  sprc_handle* h = new (std::nothrow) sprc_handle();
  if (!h) return set_error(-12, "sprc_create: out of host memory");
  if (cfg->act_dtype != 0 && cfg->act_dtype != 1) {
    delete h;
    return set_error(-22, "sprc_create: act_dtype must be 0 (bf16) or 1 (fp16)");
  }
  set_act_fp16(cfg->act_dtype);  // process-wide mode: one handle per process (include/sprc_b200.h)
  int rc = h->m.init(cfg->vit_kind, cfg->vit_depth, cfg->qf_layers, cfg->max_images, cfg->max_queries,
                     cfg->max_pairs, cfg->device);
  if (rc != 0) {
    delete h;
    return rc;
  }
  *out = h;
  return 0;
}

void sprc_destroy(sprc_handle* h) {
  if (!h) return;
  for (cudaEvent_t ev : h->done)
    if (ev) cudaEventDestroy(ev);
  for (int64_t* p : h->tok_stage)
    if (p) cudaFreeHost(p);
  delete h;
}

int sprc_load_weights(sprc_handle* h, const sprc_tensor_desc* t, int n, int* n_missing) {
  if (!h || (!t && n > 0)) return set_error(-22, "sprc_load_weights: null argument");
  SPRC_CUDA(cudaSetDevice(h->m.device));
  for (int i = 0; i < n; ++i) {
    if (!t[i].name || !t[i].data) return set_error(-22, "sprc_load_weights: tensor %d has a null name or pointer", i);
    int rc = h->m.load_tensor(t[i].name, t[i].dtype, t[i].ndim, t[i].shape, t[i].data);
    if (rc < 0) return rc;
  }
  const int miss = h->m.count_missing();
  if (n_missing) *n_missing = miss;
  return 0;
}

const char* sprc_missing_weight(sprc_handle* h, int i) {
  if (!h || i < 0 || i >= (int)h->m.missing_cache.size()) return nullptr;
  return h->m.missing_cache[i].c_str();
}

int sprc_encode_gallery(sprc_handle* h, const float* images, int B, float* feats_f32, void* feats_bf16,
                        float* raws_f32, void* raws_bf16, void* stream) {
  if (!h || !images) return set_error(-22, "sprc_encode_gallery: null argument");
  if (h->m.count_missing() != 0)
    return set_error(-61, "sprc_encode_gallery: %d weights missing (first: %s)", (int)h->m.missing_cache.size(),
                     h->m.missing_cache[0].c_str());
  return h->m.encode_gallery(images, B, feats_f32, static_cast<bf16*>(feats_bf16), raws_f32,
                             static_cast<bf16*>(raws_bf16), S(stream));
}

int sprc_encode_query(sprc_handle* h, const void* ref_raws, int ref_dtype, const int32_t* ref_rows,
                      const int64_t* input_ids, const int64_t* attention_mask, int Bq, float* fusion_f32,
                      void* fusion_bf16, void* stream) {
  if (!h || !ref_raws || !input_ids || !attention_mask) return set_error(-22, "sprc_encode_query: null argument");
  if (h->m.count_missing() != 0)
    return set_error(-61, "sprc_encode_query: %d weights missing (first: %s)", (int)h->m.missing_cache.size(),
                     h->m.missing_cache[0].c_str());
  return h->m.encode_query(ref_raws, ref_dtype, ref_rows, input_ids, attention_mask, Bq, fusion_f32,
                           static_cast<bf16*>(fusion_bf16), S(stream));
}

int sprc_encode_query_lens(sprc_handle* h, const void* ref_raws, int ref_dtype, const int32_t* ref_rows,
                           const int64_t* input_ids, const int32_t* text_len_host, int Bq, float* fusion_f32,
                           void* fusion_bf16, void* stream) {
  if (!h || !ref_raws || !input_ids || !text_len_host) return set_error(-22, "sprc_encode_query_lens: null argument");
  if (h->m.count_missing() != 0)
    return set_error(-61, "sprc_encode_query_lens: %d weights missing (first: %s)", (int)h->m.missing_cache.size(),
                     h->m.missing_cache[0].c_str());
  return h->m.encode_query_ragged(ref_raws, ref_dtype, ref_rows, input_ids, text_len_host, Bq, fusion_f32,
                                  static_cast<bf16*>(fusion_bf16), S(stream));
}

static int sim_topk_api(sprc_handle* h, const void* queries, int Q, const void* gallery, int64_t N, int64_t row_offset,
                        int k, float* out_score, int32_t* out_idx, float* out_full, int group_rows,
                        int64_t group_stride, void* stream);

int sprc_sim_topk(sprc_handle* h, const void* queries, int Q, const void* gallery, int64_t N, int64_t row_offset,
                  int k, float* out_score, int32_t* out_idx, float* out_full, void* stream) {
  return sim_topk_api(h, queries, Q, gallery, N, row_offset, k, out_score, out_idx, out_full, 0, 0, stream);
}

int sprc_sim_topk_grouped(sprc_handle* h, const void* queries, int Q, const void* gallery, int64_t N,
                          int64_t row_offset, int k, float* out_score, int32_t* out_idx, int group_rows,
                          int64_t group_stride, void* stream) {
  if (!out_score || !out_idx || group_rows <= 0 || group_stride < (int64_t)group_rows * k)
    return set_error(-22, "sprc_sim_topk_grouped: outputs, group_rows > 0 and group_stride >= group_rows * k needed");
  return sim_topk_api(h, queries, Q, gallery, N, row_offset, k, out_score, out_idx, nullptr, group_rows, group_stride,
                      stream);
}

static int sim_topk_api(sprc_handle* h, const void* queries, int Q, const void* gallery, int64_t N, int64_t row_offset,
                        int k, float* out_score, int32_t* out_idx, float* out_full, int group_rows,
                        int64_t group_stride, void* stream) {
  if (!queries || !gallery) return set_error(-22, "sprc_sim_topk: null argument");
  const bool want_topk = out_score || out_idx;
  void* ws = nullptr;
  size_t ws_bytes = 0;
  if (want_topk) {
    const size_t need = sim_topk_workspace_bytes(Q, N, k, out_full != nullptr);
    if (h) {
      SPRC_TRY(h->m.ensure_scan_ws(need));
      ws = h->m.scan_ws;
      ws_bytes = h->m.scan_ws_bytes;
    } else {
      SPRC_TRY(ensure_global_ws(need));
      ws = g_ws;
      ws_bytes = g_ws_bytes;
    }
  }
  return sim_topk(static_cast<const bf16*>(queries), Q, static_cast<const bf16*>(gallery), N, row_offset, k,
                  out_score, out_idx, out_full, ws, ws_bytes, S(stream), group_rows, static_cast<size_t>(group_stride));
}

int sprc_topk_merge(sprc_handle*, const float* cand_score, const int32_t* cand_idx, int P, int Q, int k,
                    float* out_score, int32_t* out_idx, void* stream) {
  if (!cand_score || !cand_idx) return set_error(-22, "sprc_topk_merge: null argument");
  return topk_merge(cand_score, cand_idx, P, Q, k, out_score, out_idx, S(stream));
}

int sprc_topk_merge_packed(sprc_handle*, const int32_t* cand, int P, int Q, int k, float* out_score, int32_t* out_idx,
                           void* stream) {
  if (!cand) return set_error(-22, "sprc_topk_merge_packed: null argument");
  const size_t qk = static_cast<size_t>(Q) * k;
  return topk_merge(reinterpret_cast<const float*>(cand), cand + qk, P, Q, k, out_score, out_idx, S(stream), 2 * qk);
}

int sprc_gather_scores(sprc_handle*, const void* queries, int Q, const void* gallery, int64_t N,
                       const int32_t* rows, int m, float* out, void* stream) {
  if (!queries || !gallery || !rows || !out) return set_error(-22, "sprc_gather_scores: null argument");
  return gather_scores(static_cast<const bf16*>(queries), Q, static_cast<const bf16*>(gallery), N, rows, m, out,
                       S(stream));
}

int sprc_rerank(sprc_handle* h, const void* raws_bf16, const int32_t* ref_rows, const int32_t* cand_rows,
                const int64_t* input_ids, const int64_t* attention_mask, int R, int T, float* p, void* stream) {
  if (!h || !raws_bf16 || !ref_rows || !cand_rows || !input_ids || !attention_mask || !p)
    return set_error(-22, "sprc_rerank: null argument");
  for (const char* nm : {"itm_head.weight", "itm_head.bias"})
    if (!h->m.slots[nm].loaded) return set_error(-61, "sprc_rerank: %s was never loaded", nm);
  if (h->m.count_missing() != 0) return set_error(-61, "sprc_rerank: weights missing");
  return h->m.rerank(static_cast<const bf16*>(raws_bf16), ref_rows, cand_rows, input_ids, attention_mask, nullptr, R,
                     T, p, S(stream));
}

int sprc_rerank_lens(sprc_handle* h, const void* raws_bf16, const int32_t* ref_rows, const int32_t* cand_rows,
                     const int64_t* input_ids, const int32_t* text_len_host, int R, int T, float* p, void* stream) {
  if (!h || !raws_bf16 || !ref_rows || !cand_rows || !input_ids || !text_len_host || !p)
    return set_error(-22, "sprc_rerank_lens: null argument");
  for (const char* nm : {"itm_head.weight", "itm_head.bias"})
    if (!h->m.slots[nm].loaded) return set_error(-61, "sprc_rerank_lens: %s was never loaded", nm);
  if (h->m.count_missing() != 0) return set_error(-61, "sprc_rerank_lens: weights missing");
  SPRC_REQUIRE(ragged_query_enabled(), "sprc_rerank_lens: the ragged passes are disabled (SPRC_RAGGED=0)");
  return h->m.rerank(static_cast<const bf16*>(raws_bf16), ref_rows, cand_rows, input_ids, nullptr, text_len_host, R, T,
                     p, S(stream));
}

int sprc_query_topk_host_submit(sprc_handle* h, const void* raws_bf16, const void* gallery_bf16, int64_t N,
                                const int32_t* ref_rows_host, const int64_t* ids_host, const int64_t* mask_host, int Bq,
                                int k, float* out_score_host, int32_t* out_idx_host, void* stream) {
  if (!h || !raws_bf16 || !gallery_bf16 || !ref_rows_host || !ids_host || !mask_host || !out_score_host ||
      !out_idx_host)
    return set_error(-22, "sprc_query_topk_host: null argument");
  if (h->submitted - h->awaited >= (unsigned)sprc_handle::kInflight)
    return set_error(-11, "sprc_query_topk_host_submit: %d batches already in flight, call _wait first",
                     sprc_handle::kInflight);
  Model& m = h->m;
  SPRC_REQUIRE(Bq > 0 && Bq <= m.max_queries, "sprc_query_topk_host: Bq=%d outside (0, %d]", Bq, m.max_queries);
  SPRC_REQUIRE(k > 0 && k <= 256, "sprc_query_topk_host: k=%d outside [1, 256]", k);
  cudaStream_t st = S(stream);
  SPRC_CUDA(cudaMemcpyAsync(m.d_ids, ids_host, (size_t)Bq * 32 * 8, cudaMemcpyHostToDevice, st));
  SPRC_CUDA(cudaMemcpyAsync(m.d_rows, ref_rows_host, (size_t)Bq * 4, cudaMemcpyHostToDevice, st));
  // caption lengths from the host mask: prefix masks (tokens then padding), as BertTokenizer(padding="max_length") makes
  bool prefix = true;
  std::vector<int32_t> lens((size_t)Bq);
  for (int b = 0; b < Bq && prefix; ++b) {
    int L = 0;
    while (L < 32 && mask_host[(size_t)b * 32 + L] != 0) ++L;
    for (int t = L; t < 32; ++t) prefix = prefix && mask_host[(size_t)b * 32 + t] == 0;
    lens[b] = L;
  }
  if (prefix && ragged_query_enabled()) {
    SPRC_TRY(sprc_encode_query_lens(h, raws_bf16, SPRC_BF16, m.d_rows, m.d_ids, lens.data(), Bq, nullptr, m.d_fusion,
                                    stream));
  } else {
    SPRC_CUDA(cudaMemcpyAsync(m.d_mask, mask_host, (size_t)Bq * 32 * 8, cudaMemcpyHostToDevice, st));
    SPRC_TRY(sprc_encode_query(h, raws_bf16, SPRC_BF16, m.d_rows, m.d_ids, m.d_mask, Bq, nullptr, m.d_fusion, stream));
  }
  SPRC_TRY(sprc_sim_topk(h, m.d_fusion, Bq, gallery_bf16, N, 0, k, m.d_topk_score, m.d_topk_idx, nullptr, stream));
  SPRC_CUDA(cudaMemcpyAsync(out_score_host, m.d_topk_score, (size_t)Bq * k * 4, cudaMemcpyDeviceToHost, st));
  SPRC_CUDA(cudaMemcpyAsync(out_idx_host, m.d_topk_idx, (size_t)Bq * k * 4, cudaMemcpyDeviceToHost, st));
  if (!h->done[0])
    for (cudaEvent_t& e : h->done) SPRC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  cudaEvent_t& ev = h->done[h->submitted % sprc_handle::kInflight];
  SPRC_CUDA(cudaEventRecord(ev, st));
  ++h->submitted;
  return 0;
}

int sprc_query_topk_strings_submit(sprc_handle* h, const sprc_tokenizer* tok, const void* raws_bf16,
                                   const void* gallery_bf16, int64_t N, const int32_t* ref_rows_host, const char* texts,
                                   const int64_t* offsets, int Bq, int k, int threads, float* out_score_host,
                                   int32_t* out_idx_host, void* stream) {
  if (!h || !tok || !texts || !offsets || !ref_rows_host)
    return set_error(-22, "sprc_query_topk_strings_submit: null argument");
  if (h->submitted - h->awaited >= (unsigned)sprc_handle::kInflight)
    return set_error(-11, "sprc_query_topk_strings_submit: %d batches already in flight, call _wait first",
                     sprc_handle::kInflight);
  Model& m = h->m;
  SPRC_REQUIRE(Bq > 0 && Bq <= m.max_queries, "sprc_query_topk_strings: Bq=%d outside (0, %d]", Bq, m.max_queries);
  // the slot's previous batch (submitted - kInflight) has been awaited, so its H2D copies are done
  const size_t row = (size_t)m.max_queries * 32;
  if (!h->tok_stage[0])   // all slots at the first call (page-locked allocation synchronises the device)
    for (int64_t*& p : h->tok_stage)
      SPRC_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&p), (2 * row + m.max_queries) * 8, cudaHostAllocDefault));
  int64_t*& stage = h->tok_stage[h->submitted % sprc_handle::kInflight];
  int64_t* ids = stage;
  int64_t* mask = stage + row;
  int32_t* rows = reinterpret_cast<int32_t*>(stage + 2 * row);
  std::vector<uint8_t> flags((size_t)Bq);
  SPRC_TRY(sprc_tokenize_host(tok, texts, offsets, Bq, 32, threads, ids, mask, nullptr, flags.data()));
  for (int i = 0; i < Bq; ++i)
    if (flags[i])
      return set_error(-84, "sprc_query_topk_strings: caption %d holds neighbour-dependent characters (combining "
                            "marks / final sigma); tokenise this batch on the host and use sprc_query_topk_host_submit", i);
  for (int i = 0; i < Bq; ++i) rows[i] = ref_rows_host[i];   // the caller's buffer need not outlive this call
  return sprc_query_topk_host_submit(h, raws_bf16, gallery_bf16, N, rows, ids, mask, Bq, k, out_score_host,
                                     out_idx_host, stream);
}

int sprc_query_topk_host_wait(sprc_handle* h) {
  if (!h) return set_error(-22, "sprc_query_topk_host_wait: null handle");
  if (h->awaited == h->submitted) return set_error(-22, "sprc_query_topk_host_wait: nothing in flight");
  SPRC_CUDA(cudaEventSynchronize(h->done[h->awaited % sprc_handle::kInflight]));
  ++h->awaited;
  return 0;
}

int sprc_query_topk_host(sprc_handle* h, const void* raws_bf16, const void* gallery_bf16, int64_t N,
                         const int32_t* ref_rows_host, const int64_t* ids_host, const int64_t* mask_host, int Bq,
                         int k, float* out_score_host, int32_t* out_idx_host, void* stream) {
  if (h && h->submitted != h->awaited)
    return set_error(-11, "sprc_query_topk_host: %u submitted batches are still in flight", h->submitted - h->awaited);
  SPRC_TRY(sprc_query_topk_host_submit(h, raws_bf16, gallery_bf16, N, ref_rows_host, ids_host, mask_host, Bq, k,
                                       out_score_host, out_idx_host, stream));
  return sprc_query_topk_host_wait(h);
}

}  // extern "C"
