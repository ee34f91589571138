// Model handle: weight packing (reference state-dict keys -> kernel layouts) and the forward passes.
//
// Reference behaviour restated here (paths relative to /root/reference/src/lavis/models):
//   ViT          eva_vit.py:118-148,173-180,324-340 (EVA-g) / clip_vit.py:114-139,171-185 (CLIP-L)
//   ln_vision    blip2_models/blip2.py:81,193-199
//   Q-Former     blip2_models/Qformer.py:78-114 (embeddings), 175-281 (attention), 408-480 (layer routing)
//   heads        blip2_models/blip2_qformer_cir_align_prompt.py:348-350,385 ; rerank: blip2_qformer_cir_rerank.py:399-445
// Data layout in HBM: activations are row-major [tokens, features]; the residual stream and every
// LayerNorm input are fp32, GEMM operands are bf16 copies written by the producing kernel's epilogue.
#include <vector>
#include <stdlib.h>
#include "model.h"

#include <string.h>

#include "../../include/sprc_b200.h"

namespace sprc {

Model::~Model() {
  for (void* p : allocs) cudaFree(p);
  if (staging) cudaFree(staging);
  if (scan_ws) cudaFree(scan_ws);
  for (int i = 0; i < kMetaRing; ++i) {
    if (h_meta_pin[i]) cudaFreeHost(h_meta_pin[i]);
    if (meta_ev[i]) cudaEventDestroy(meta_ev[i]);
  }
}

int Model::alloc(void** p, size_t bytes) {
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) {
    *p = nullptr;
    return set_error(-12, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
  }
  allocs.push_back(*p);
  return 0;
}

int Model::ensure_scan_ws(size_t bytes) {
  if (bytes <= scan_ws_bytes) return 0;
  if (scan_ws) cudaFree(scan_ws);
  scan_ws = nullptr;
  scan_ws_bytes = 0;
  cudaError_t e = cudaMalloc(&scan_ws, bytes);
  if (e != cudaSuccess) return set_error(-12, "cudaMalloc of %zu bytes (scan workspace) failed", bytes);
  scan_ws_bytes = bytes;
  return 0;
}

void Model::add_slot(const std::string& name, void* dst, int dtype, int64_t rows, int64_t cols, int64_t ld,
                     int64_t row_off, bool required, bool flexible) {
  WeightSlot s;
  s.dst = dst;
  s.dst_dtype = dtype;
  s.rows = rows;
  s.cols = cols;
  s.ld = ld ? ld : cols;
  s.row_off = row_off;
  s.required = required;
  s.flexible_rows = flexible;
  slots[name] = s;
}

int Model::init(int kind, int vit_depth, int qf_layers_, int max_images_, int max_queries_, int max_pairs_,
                int device_) {
  SPRC_REQUIRE(kind == SPRC_VIT_EVA_G || kind == SPRC_VIT_CLIP_L, "unknown vit_kind %d", kind);
  device = device_;
  act_dtype_fp16 = act_fp16();
  SPRC_CUDA(cudaSetDevice(device));
  {
    cudaDeviceProp prop;
    SPRC_CUDA(cudaGetDeviceProperties(&prop, device));
    SPRC_REQUIRE(prop.major == 10, "libsprc_b200 is built for sm_100a only; device %d is sm_%d%d", device,
                 prop.major, prop.minor);
  }
  vit_kind = kind;
  if (kind == SPRC_VIT_EVA_G) {
    Dv = 1408, depth = 39, dh = 88, mlp = 6144, vit_eps = 1e-6f, vit_act = ACT_GELU;
  } else {
    Dv = 1024, depth = 23, dh = 64, mlp = 4096, vit_eps = 1e-5f, vit_act = ACT_QUICKGELU;
  }
  if (vit_depth > 0) depth = vit_depth;
  qf_layers = qf_layers_ > 0 ? qf_layers_ : 12;
  n_cross = (qf_layers + 1) / 2;  // cross_attention_freq = 2 -> layers 0,2,4,...
  max_images = max_images_ > 0 ? max_images_ : 1;
  max_queries = max_queries_ > 0 ? max_queries_ : 1;
  max_pairs = max_pairs_ > 0 ? max_pairs_ : 0;

  // ------------------------------ weights ------------------------------
  const std::string ve = "visual_encoder.";
  SPRC_TRY(alloc_t(&cls, Dv));
  SPRC_TRY(alloc_t(&pos, (size_t)257 * Dv));
  SPRC_TRY(alloc_t(&patch_w, (size_t)Dv * KP));
  SPRC_CUDA(cudaMemset(patch_w, 0, (size_t)Dv * KP * sizeof(bf16)));
  blocks.resize(depth);
  if (kind == SPRC_VIT_EVA_G) {
    SPRC_TRY(alloc_t(&patch_b, Dv));
    add_slot(ve + "cls_token", cls, SPRC_F32, 1, Dv);
    add_slot(ve + "pos_embed", pos, SPRC_F32, 257, Dv);
    add_slot(ve + "patch_embed.proj.weight", patch_w, SPRC_BF16, Dv, 588, KP);
    add_slot(ve + "patch_embed.proj.bias", patch_b, SPRC_F32, 1, Dv);
  } else {
    SPRC_TRY(alloc_t(&ln_pre_g, Dv));
    SPRC_TRY(alloc_t(&ln_pre_b, Dv));
    add_slot(ve + "class_embedding", cls, SPRC_F32, 1, Dv);
    add_slot(ve + "positional_embedding", pos, SPRC_F32, 257, Dv);
    add_slot(ve + "conv1.weight", patch_w, SPRC_BF16, Dv, 588, KP);
    add_slot(ve + "ln_pre.weight", ln_pre_g, SPRC_F32, 1, Dv);
    add_slot(ve + "ln_pre.bias", ln_pre_b, SPRC_F32, 1, Dv);
  }
  for (int i = 0; i < depth; ++i) {
    VitBlock& b = blocks[i];
    SPRC_TRY(alloc_t(&b.ln1_g, Dv));
    SPRC_TRY(alloc_t(&b.ln1_b, Dv));
    SPRC_TRY(alloc_t(&b.ln2_g, Dv));
    SPRC_TRY(alloc_t(&b.ln2_b, Dv));
    SPRC_TRY(alloc_t(&b.qkv_w, (size_t)3 * Dv * Dv));
    SPRC_TRY(alloc_t(&b.qkv_b, (size_t)3 * Dv));
    SPRC_TRY(alloc_t(&b.proj_w, (size_t)Dv * Dv));
    SPRC_TRY(alloc_t(&b.proj_b, Dv));
    SPRC_TRY(alloc_t(&b.fc1_w, (size_t)mlp * Dv));
    SPRC_TRY(alloc_t(&b.fc1_b, mlp));
    SPRC_TRY(alloc_t(&b.fc2_w, (size_t)Dv * mlp));
    SPRC_TRY(alloc_t(&b.fc2_b, Dv));
    if (kind == SPRC_VIT_EVA_G) {
      const std::string p = ve + "blocks." + std::to_string(i) + ".";
      // qkv bias = cat(q_bias, zeros, v_bias)  (eva_vit.py:122)
      SPRC_CUDA(cudaMemset(b.qkv_b, 0, (size_t)3 * Dv * sizeof(float)));
      add_slot(p + "norm1.weight", b.ln1_g, SPRC_F32, 1, Dv);
      add_slot(p + "norm1.bias", b.ln1_b, SPRC_F32, 1, Dv);
      add_slot(p + "norm2.weight", b.ln2_g, SPRC_F32, 1, Dv);
      add_slot(p + "norm2.bias", b.ln2_b, SPRC_F32, 1, Dv);
      add_slot(p + "attn.q_bias", b.qkv_b, SPRC_F32, 1, Dv);
      add_slot(p + "attn.v_bias", b.qkv_b + 2 * Dv, SPRC_F32, 1, Dv);
      add_slot(p + "attn.qkv.weight", b.qkv_w, SPRC_BF16, 3 * Dv, Dv);
      add_slot(p + "attn.proj.weight", b.proj_w, SPRC_BF16, Dv, Dv);
      add_slot(p + "attn.proj.bias", b.proj_b, SPRC_F32, 1, Dv);
      add_slot(p + "mlp.fc1.weight", b.fc1_w, SPRC_BF16, mlp, Dv);
      add_slot(p + "mlp.fc1.bias", b.fc1_b, SPRC_F32, 1, mlp);
      add_slot(p + "mlp.fc2.weight", b.fc2_w, SPRC_BF16, Dv, mlp);
      add_slot(p + "mlp.fc2.bias", b.fc2_b, SPRC_F32, 1, Dv);
    } else {
      const std::string p = ve + "transformer.resblocks." + std::to_string(i) + ".";
      add_slot(p + "ln_1.weight", b.ln1_g, SPRC_F32, 1, Dv);
      add_slot(p + "ln_1.bias", b.ln1_b, SPRC_F32, 1, Dv);
      add_slot(p + "ln_2.weight", b.ln2_g, SPRC_F32, 1, Dv);
      add_slot(p + "ln_2.bias", b.ln2_b, SPRC_F32, 1, Dv);
      add_slot(p + "attn.in_proj_weight", b.qkv_w, SPRC_BF16, 3 * Dv, Dv);
      add_slot(p + "attn.in_proj_bias", b.qkv_b, SPRC_F32, 1, 3 * Dv);
      add_slot(p + "attn.out_proj.weight", b.proj_w, SPRC_BF16, Dv, Dv);
      add_slot(p + "attn.out_proj.bias", b.proj_b, SPRC_F32, 1, Dv);
      add_slot(p + "mlp.c_fc.weight", b.fc1_w, SPRC_BF16, mlp, Dv);
      add_slot(p + "mlp.c_fc.bias", b.fc1_b, SPRC_F32, 1, mlp);
      add_slot(p + "mlp.c_proj.weight", b.fc2_w, SPRC_BF16, Dv, mlp);
      add_slot(p + "mlp.c_proj.bias", b.fc2_b, SPRC_F32, 1, Dv);
    }
  }
  SPRC_TRY(alloc_t(&lnv_g, Dv));
  SPRC_TRY(alloc_t(&lnv_b, Dv));
  add_slot("ln_vision.weight", lnv_g, SPRC_F32, 1, Dv);
  add_slot("ln_vision.bias", lnv_b, SPRC_F32, 1, Dv);

  SPRC_TRY(alloc_t(&query_tokens, 32 * 768));
  SPRC_TRY(alloc_t(&word_emb, (size_t)30523 * 768));
  SPRC_TRY(alloc_t(&pos_emb, (size_t)512 * 768));
  SPRC_TRY(alloc_t(&emb_g, 768));
  SPRC_TRY(alloc_t(&emb_b, 768));
  add_slot("query_tokens", query_tokens, SPRC_F32, 32, 768);
  const std::string qb = "Qformer.bert.";
  add_slot(qb + "embeddings.word_embeddings.weight", word_emb, SPRC_F32, 30523, 768, 768, 0, true, true);
  add_slot(qb + "embeddings.position_embeddings.weight", pos_emb, SPRC_F32, 512, 768);
  add_slot(qb + "embeddings.LayerNorm.weight", emb_g, SPRC_F32, 1, 768);
  add_slot(qb + "embeddings.LayerNorm.bias", emb_b, SPRC_F32, 1, 768);
  SPRC_TRY(alloc_t(&kv_w, (size_t)n_cross * 1536 * Dv));
  SPRC_TRY(alloc_t(&kv_b, (size_t)n_cross * 1536));
  layers.resize(qf_layers);
  for (int l = 0; l < qf_layers; ++l) {
    QfLayer& L = layers[l];
    const std::string p = qb + "encoder.layer." + std::to_string(l) + ".";
    SPRC_TRY(alloc_t(&L.qkv_w, (size_t)2304 * 768));
    SPRC_TRY(alloc_t(&L.qkv_b, 2304));
    SPRC_TRY(alloc_t(&L.so_w, (size_t)768 * 768));
    SPRC_TRY(alloc_t(&L.so_b, 768));
    SPRC_TRY(alloc_t(&L.so_g, 768));
    SPRC_TRY(alloc_t(&L.so_beta, 768));
    const char* qkvn[3] = {"query", "key", "value"};
    for (int j = 0; j < 3; ++j) {
      add_slot(p + "attention.self." + qkvn[j] + ".weight", L.qkv_w, SPRC_BF16, 768, 768, 768, j * 768);
      add_slot(p + "attention.self." + qkvn[j] + ".bias", L.qkv_b + j * 768, SPRC_F32, 1, 768);
    }
    add_slot(p + "attention.output.dense.weight", L.so_w, SPRC_BF16, 768, 768);
    add_slot(p + "attention.output.dense.bias", L.so_b, SPRC_F32, 1, 768);
    add_slot(p + "attention.output.LayerNorm.weight", L.so_g, SPRC_F32, 1, 768);
    add_slot(p + "attention.output.LayerNorm.bias", L.so_beta, SPRC_F32, 1, 768);
    L.has_cross = (l % 2 == 0);
    if (L.has_cross) {
      const int ci = l / 2;
      SPRC_TRY(alloc_t(&L.cq_w, (size_t)768 * 768));
      SPRC_TRY(alloc_t(&L.cq_b, 768));
      SPRC_TRY(alloc_t(&L.co_w, (size_t)768 * 768));
      SPRC_TRY(alloc_t(&L.co_b, 768));
      SPRC_TRY(alloc_t(&L.co_g, 768));
      SPRC_TRY(alloc_t(&L.co_beta, 768));
      add_slot(p + "crossattention.self.query.weight", L.cq_w, SPRC_BF16, 768, 768);
      add_slot(p + "crossattention.self.query.bias", L.cq_b, SPRC_F32, 1, 768);
      add_slot(p + "crossattention.self.key.weight", kv_w, SPRC_BF16, 768, Dv, Dv, (int64_t)ci * 1536);
      add_slot(p + "crossattention.self.key.bias", kv_b + ci * 1536, SPRC_F32, 1, 768);
      add_slot(p + "crossattention.self.value.weight", kv_w, SPRC_BF16, 768, Dv, Dv, (int64_t)ci * 1536 + 768);
      add_slot(p + "crossattention.self.value.bias", kv_b + ci * 1536 + 768, SPRC_F32, 1, 768);
      add_slot(p + "crossattention.output.dense.weight", L.co_w, SPRC_BF16, 768, 768);
      add_slot(p + "crossattention.output.dense.bias", L.co_b, SPRC_F32, 1, 768);
      add_slot(p + "crossattention.output.LayerNorm.weight", L.co_g, SPRC_F32, 1, 768);
      add_slot(p + "crossattention.output.LayerNorm.bias", L.co_beta, SPRC_F32, 1, 768);
    }
    SPRC_TRY(alloc_t(&L.ti_w, (size_t)3072 * 768));
    SPRC_TRY(alloc_t(&L.ti_b, 3072));
    SPRC_TRY(alloc_t(&L.to_w, (size_t)768 * 3072));
    SPRC_TRY(alloc_t(&L.to_b, 768));
    SPRC_TRY(alloc_t(&L.to_g, 768));
    SPRC_TRY(alloc_t(&L.to_beta, 768));
    SPRC_TRY(alloc_t(&L.qi_w, (size_t)3072 * 768));
    SPRC_TRY(alloc_t(&L.qi_b, 3072));
    SPRC_TRY(alloc_t(&L.qo_w, (size_t)768 * 3072));
    SPRC_TRY(alloc_t(&L.qo_b, 768));
    SPRC_TRY(alloc_t(&L.qo_g, 768));
    SPRC_TRY(alloc_t(&L.qo_beta, 768));
    add_slot(p + "intermediate.dense.weight", L.ti_w, SPRC_BF16, 3072, 768);
    add_slot(p + "intermediate.dense.bias", L.ti_b, SPRC_F32, 1, 3072);
    add_slot(p + "output.dense.weight", L.to_w, SPRC_BF16, 768, 3072);
    add_slot(p + "output.dense.bias", L.to_b, SPRC_F32, 1, 768);
    add_slot(p + "output.LayerNorm.weight", L.to_g, SPRC_F32, 1, 768);
    add_slot(p + "output.LayerNorm.bias", L.to_beta, SPRC_F32, 1, 768);
    add_slot(p + "intermediate_query.dense.weight", L.qi_w, SPRC_BF16, 3072, 768);
    add_slot(p + "intermediate_query.dense.bias", L.qi_b, SPRC_F32, 1, 3072);
    add_slot(p + "output_query.dense.weight", L.qo_w, SPRC_BF16, 768, 3072);
    add_slot(p + "output_query.dense.bias", L.qo_b, SPRC_F32, 1, 768);
    add_slot(p + "output_query.LayerNorm.weight", L.qo_g, SPRC_F32, 1, 768);
    add_slot(p + "output_query.LayerNorm.bias", L.qo_beta, SPRC_F32, 1, 768);
  }
  SPRC_TRY(alloc_t(&vproj_w, 256 * 768));
  SPRC_TRY(alloc_t(&vproj_b, 256));
  SPRC_TRY(alloc_t(&tproj_w, 256 * 768));
  SPRC_TRY(alloc_t(&tproj_b, 256));
  SPRC_TRY(alloc_t(&itm_w, 2 * 768));
  SPRC_TRY(alloc_t(&itm_b, 2));
  add_slot("vision_proj.weight", vproj_w, SPRC_BF16, 256, 768);
  add_slot("vision_proj.bias", vproj_b, SPRC_F32, 1, 256);
  add_slot("text_proj.weight", tproj_w, SPRC_BF16, 256, 768);
  add_slot("text_proj.bias", tproj_b, SPRC_F32, 1, 256);
  add_slot("itm_head.weight", itm_w, SPRC_F32, 2, 768, 768, 0, /*required=*/false);
  add_slot("itm_head.bias", itm_b, SPRC_F32, 1, 2, 2, 0, /*required=*/false);

  // ------------------------------ workspace ------------------------------
  vit_cap = max_images;
  enc_cap = max_images;
  if (max_queries > enc_cap) enc_cap = max_queries;
  if (2 * max_pairs > enc_cap) enc_cap = 2 * max_pairs;
  qf_rows = max_images * 32;
  if (max_queries * 64 > qf_rows) qf_rows = max_queries * 64;
  if (max_pairs * 64 > qf_rows) qf_rows = max_pairs * 64;
  const size_t T = (size_t)vit_cap * 257;
  SPRC_TRY(alloc_t(&patches, (size_t)vit_cap * 256 * KP));
  SPRC_TRY(alloc_t(&patch_out, (size_t)vit_cap * 256 * Dv));
  SPRC_TRY(alloc_t(&x, T * Dv));
  SPRC_TRY(alloc_t(&xn, T * Dv));
  SPRC_TRY(alloc_t(&qkv, T * 3 * Dv));
  SPRC_TRY(alloc_t(&att, T * Dv));
  SPRC_TRY(alloc_t(&h1, T * mlp));
  SPRC_TRY(alloc_t(&raws, (size_t)enc_cap * 257 * Dv));
  SPRC_TRY(alloc_t(&kv, (size_t)enc_cap * 257 * n_cross * 1536));
  const size_t R = (size_t)qf_rows;
  SPRC_TRY(alloc_t(&qh, R * 768));
  SPRC_TRY(alloc_t(&qt, R * 768));
  SPRC_TRY(alloc_t(&qhb, R * 768));
  SPRC_TRY(alloc_t(&qqkv, R * 2304));
  SPRC_TRY(alloc_t(&qctx, R * 768));
  SPRC_TRY(alloc_t(&qcq, R * 768));
  SPRC_TRY(alloc_t(&qffn, R * 3072));
  SPRC_TRY(alloc_t(&qproj, R * 256));
  SPRC_CUDA(cudaMemset(qctx, 0, R * 768 * sizeof(bf16)));
  SPRC_CUDA(cudaMemset(qcq, 0, R * 768 * sizeof(bf16)));
  SPRC_CUDA(cudaMemset(qffn, 0, R * 3072 * sizeof(bf16)));
  const size_t nq = (size_t)(max_queries > max_pairs ? max_queries : max_pairs);
  SPRC_TRY(alloc_t(&qmask, nq * 64));
  SPRC_TRY(alloc_t(&d_ids, nq * 32));
  SPRC_TRY(alloc_t(&d_mask, nq * 32));
  SPRC_TRY(alloc_t(&d_rows, nq + 16));
  SPRC_TRY(alloc_t(&d_rows2, nq + 16));
  meta_cap = nq * 40 + 64;
  SPRC_TRY(alloc_t(&d_meta, meta_cap));
  SPRC_TRY(alloc_t(&d_fusion, nq * 256));
  SPRC_TRY(alloc_t(&d_topk_score, nq * 256));
  SPRC_TRY(alloc_t(&d_topk_idx, nq * 256));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// weight loading
// ------------------------------------------------------------------------------------------------
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
pack2d_kernel(const TIn* __restrict__ src, int64_t rows, int64_t cols, TOut* __restrict__ dst, int64_t ld) {
  const int64_t n = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i % cols;
    dst[r * ld + c] = static_cast<TOut>(static_cast<float>(src[i]));
  }
}

template <typename TIn, typename TOut>
static void launch_pack(const void* src, int64_t rows, int64_t cols, void* dst, int64_t ld) {
  const int64_t n = rows * cols;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  pack2d_kernel<TIn, TOut><<<(int)blocks, 256>>>(static_cast<const TIn*>(src), rows, cols, static_cast<TOut*>(dst), ld);
}

int Model::load_tensor(const char* name, int dtype, int ndim, const int64_t* shape, const void* data) {
  auto it = slots.find(name);
  if (it == slots.end()) return 1;  // unknown key: ignored (strict=False)
  WeightSlot& s = it->second;
  SPRC_REQUIRE(dtype == SPRC_F32 || dtype == SPRC_F16 || dtype == SPRC_BF16, "%s: unsupported dtype %d", name,
               dtype);
  int64_t numel = 1;
  for (int i = 0; i < ndim; ++i) numel *= shape[i];
  // layout, not just the element count: a transposed [cols, rows] matrix has the right numel and the wrong meaning
  // (leading singleton dimensions are ignored: cls_token [1,1,Dv], pos_embed [1,257,Dv], query_tokens [1,32,768])
  {
    int first = 0;
    while (first < ndim - 1 && shape[first] == 1) ++first;
    if (s.rows > 1 && ndim - first >= 2) {
      int64_t inner = 1;
      for (int i = first + 1; i < ndim; ++i) inner *= shape[i];
      SPRC_REQUIRE(inner == s.cols && (s.flexible_rows || shape[first] == s.rows),
                   "%s: expected a [%lld, %lld] matrix, got leading dimension %lld and %lld elements per row", name,
                   (long long)s.rows, (long long)s.cols, (long long)shape[first], (long long)inner);
    }
  }
  int64_t rows = s.rows;
  if (s.flexible_rows) {
    SPRC_REQUIRE(numel % s.cols == 0 && numel / s.cols <= s.rows && numel > 0, "%s: bad shape (numel %lld)", name,
                 (long long)numel);
    rows = numel / s.cols;
    if (strstr(name, "word_embeddings")) vocab = (int)rows;
  } else {
    SPRC_REQUIRE(numel == s.rows * s.cols, "%s: expected %lld x %lld = %lld elements, got %lld", name,
                 (long long)s.rows, (long long)s.cols, (long long)(s.rows * s.cols), (long long)numel);
  }
  const size_t esz = dtype == SPRC_F32 ? 4 : 2;
  const size_t bytes = (size_t)numel * esz;
  if (bytes > staging_bytes) {
    if (staging) cudaFree(staging);
    staging = nullptr;
    staging_bytes = 0;
    SPRC_CUDA(cudaMalloc(&staging, bytes));
    staging_bytes = bytes;
  }
  SPRC_CUDA(cudaMemcpy(staging, data, bytes, cudaMemcpyDefault));
  const size_t dsz = s.dst_dtype == SPRC_F32 ? 4 : 2;
  void* dst = static_cast<char*>(s.dst) + (size_t)s.row_off * s.ld * dsz;
  if (s.dst_dtype == SPRC_F32) {
    if (dtype == SPRC_F32)
      launch_pack<float, float>(staging, rows, s.cols, dst, s.ld);
    else if (dtype == SPRC_F16)
      launch_pack<__half, float>(staging, rows, s.cols, dst, s.ld);
    else
      launch_pack<bf16, float>(staging, rows, s.cols, dst, s.ld);
  } else {
    // 16-bit GEMM operands are stored in the library's active format (bf16, or fp16 in fp16 mode)
    if (act_fp16()) {
      if (dtype == SPRC_F32)
        launch_pack<float, __half>(staging, rows, s.cols, dst, s.ld);
      else if (dtype == SPRC_F16)
        launch_pack<__half, __half>(staging, rows, s.cols, dst, s.ld);
      else
        launch_pack<bf16, __half>(staging, rows, s.cols, dst, s.ld);
    } else if (dtype == SPRC_F32)
      launch_pack<float, bf16>(staging, rows, s.cols, dst, s.ld);
    else if (dtype == SPRC_F16)
      launch_pack<__half, bf16>(staging, rows, s.cols, dst, s.ld);
    else
      launch_pack<bf16, bf16>(staging, rows, s.cols, dst, s.ld);
  }
  SPRC_CUDA(cudaGetLastError());
  SPRC_CUDA(cudaDeviceSynchronize());
  s.loaded = true;
  return 0;
}

int Model::count_missing() {
  missing_cache.clear();
  for (auto& kv_ : slots)
    if (kv_.second.required && !kv_.second.loaded) missing_cache.push_back(kv_.first);
  return (int)missing_cache.size();
}

// ------------------------------------------------------------------------------------------------
// forward passes
// ------------------------------------------------------------------------------------------------
static int linear(const bf16* A, int M, int K, int lda, const bf16* W, int N, const float* bias, int act,
                  const float* residual, float* out_f32, bf16* out_bf16, int ldc, int grp_rows, int grp_stride,
                  cudaStream_t st) {
  GemmDesc d;
  d.A = A;
  d.W = W;
  d.M = M;
  d.N = N;
  d.K = K;
  d.lda = lda;
  d.ldw = K;
  d.grp_rows = grp_rows;
  d.grp_stride = grp_stride;
  d.bias = bias;
  d.residual = residual;
  d.out_f32 = out_f32;
  d.out_bf16 = out_bf16;
  d.ldc = ldc;
  d.act = act;
  return gemm_bf16_tcgen05(d, st);
}

// dense + residual + LayerNorm of a post-LN Q-Former sublayer (Qformer.py:291-295, 373-381): x (fp32, in place) and
// xb (16-bit copy) <- LayerNorm(A W^T + bias + x): GEMM with TMA reduce-add into x, then the LayerNorm kernel.
// Fused forms were built, measured on B200 and removed: a 3-CTA cluster kernel exchanging row
// statistics through DSMEM (profiles/r01c_gemm_ln_*), and LayerNorm folded algebraically into the neighbouring GEMMs'
// epilogues, twice (profiles/r02a_*: LayerNorm 7.3 -> 0.8 ms per step but GEMMs 43.7 -> 61.2 ms; profiles/r03a_*, commit
// 649ae55, with a TMA-prefetched producer epilogue and a statistics warp: GEMMs 44.3 -> 50.5 ms, 37.3k vs 37.2k q/s).
static bool dual_ffn_enabled() {
  static const bool on = [] {
    const char* e = getenv("SPRC_DUAL_FFN");   // SPRC_DUAL_FFN=0: separate launches for query-row and text-row FFNs
    return !(e && e[0] == '0');
  }();
  return on;
}
static int linear_ln(const bf16* A, int M, int K, int lda, const bf16* W, const float* bias, const float* gamma,
                     const float* beta, float eps, float* x, bf16* xb, int grp_rows, int grp_stride, cudaStream_t st) {
  SPRC_TRY(linear(A, M, K, lda, W, 768, bias, ACT_NONE, x, x, nullptr, 768, grp_rows, grp_stride, st));
  return layernorm(x, M, 768, gamma, beta, eps, grp_rows, grp_stride, x, xb, st);
}

int Model::vit_forward(const float* images, int B, float* raws_f32, bf16* raws_bf16, cudaStream_t st) {
  SPRC_REQUIRE(B > 0 && B <= vit_cap, "vit_forward: B=%d outside (0, %d]", B, vit_cap);
  const int T = B * 257;
  SPRC_TRY(im2col_patches(images, B, patches, KP, st));
  SPRC_TRY(linear(patches, B * 256, KP, KP, patch_w, Dv, patch_b, ACT_NONE, nullptr, patch_out, nullptr, Dv, 0, 0,
                  st));
  SPRC_TRY(vit_assemble_tokens(patch_out, cls, pos, B, Dv, x, st));
  if (vit_kind == SPRC_VIT_CLIP_L) SPRC_TRY(layernorm(x, T, Dv, ln_pre_g, ln_pre_b, 1e-5f, 0, 0, x, nullptr, st));
  const float scale = 1.0f / sqrtf((float)dh);
  for (int i = 0; i < depth; ++i) {
    const VitBlock& b = blocks[i];
    SPRC_TRY(layernorm(x, T, Dv, b.ln1_g, b.ln1_b, vit_eps, 0, 0, nullptr, xn, st));
    SPRC_TRY(linear(xn, T, Dv, Dv, b.qkv_w, 3 * Dv, b.qkv_b, ACT_NONE, nullptr, nullptr, qkv, 3 * Dv, 0, 0, st));
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
    SPRC_TRY(linear(att, T, Dv, Dv, b.proj_w, Dv, b.proj_b, ACT_NONE, x, x, nullptr, Dv, 0, 0, st));
    SPRC_TRY(layernorm(x, T, Dv, b.ln2_g, b.ln2_b, vit_eps, 0, 0, nullptr, xn, st));
    SPRC_TRY(linear(xn, T, Dv, Dv, b.fc1_w, mlp, b.fc1_b, vit_act, nullptr, nullptr, h1, mlp, 0, 0, st));
    SPRC_TRY(linear(h1, T, mlp, mlp, b.fc2_w, Dv, b.fc2_b, ACT_NONE, x, x, nullptr, Dv, 0, 0, st));
  }
  SPRC_TRY(layernorm(x, T, Dv, lnv_g, lnv_b, 1e-5f, 0, 0, raws_f32, raws_bf16, st));
  return 0;
}

// K/V projections of all cross-attention layers in one GEMM.  head_major: every (layer, K|V, head) block is written as
// a contiguous [n_img * 257, 64] matrix (GemmDesc::out_col_block), which the tcgen05 cross-attention kernel reads as
// whole 33 KB blocks per (sample, head) instead of 128-byte pieces of 18 KB-pitch rows (all callers use it now: the
// rerank path addresses the blocks of its reference / candidate images through image index tables).
int Model::cross_kv(const bf16* raws_bf16, int n_img, bool head_major, cudaStream_t st) {
  SPRC_REQUIRE(n_img > 0 && n_img <= enc_cap, "cross_kv: %d images outside (0, %d]", n_img, enc_cap);
  GemmDesc d;
  d.A = raws_bf16;
  d.W = kv_w;
  d.M = n_img * 257;
  d.N = n_cross * 1536;
  d.K = Dv;
  d.lda = Dv;
  d.ldw = Dv;
  d.bias = kv_b;
  d.out_bf16 = kv;
  d.ldc = n_cross * 1536;
  d.out_col_block = head_major ? 64 : 0;
  kv_table_rows = (long long)n_img * 257;
  kv_head_major = head_major;
  return gemm_bf16_tcgen05(d, st);
}

int Model::qformer_layers(int B, int S, bool with_enc, int Lk, const int32_t* kv_idx0, const int32_t* kv_idx1,
                          const float* key_mask, int live_out, long long kv_rows, cudaStream_t st) {
  const int rows = B * S;
  SPRC_REQUIRE(rows <= qf_rows, "qformer: %d rows exceed workspace (%d)", rows, qf_rows);
  SPRC_REQUIRE(live_out == QF_OUT_ALL || S == 64, "qformer: row-restricted output needs S = 64");
  const int g = (S == 64) ? 32 : 0;  // row grouping for "first/last 32 rows of each 64-row sample"
  const int gs = (S == 64) ? 64 : 0;
  const int ldkv = n_cross * 1536;
  for (int l = 0; l < qf_layers; ++l) {
    const QfLayer& L = layers[l];
    // Rows whose output nobody reads are not computed in the LAST layer (their keys/values still are): the
    // fusion pass is consumed through its 32 query rows only (align_prompt.py:343 `fusion_output[:, :32]`,
    // rerank.py:440 `[:, :query_tokens.size(1)]`), the text pass through row 32 only (align_prompt.py:348).
    const int live = (l == qf_layers - 1) ? live_out : QF_OUT_ALL;
    // ---- self-attention over all S rows ----
    SPRC_TRY(linear(qhb, rows, 768, 768, L.qkv_w, 2304, L.qkv_b, ACT_NONE, nullptr, nullptr, qqkv, 2304, 0, 0, st));
    AttnDesc a;
    a.Q = qqkv;
    a.K = qqkv + 768;
    a.V = qqkv + 1536;
    a.O = qctx;
    a.B = B;
    a.H = 12;
    a.dh = 64;
    a.Lq = a.Lk = S;
    a.ldq = a.ldk = a.ldv = 2304;
    a.ldo = 768;
    a.q_batch_rows = a.kv_batch_rows = S;
    a.key_mask = key_mask;
    a.scale = 0.125f;
    SPRC_TRY(attention(a, st));
    // post-LN residual sublayers (Qformer.py:291-295): qh += dense(ctx) by TMA reduce-add, then LayerNorm in place
    if (live == QF_OUT_ALL) {
      SPRC_TRY(linear_ln(qctx, rows, 768, 768, L.so_w, L.so_b, L.so_g, L.so_beta, 1e-12f, qh, qhb, 0, 0, st));
    } else {
      // QF_OUT_QUERY_ROWS: rows [0,32) of each sample; QF_OUT_TEXT_CLS: row 32 of each sample
      const int gr = live == QF_OUT_QUERY_ROWS ? 32 : 1;
      const size_t o = live == QF_OUT_QUERY_ROWS ? 0 : 32;
      const int m = B * gr;
      SPRC_TRY(linear_ln(qctx + o * 768, m, 768, 768, L.so_w, L.so_b, L.so_g, L.so_beta, 1e-12f, qh + o * 768,
                         qhb + o * 768, gr, 64, st));
    }
    if (with_enc) {
      if (L.has_cross) {
        const int ci = l / 2;
        // query rows only (Qformer.py:436): Q projection, attention over the image tokens, output + LN
        SPRC_TRY(linear(qhb, B * 32, 768, 768, L.cq_w, 768, L.cq_b, ACT_NONE, nullptr, nullptr, qcq, 768, g, gs, st));
        AttnDesc c;
        c.Q = qcq;
        if (kv_idx0 && kv_head_major) kv_rows = kv_table_rows;   // rerank: image index tables over the whole table
        if (kv_rows > 0) {  // head-major blocks (cross_kv): block index = ci * 24 + {K: 0, V: 12} + head
          c.K = kv + (size_t)ci * 24 * kv_rows * 64;
          c.V = kv + ((size_t)ci * 24 + 12) * kv_rows * 64;
          c.kv_head_stride = kv_rows * 64;
          c.kv_rows_total = kv_rows;
        } else {
          c.K = kv + (size_t)ci * 1536;
          c.V = kv + (size_t)ci * 1536 + 768;
        }
        c.O = qctx;
        c.B = B;
        c.H = 12;
        c.dh = 64;
        c.Lq = 32;
        c.Lk = Lk;
        c.ldq = 768;
        c.ldk = c.ldv = kv_rows > 0 ? 64 : ldkv;
        c.ldo = 768;
        c.q_batch_rows = S;
        c.kv_batch_rows = 257;
        c.scale = 0.125f;
        c.kv_idx0 = kv_idx0;
        c.kv_idx1 = kv_idx1;
        c.Lk1 = 257;
        SPRC_TRY(attention(c, st));
        SPRC_TRY(linear_ln(qctx, B * 32, 768, 768, L.co_w, L.co_b, L.co_g, L.co_beta, 1e-12f, qh, qhb, g, gs, st));
      }
      // query rows -> *_query FFN; text rows -> text FFN (Qformer.py:455-468)
      SPRC_TRY(linear(qhb, B * 32, 768, 768, L.qi_w, 3072, L.qi_b, ACT_GELU, nullptr, nullptr, qffn, 3072, g, gs, st));
      SPRC_TRY(linear_ln(qffn, B * 32, 3072, 3072, L.qo_w, L.qo_b, L.qo_g, L.qo_beta, 1e-12f, qh, qhb, g, gs, st));
      if (S == 64 && live == QF_OUT_ALL) {
        const size_t o = 32;
        SPRC_TRY(linear(qhb + o * 768, B * 32, 768, 768, L.ti_w, 3072, L.ti_b, ACT_GELU, nullptr, nullptr,
                        qffn + o * 3072, 3072, g, gs, st));
        SPRC_TRY(linear_ln(qffn + o * 3072, B * 32, 3072, 3072, L.to_w, L.to_b, L.to_g, L.to_beta, 1e-12f,
                           qh + o * 768, qhb + o * 768, g, gs, st));
      }
    } else if (live == QF_OUT_TEXT_CLS) {
      const size_t o = 32;
      SPRC_TRY(linear(qhb + o * 768, B, 768, 768, L.ti_w, 3072, L.ti_b, ACT_GELU, nullptr, nullptr, qffn + o * 3072,
                      3072, 1, 64, st));
      SPRC_TRY(linear_ln(qffn + o * 3072, B, 3072, 3072, L.to_w, L.to_b, L.to_g, L.to_beta, 1e-12f, qh + o * 768,
                         qhb + o * 768, 1, 64, st));
    } else {
      // no encoder states: every row takes the text FFN (Qformer.py:469-475, the "baiyang change" at :434-435)
      SPRC_TRY(linear(qhb, rows, 768, 768, L.ti_w, 3072, L.ti_b, ACT_GELU, nullptr, nullptr, qffn, 3072, 0, 0, st));
      SPRC_TRY(linear_ln(qffn, rows, 3072, 3072, L.to_w, L.to_b, L.to_g, L.to_beta, 1e-12f, qh, qhb, 0, 0, st));
    }
  }
  return 0;
}

int Model::encode_gallery(const float* images, int B, float* feats_f32, bf16* feats_bf16, float* raws_f32,
                          bf16* raws_bf16, cudaStream_t st) {
  SPRC_REQUIRE(B > 0 && B <= max_images, "encode_gallery: B=%d outside (0, %d]", B, max_images);
  bf16* rb = raws_bf16 ? raws_bf16 : raws;
  SPRC_TRY(vit_forward(images, B, raws_f32, rb, st));
  if (!feats_f32 && !feats_bf16) return 0;
  SPRC_TRY(cross_kv(rb, B, true, st));
  // embeddings = LayerNorm(query_tokens)  (Qformer.py:110-112), broadcast over the batch
  SPRC_TRY(qformer_embed_rows(query_tokens, 0, nullptr, 1, word_emb, pos_emb, vocab, B, qt, st));
  SPRC_TRY(layernorm(qt, B * 32, 768, emb_g, emb_b, 1e-12f, 0, 0, qh, qhb, st));
  SPRC_TRY(qformer_layers(B, 32, true, 257, nullptr, nullptr, nullptr, QF_OUT_ALL, (long long)B * 257, st));
  SPRC_TRY(linear(qhb, B * 32, 768, 768, vproj_w, 256, vproj_b, ACT_NONE, nullptr, qproj, nullptr, 256, 0, 0, st));
  SPRC_TRY(l2norm_rows256(qproj, 256, B * 32, feats_f32, feats_bf16, st));
  return 0;
}

int Model::encode_query(const void* ref_raws, int ref_dtype, const int32_t* ref_rows, const int64_t* ids,
                        const int64_t* mask, int Bq, float* fusion_f32, bf16* fusion_bf16, cudaStream_t st) {
  SPRC_REQUIRE(Bq > 0 && Bq <= max_queries, "encode_query: Bq=%d outside (0, %d]", Bq, max_queries);
  SPRC_REQUIRE(ref_dtype == SPRC_F32 || ref_dtype == SPRC_BF16, "encode_query: ref dtype %d unsupported", ref_dtype);
  const size_t row_elems = (size_t)257 * Dv;
  const bf16* rb;
  if (ref_rows) {
    SPRC_TRY(gather_rows_bf16(ref_raws, ref_dtype, ref_rows, Bq, row_elems, raws, st));
    rb = raws;
  } else if (ref_dtype == SPRC_F32) {
    SPRC_TRY(convert_f32_to_bf16(static_cast<const float*>(ref_raws), raws, (size_t)Bq * row_elems, st));
    rb = raws;
  } else {
    rb = static_cast<const bf16*>(ref_raws);
  }
  SPRC_TRY(cross_kv(rb, Bq, true, st));
  SPRC_TRY(qformer_key_mask(mask, 1, Bq, qmask, st));
  // pass 1: fusion = Qformer(text, query_tokens, enc = reference embeds)   (align_prompt.py:332-339)
  SPRC_TRY(qformer_embed_rows(query_tokens, 0, ids, 1, word_emb, pos_emb, vocab, Bq, qt, st));
  SPRC_TRY(layernorm(qt, Bq * 64, 768, emb_g, emb_b, 1e-12f, 0, 0, qh, qhb, st));
  SPRC_TRY(qformer_layers(Bq, 64, true, 257, nullptr, nullptr, qmask, QF_OUT_QUERY_ROWS, (long long)Bq * 257, st));
  // pass 2: text = Qformer(text, query_embeds = fusion[:, :32])  with no encoder states (:341-346)
  SPRC_TRY(qformer_embed_rows(qh, 64, ids, 1, word_emb, pos_emb, vocab, Bq, qt, st));
  SPRC_TRY(layernorm(qt, Bq * 64, 768, emb_g, emb_b, 1e-12f, 0, 0, qh, qhb, st));
  SPRC_TRY(qformer_layers(Bq, 64, false, 0, nullptr, nullptr, qmask, QF_OUT_TEXT_CLS, 0, st));
  // fusion_feats = normalize(text_proj(h[:, 32]))   (:348-350): row 32 of every 64-row sample
  SPRC_TRY(linear(qhb + (size_t)32 * 768, Bq, 768, 768, tproj_w, 256, tproj_b, ACT_NONE, nullptr, qproj, nullptr, 256,
                  1, 64, st));
  SPRC_TRY(l2norm_rows256(qproj, (size_t)64 * 256, Bq, fusion_f32, fusion_bf16, st));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// composed-query fusion over the ragged row layout
// ------------------------------------------------------------------------------------------------
bool ragged_query_enabled();
static bool ragged_enabled() { return ragged_query_enabled(); }
bool ragged_query_enabled() {
  static const bool on = [] {
    const char* e = getenv("SPRC_RAGGED");   // SPRC_RAGGED=0: always run the padded 64-rows-per-sample passes
    return !(e && e[0] == '0');
  }();
  return on;
}

// One Q-Former pass over rows [0, 32 B) (query rows) + [32 B, 32 B + T8) (live text rows, attention_qfr.cu).
// with_enc: fusion pass (cross-attention + query FFN on the query rows, text FFN on the text rows; the last layer
// computes the query rows only, align_prompt.py:343).  !with_enc: text pass (text FFN on every row, Qformer.py:434-435,
// 469-475; the last layer computes the [CLS] rows only, gathered into dense [B, 768] buffers: qt = fp32, qcq = 16-bit,
// align_prompt.py:348).
int Model::qformer_layers_ragged(int B, int T8, bool with_enc, int Lk, const int32_t* kv_idx0, const int32_t* kv_idx1,
                                 cudaStream_t st) {
  const int qrows = 32 * B, rows_all = qrows + T8;
  SPRC_REQUIRE(rows_all <= qf_rows, "qformer: %d rows exceed workspace (%d)", rows_all, qf_rows);
  const size_t to = (size_t)qrows;   // first text row
  float* x_cls = qt;
  bf16* ctx_cls = qcq;
  bf16* x_cls_b = qcq + (size_t)B * 768;
  for (int l = 0; l < qf_layers; ++l) {
    const QfLayer& L = layers[l];
    const bool last = l == qf_layers - 1;
    SPRC_TRY(linear(qhb, rows_all, 768, 768, L.qkv_w, 2304, L.qkv_b, ACT_NONE, nullptr, nullptr, qqkv, 2304, 0, 0, st));
    SPRC_TRY(attention_qf_ragged(qqkv, 2304, qctx, 768, B, rows_all, static_cast<const int4*>(m_pairs), 0.125f, st));
    if (!last) {
      SPRC_TRY(linear_ln(qctx, rows_all, 768, 768, L.so_w, L.so_b, L.so_g, L.so_beta, 1e-12f, qh, qhb, 0, 0, st));
    } else if (with_enc) {
      SPRC_TRY(linear_ln(qctx, qrows, 768, 768, L.so_w, L.so_b, L.so_g, L.so_beta, 1e-12f, qh, qhb, 0, 0, st));
    } else {
      SPRC_TRY(gather_rows768(qh, qctx, m_cls, qrows, B, x_cls, ctx_cls, st));
      SPRC_TRY(linear_ln(ctx_cls, B, 768, 768, L.so_w, L.so_b, L.so_g, L.so_beta, 1e-12f, x_cls, x_cls_b, 0, 0, st));
    }
    if (with_enc) {
      if (L.has_cross) {
        const int ci = l / 2;
        const long long kv_rows = kv_idx0 ? kv_table_rows : (long long)B * 257;
        SPRC_TRY(linear(qhb, qrows, 768, 768, L.cq_w, 768, L.cq_b, ACT_NONE, nullptr, nullptr, qcq, 768, 0, 0, st));
        AttnDesc c;
        c.Q = qcq;
        c.kv_rows_total = kv_rows;
        if (kv_idx0 && !kv_head_major) {   // rerank over plain K/V rows, keys = cat(reference image, candidate image)
          c.K = kv + (size_t)ci * 1536;
          c.V = kv + (size_t)ci * 1536 + 768;
          c.ldk = c.ldv = n_cross * 1536;
          c.kv_idx0 = kv_idx0;
          c.kv_idx1 = kv_idx1;
        } else {         // head-major blocks (cross_kv); rerank: image index tables over them
          c.kv_idx0 = kv_idx0;
          c.kv_idx1 = kv_idx1;
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
        SPRC_TRY(linear_ln(qctx, qrows, 768, 768, L.co_w, L.co_b, L.co_g, L.co_beta, 1e-12f, qh, qhb, 0, 0, st));
      }
      if (!last && T8 > 0 && qrows % 256 == 0 && dual_ffn_enabled()) {
        // both FFNs as ONE grid per GEMM: query rows read the *_query weights, text rows the text weights
        // (GemmDesc::W2) - a 27.9k-row launch instead of an 18.9k-row one plus a half-empty 9k-row one
        GemmDesc d;
        d.A = qhb;
        d.M = rows_all;
        d.m_split = qrows;
        d.K = d.lda = d.ldw = 768;
        d.N = d.ldc = 3072;
        d.W = L.qi_w, d.bias = L.qi_b;
        d.W2 = L.ti_w, d.bias2 = L.ti_b;
        d.act = ACT_GELU;
        d.out_bf16 = qffn;
        SPRC_TRY(gemm_bf16_tcgen05(d, st));
        GemmDesc e;
        e.A = qffn;
        e.M = rows_all;
        e.m_split = qrows;
        e.K = e.lda = e.ldw = 3072;
        e.N = e.ldc = 768;
        e.W = L.qo_w, e.bias = L.qo_b;
        e.W2 = L.to_w, e.bias2 = L.to_b;
        e.residual = e.out_f32 = qh;
        SPRC_TRY(gemm_bf16_tcgen05(e, st));
        SPRC_TRY(layernorm(qh, qrows, 768, L.qo_g, L.qo_beta, 1e-12f, 0, 0, qh, qhb, st));
        SPRC_TRY(layernorm(qh + to * 768, T8, 768, L.to_g, L.to_beta, 1e-12f, 0, 0, qh + to * 768, qhb + to * 768, st));
      } else {
        SPRC_TRY(linear(qhb, qrows, 768, 768, L.qi_w, 3072, L.qi_b, ACT_GELU, nullptr, nullptr, qffn, 3072, 0, 0, st));
        SPRC_TRY(linear_ln(qffn, qrows, 3072, 3072, L.qo_w, L.qo_b, L.qo_g, L.qo_beta, 1e-12f, qh, qhb, 0, 0, st));
        if (!last) {
          SPRC_TRY(linear(qhb + to * 768, T8, 768, 768, L.ti_w, 3072, L.ti_b, ACT_GELU, nullptr, nullptr,
                          qffn + to * 3072, 3072, 0, 0, st));
          SPRC_TRY(linear_ln(qffn + to * 3072, T8, 3072, 3072, L.to_w, L.to_b, L.to_g, L.to_beta, 1e-12f,
                             qh + to * 768, qhb + to * 768, 0, 0, st));
        }
      }
    } else if (!last) {
      SPRC_TRY(linear(qhb, rows_all, 768, 768, L.ti_w, 3072, L.ti_b, ACT_GELU, nullptr, nullptr, qffn, 3072, 0, 0, st));
      SPRC_TRY(linear_ln(qffn, rows_all, 3072, 3072, L.to_w, L.to_b, L.to_g, L.to_beta, 1e-12f, qh, qhb, 0, 0, st));
    } else {
      SPRC_TRY(linear(x_cls_b, B, 768, 768, L.ti_w, 3072, L.ti_b, ACT_GELU, nullptr, nullptr, qffn, 3072, 0, 0, st));
      SPRC_TRY(linear_ln(qffn, B, 3072, 3072, L.to_w, L.to_b, L.to_g, L.to_beta, 1e-12f, x_cls, x_cls_b, 0, 0, st));
    }
  }
  return 0;
}

// Row tables of the ragged layout (attention_qfr.cu), uploaded to d_meta: toff[B] | len[B] | cls[B] | row_sample[T8] |
// pairs int4 [ceil(B / 2)].  Text rows: the two samples of a pair are adjacent, each pair's slot is rounded up to 8.
int Model::build_ragged_meta(const int32_t* lens_host, int repeat, int B, int* T8_out, cudaStream_t st) {
  const int32_t* text_len_host = lens_host;
  if (repeat < 1) repeat = 1;
  // ---- row tables: toff[B] | len[B] | cls[B] | row_sample[T8] | pairs int4 [ceil(B / 2)] ----
  // text rows: the two samples of a pair are adjacent, each pair's slot is rounded up to 8 rows
  const int P = (B + 1) / 2;
  h_meta.assign((size_t)B * 3, 0);
  int T8 = 0;
  std::vector<int32_t> rsmp;
  rsmp.reserve((size_t)B * 20);
  for (int g = 0; g < P; ++g) {
    int used = 0;
    for (int b = 2 * g; b < 2 * g + 2 && b < B; ++b) {
      int L = text_len_host[b / repeat];
      SPRC_REQUIRE(L >= 0 && L <= 32, "encode_query: caption %d has %d live tokens", b, L);
      if (L < 1) L = 1;   // an all-masked caption still owns its [CLS] row (the padded path reads row 32 regardless)
      h_meta[b] = T8 + used;
      h_meta[B + b] = L;
      h_meta[2 * B + b] = T8 + used;   // [CLS] = first text row of the sample
      for (int i = 0; i < L; ++i) rsmp.push_back(b);
      used += L;
    }
    const int slot8 = (used + 7) & ~7;
    const int last = (2 * g + 1 < B) ? 2 * g + 1 : 2 * g;
    for (int i = used; i < slot8; ++i) rsmp.push_back(last);   // slack rows (t >= L: written as zeros)
    T8 += slot8;
  }
  const size_t off_slot = (size_t)B * 3;
  size_t off_pairs = off_slot + rsmp.size();
  off_pairs = (off_pairs + 3) & ~size_t(3);   // int4 alignment
  h_meta.resize(off_pairs + (size_t)P * 4, 0);
  for (size_t i = 0; i < rsmp.size(); ++i) h_meta[off_slot + i] = rsmp[i];
  for (int g = 0; g < P; ++g) {
    const int b0 = 2 * g, b1 = 2 * g + 1;
    const int L0 = h_meta[B + b0], L1 = b1 < B ? h_meta[B + b1] : 0;
    h_meta[off_pairs + 4 * g] = h_meta[b0];
    h_meta[off_pairs + 4 * g + 1] = L0;
    h_meta[off_pairs + 4 * g + 2] = h_meta[b0] + L0;   // sample 1 follows sample 0 directly
    h_meta[off_pairs + 4 * g + 3] = L1;
  }
  SPRC_REQUIRE(h_meta.size() <= meta_cap, "ragged row tables (%zu ints) exceed their buffer (%zu)", h_meta.size(),
               meta_cap);
  // pinned ring slot -> d_meta, fully asynchronous: the host may prepare and enqueue up to kMetaRing batches ahead
  const int slot = static_cast<int>(meta_seq++ % kMetaRing);
  if (!h_meta_pin[0]) {
    // the whole ring at the first call: page-locked allocation synchronises the device, and a slot allocated lazily
    // inside a query loop drains the queue in front of that batch (seen as one 60 ms bubble per new slot in bench.py)
    for (int i = 0; i < kMetaRing; ++i) {
      SPRC_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h_meta_pin[i]), meta_cap * sizeof(int32_t)));
      SPRC_CUDA(cudaEventCreateWithFlags(&meta_ev[i], cudaEventDisableTiming));
      SPRC_CUDA(cudaEventRecord(meta_ev[i], st));
    }
  }
  SPRC_CUDA(cudaEventSynchronize(meta_ev[slot]));   // the copy that last used this slot has executed
  memcpy(h_meta_pin[slot], h_meta.data(), h_meta.size() * sizeof(int32_t));
  SPRC_CUDA(cudaMemcpyAsync(d_meta, h_meta_pin[slot], h_meta.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  SPRC_CUDA(cudaEventRecord(meta_ev[slot], st));
  m_toff = d_meta;
  m_len = d_meta + B;
  m_cls = d_meta + 2 * B;
  m_slot = d_meta + off_slot;
  m_pairs = d_meta + off_pairs;

  *T8_out = T8;
  return 0;
}

int Model::encode_query_ragged(const void* ref_raws, int ref_dtype, const int32_t* ref_rows, const int64_t* ids,
                               const int32_t* text_len_host, int Bq, float* fusion_f32, bf16* fusion_bf16,
                               cudaStream_t st) {
  SPRC_REQUIRE(Bq > 0 && Bq <= max_queries, "encode_query: Bq=%d outside (0, %d]", Bq, max_queries);
  SPRC_REQUIRE(ref_dtype == SPRC_F32 || ref_dtype == SPRC_BF16, "encode_query: ref dtype %d unsupported", ref_dtype);
  const int B = Bq;
  int T8 = 0;
  SPRC_TRY(build_ragged_meta(text_len_host, 1, B, &T8, st));
  const size_t row_elems = (size_t)257 * Dv;
  const bf16* rb;
  if (ref_rows) {
    SPRC_TRY(gather_rows_bf16(ref_raws, ref_dtype, ref_rows, Bq, row_elems, raws, st));
    rb = raws;
  } else if (ref_dtype == SPRC_F32) {
    SPRC_TRY(convert_f32_to_bf16(static_cast<const float*>(ref_raws), raws, (size_t)Bq * row_elems, st));
    rb = raws;
  } else {
    rb = static_cast<const bf16*>(ref_raws);
  }
  SPRC_TRY(cross_kv(rb, Bq, true, st));
  const int rows_all = 32 * B + T8;
  // pass 1: fusion = Qformer(text, query_tokens, enc = reference embeds)   (align_prompt.py:332-339)
  SPRC_TRY(qformer_embed_ragged(query_tokens, 0, ids, 1, m_slot, m_toff, m_len, word_emb, pos_emb, vocab, B, rows_all,
                                qt, st));
  SPRC_TRY(layernorm(qt, rows_all, 768, emb_g, emb_b, 1e-12f, 0, 0, qh, qhb, st));
  SPRC_TRY(qformer_layers_ragged(B, T8, true, 257, nullptr, nullptr, st));
  // pass 2: text = Qformer(text, query_embeds = fusion[:, :32]) with no encoder states (:341-346); the query rows of
  // pass 1's output are rows [0, 32 B) of qh
  SPRC_TRY(qformer_embed_ragged(qh, 1, ids, 1, m_slot, m_toff, m_len, word_emb, pos_emb, vocab, B, rows_all, qt, st));
  SPRC_TRY(layernorm(qt, rows_all, 768, emb_g, emb_b, 1e-12f, 0, 0, qh, qhb, st));
  SPRC_TRY(qformer_layers_ragged(B, T8, false, 0, nullptr, nullptr, st));
  // fusion_feats = normalize(text_proj(h[:, 32]))   (:348-350): the gathered [CLS] rows (dense [B, 768] in qcq + B*768)
  SPRC_TRY(linear(qcq + (size_t)B * 768, Bq, 768, 768, tproj_w, 256, tproj_b, ACT_NONE, nullptr, qproj, nullptr, 256, 0,
                  0, st));
  SPRC_TRY(l2norm_rows256(qproj, (size_t)256, Bq, fusion_f32, fusion_bf16, st));
  return 0;
}

static __global__ void rerank_pair_images_kernel(int32_t* ref_img, int32_t* cand_img, int pairs, int T, int r) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < pairs) {
    ref_img[i] = i / T;    // reference image of the pair: KV rows [0, 257)
    cand_img[i] = r + i;   // candidate image: KV rows [257, 514)
  }
}

int Model::rerank(const bf16* raws_table, const int32_t* ref_rows, const int32_t* cand_rows, const int64_t* ids,
                  const int64_t* mask, const int32_t* text_len_host, int R, int T, float* p, cudaStream_t st) {
  SPRC_REQUIRE(max_pairs > 0, "rerank: handle was created with max_pairs = 0");
  SPRC_REQUIRE(R > 0 && T > 0 && T <= max_pairs, "rerank: R=%d T=%d (max_pairs %d)", R, T, max_pairs);
  const size_t row_elems = (size_t)257 * Dv;
  int rc = max_pairs / T;  // queries per chunk
  if (rc < 1) rc = 1;
  for (int r0 = 0; r0 < R; r0 += rc) {
    const int r = (R - r0) < rc ? (R - r0) : rc;
    const int pairs = r * T;
    const int n_img = r + pairs;
    SPRC_REQUIRE(n_img <= enc_cap, "rerank: %d images exceed workspace (%d)", n_img, enc_cap);
    // hoist the image-only K/V projections out of the pair loop (SURVEY.md §7 "Rerank cost"):
    // raws buffer = [refs of this chunk ; candidates of this chunk]
    SPRC_TRY(gather_rows_bf16(raws_table, SPRC_BF16, ref_rows + r0, r, row_elems, raws, st));
    SPRC_TRY(gather_rows_bf16(raws_table, SPRC_BF16, cand_rows + (size_t)r0 * T, pairs, row_elems,
                              raws + (size_t)r * row_elems, st));
    SPRC_TRY(cross_kv(raws, n_img, true, st));   // head-major: every (image, head) K / V block is contiguous
    // pair i reads the K/V rows of image i / T (its reference) and of image r + i (its candidate): written on the
    // device, so the call enqueues only (no host staging, no synchronisation)
    rerank_pair_images_kernel<<<(pairs + 255) / 256, 256, 0, st>>>(d_rows, d_rows2, pairs, T, r);
    count_launch();
    SPRC_CUDA(cudaGetLastError());
    if (text_len_host && ragged_enabled()) {
      // ragged rows: every pair owns its 32 query rows + the live tokens of its caption (attention_qfr.cu)
      int T8 = 0;
      SPRC_TRY(build_ragged_meta(text_len_host + r0, T, pairs, &T8, st));
      const int rows_all = 32 * pairs + T8;
      SPRC_TRY(qformer_embed_ragged(query_tokens, 0, ids + (size_t)r0 * 32, T, m_slot, m_toff, m_len, word_emb, pos_emb,
                                    vocab, pairs, rows_all, qt, st));
      SPRC_TRY(layernorm(qt, rows_all, 768, emb_g, emb_b, 1e-12f, 0, 0, qh, qhb, st));
      SPRC_TRY(qformer_layers_ragged(pairs, T8, true, 514, d_rows, d_rows2, st));
      SPRC_TRY(itm_head_prob(qh, 32, pairs, itm_w, itm_b, p + (size_t)r0 * T, st));
      continue;
    }
    SPRC_REQUIRE(mask != nullptr, "rerank: attention mask or caption lengths needed");
    SPRC_TRY(qformer_key_mask(mask + (size_t)r0 * 32, T, pairs, qmask, st));
    SPRC_TRY(qformer_embed_rows(query_tokens, 0, ids + (size_t)r0 * 32, T, word_emb, pos_emb, vocab, pairs, qt, st));
    SPRC_TRY(layernorm(qt, pairs * 64, 768, emb_g, emb_b, 1e-12f, 0, 0, qh, qhb, st));
    SPRC_TRY(qformer_layers(pairs, 64, true, 514, d_rows, d_rows2, qmask, QF_OUT_QUERY_ROWS, 0, st));
    SPRC_TRY(itm_head_prob(qh, 64, pairs, itm_w, itm_b, p + (size_t)r0 * T, st));
  }
  return 0;
}

}  // namespace sprc
