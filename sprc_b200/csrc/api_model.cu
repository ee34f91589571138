// C-ABI: model-level entry points (handle, weights, gallery/query encoders, scan, rerank).
#include "../../include/sprc_b200.h"
#include "common.h"
#include "ops.h"

using namespace sprc;

#define SPRC_TODO(name) return set_error(-38, name ": not implemented yet")

extern "C" {

int sprc_create(const sprc_config*, sprc_handle**) { SPRC_TODO("sprc_create"); }
void sprc_destroy(sprc_handle*) {}
int sprc_load_weights(sprc_handle*, const sprc_tensor_desc*, int, int*) { SPRC_TODO("sprc_load_weights"); }
const char* sprc_missing_weight(sprc_handle*, int) { return nullptr; }
int sprc_encode_gallery(sprc_handle*, const float*, int, float*, void*, float*, void*, void*) {
  SPRC_TODO("sprc_encode_gallery");
}
int sprc_encode_query(sprc_handle*, const void*, int, const int32_t*, const int64_t*, const int64_t*, int, float*,
                      void*, void*) {
  SPRC_TODO("sprc_encode_query");
}
int sprc_sim_topk(sprc_handle*, const void*, int, const void*, int64_t, int64_t, int, float*, int32_t*, float*,
                  void*) {
  SPRC_TODO("sprc_sim_topk");
}
int sprc_topk_merge(sprc_handle*, const float*, const int32_t*, int, int, int, float*, int32_t*, void*) {
  SPRC_TODO("sprc_topk_merge");
}
int sprc_gather_scores(sprc_handle*, const void*, int, const void*, int64_t, const int32_t*, int, float*, void*) {
  SPRC_TODO("sprc_gather_scores");
}
int sprc_rerank(sprc_handle*, const void*, const int32_t*, const int32_t*, const int64_t*, const int64_t*, int, int,
                float*, void*) {
  SPRC_TODO("sprc_rerank");
}
int sprc_query_topk_host(sprc_handle*, const void*, const void*, int64_t, const int32_t*, const int64_t*,
                         const int64_t*, int, int, float*, int32_t*, void*) {
  SPRC_TODO("sprc_query_topk_host");
}

}  // extern "C"
