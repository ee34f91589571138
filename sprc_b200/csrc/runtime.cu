// Error reporting and device queries for libsprc_b200 (see include/sprc_b200.h conventions).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.h"
#include "ops.h"

#include <string>
#include <vector>

namespace sprc {

static thread_local char g_err[1024] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

const char* last_error() { return g_err; }

static int g_act_fp16 = 0;
int act_fp16() { return g_act_fp16; }
void set_act_fp16(int on) { g_act_fp16 = on ? 1 : 0; }

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SPRC_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

int device_sm_count() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms[dev] = n;
  }
  return sms[dev];
}


// ------------------------------------------------------------------------------------------------
// optional per-launch timing (CUDA events on the launching stream), by kernel category
// ------------------------------------------------------------------------------------------------
struct ProfRec {
  cudaEvent_t a, b;
  int cat;
  double flops, bytes;
  char tag[56];
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_free_events;
static cudaEvent_t g_pending = nullptr;

static cudaEvent_t get_event() {
  if (!g_free_events.empty()) {
    cudaEvent_t e = g_free_events.back();
    g_free_events.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

bool prof_enabled() { return g_prof_on; }
int next_sweep_reverse() {
  static thread_local unsigned sweep = 0;
  static const bool enabled = [] {
    const char* e = getenv("SPRC_SWEEP");  // SPRC_SWEEP=0: always ascending (A/B measurements)
    return !(e && e[0] == '0');
  }();
  return enabled ? static_cast<int>(sweep++ & 1u) : 0;
}

void prof_begin(cudaStream_t st) {
  if (!g_prof_on) return;
  g_pending = get_event();
  cudaEventRecord(g_pending, st);
}
void prof_end(int cat, double flops, double bytes, cudaStream_t st, const char* tag) {
  if (!g_prof_on || !g_pending) return;
  ProfRec r;
  r.a = g_pending;
  r.b = get_event();
  r.cat = cat;
  r.flops = flops;
  r.bytes = bytes;
  snprintf(r.tag, sizeof(r.tag), "%s", tag ? tag : "");
  cudaEventRecord(r.b, st);
  g_recs.push_back(r);
  g_pending = nullptr;
}
void prof_set(bool on) {
  g_prof_on = on;
  for (auto& r : g_recs) {
    g_free_events.push_back(r.a);
    g_free_events.push_back(r.b);
  }
  g_recs.clear();
}
// out[cat*4 + {0,1,2,3}] = {total ms, flops, bytes, launches}
int prof_read(double* out, int ncat) {
  for (int i = 0; i < ncat * 4; ++i) out[i] = 0.0;
  for (auto& r : g_recs) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) return set_error(-5, "profile: event sync failed");
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    if (r.cat >= 0 && r.cat < ncat) {
      out[r.cat * 4 + 0] += ms;
      out[r.cat * 4 + 1] += r.flops;
      out[r.cat * 4 + 2] += r.bytes;
      out[r.cat * 4 + 3] += 1.0;
    }
  }
  return 0;
}


// CSV of per-(category, tag) aggregates: cat,tag,launches,total_ms,flops,bytes
int prof_dump(const char* path) {
  FILE* f = fopen(path, "w");
  if (!f) return set_error(-2, "profile: cannot open %s", path);
  struct Agg {
    int cat;
    std::string tag;
    double n, ms, flops, bytes;
  };
  std::vector<Agg> aggs;
  for (auto& r : g_recs) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    Agg* a = nullptr;
    for (auto& x : aggs)
      if (x.cat == r.cat && x.tag == r.tag) a = &x;
    if (!a) {
      aggs.push_back({r.cat, r.tag, 0, 0, 0, 0});
      a = &aggs.back();
    }
    a->n += 1;
    a->ms += ms;
    a->flops += r.flops;
    a->bytes += r.bytes;
  }
  fprintf(f, "cat,tag,launches,total_ms,flops,bytes\n");
  for (auto& a : aggs) fprintf(f, "%d,%s,%.0f,%.6f,%.6e,%.6e\n", a.cat, a.tag.c_str(), a.n, a.ms, a.flops, a.bytes);
  fclose(f);
  return 0;
}

}  // namespace sprc
