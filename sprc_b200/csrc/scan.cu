// Similarity scan + top-k:
//   sim[q, n] = max_{t<32} < query[q, :], gallery[n, t, :] >        blip2_qformer_cir_align_prompt.py:353-358
//   ranking   = argsort(1 - sim)[:, :k]  (ties: lower gallery row)   validate_blip.py:44-46,253-255
//
// The reference expands the gallery Bq times and sorts all N scores per query.  Here one persistent
// kernel streams the bf16 gallery [N*32, 256] from HBM exactly once per 128-query tile:
//   warp 0      TMA producer: 64 gallery tokens (2 images) x 256 dims per ring slot, 4 slots
//   warp 1      tcgen05.mma issuer: D[128 queries, 128 tokens] = Qtile[128,256] (TMEM) * G[128,256]^T (slot pair), fp32 in TMEM
//               (queries on TMEM lanes so that the max over an image's 32 tokens is a per-thread max
//               over 32 accumulator columns - no shuffles)
//   warps 2..5  epilogue: one thread per query; keeps that query's running top-k as a binary min-heap in
//               shared memory (slot-major layout -> bank-conflict free for any mix of heap positions)
// CTAs = query tiles x gallery splits; per-(split, query) heaps go to a scratch buffer and a second
// small kernel merges them (the same kernel merges per-GPU candidates after the NCCL all-gather).
// Bytes per gallery pass: N*32*256*2 (819.2 MB at N = 50k); the scan is HBM-bound for Q <= 128.
#include <float.h>

#include "ops.h"
#include "ptx.cuh"

namespace sprc {

int make_tmap_bf16(CUtensorMap* tm, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
                   uint64_t stride2, uint32_t b0, uint32_t b1, uint32_t b2, int rank);

static constexpr int SQ = 128;        // queries per CTA (UMMA M)
static constexpr int ST = 64;         // gallery tokens per ring slot = 2 images (one TMA transaction group)
static constexpr int PT = 2 * ST;     // gallery tokens per MMA (UMMA N) = one PAIR of adjacent ring slots = 4 images
static constexpr int SSLOTS = 4;      // ring slots (32 KB each): one pair being consumed, one being filled
static constexpr int AS = 3;          // accumulator stages in TMEM (128 fp32 columns each)
static constexpr int KCAP = 64;       // heap capacity (fused path handles k <= 64)
static constexpr int PEND = 8;        // per-query pending candidates between heap flushes
static constexpr int SEG = 4096;      // segment width of the large-k path
static constexpr int SLOT_BYTES = ST * 256 * 2;     // 32 KB
static constexpr int SLAB_BYTES = ST * 128;         // one K block (64 dims) of one slot: 64 rows x 128 B
static constexpr int KB_STRIDE = SSLOTS * SLAB_BYTES;  // K-block regions hold the slabs of all slots back to back
static constexpr int TM_COL_A = 0;    // query tile: 128 lanes x 128 columns of packed 16-bit pairs (K = 256)
static constexpr int TM_COL_D = 128;  // AS x 128 fp32 accumulator columns

typedef unsigned long long u64;

__device__ __forceinline__ u64 make_key(float score, uint32_t idx) {
  uint32_t f = __float_as_uint(score);
  f = (f & 0x80000000u) ? ~f : (f | 0x80000000u);  // monotone float -> uint
  return (static_cast<u64>(f) << 32) | static_cast<u64>(0xFFFFFFFFu - idx);  // larger key = better
}
__device__ __forceinline__ void decode_key(u64 key, float& score, int32_t& idx) {
  const uint32_t hi = static_cast<uint32_t>(key >> 32);
  if (hi == 0) {  // never filled
    score = -INFINITY;
    idx = -1;
    return;
  }
  const uint32_t f = (hi & 0x80000000u) ? (hi & 0x7FFFFFFFu) : ~hi;
  score = __uint_as_float(f);
  idx = static_cast<int32_t>(0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFu));
}

struct ScanParams {
  const uint4* queries;      // [Q, 256] 16-bit, rows 512 B
  int Q;
  long long N;
  long long row_offset;
  int k;
  int qtiles, splits;
  long long imgs_per_split;  // multiple of 4
  int fp16;                  // operand format of queries / gallery
  float* out_full;           // [Q, N] or null
  u64* cand;                 // [splits][k][qtiles*128] or null
};

// Design notes (measured, profiles/r01b_* and the exp1/exp2 bench logs):
//  * The query tile is the A operand of every MMA of the CTA, so it lives in TENSOR MEMORY (tcgen05.mma with A in
//    TMEM) and shared memory carries only the gallery stream.
//  * The 16 K-step MMAs of a score tile accumulate into the same TMEM columns, i.e. they form a dependent chain: with
//    N = 64 every MMA took ~73 cycles (twice its 32-cycle issue floor) and the tensor pipe, not HBM or L2, paced the
//    scan at 44 % activity for 1 and for 2 query tiles per CTA alike.  The ring therefore stores the K-block slabs of
//    ADJACENT slots back to back (slot s, K block kb at kb * KB_STRIDE + s * SLAB_BYTES), so that two 64-token slots
//    form one contiguous 128-row K-major operand and each MMA is 128 x 128 x 16 (64-cycle floor): half the
//    instructions per gallery byte while TMA transactions stay 32 KB.
//  * A candidate that beats the heap minimum is only QUEUED; the warp drains the queues together when one of them is
//    nearly full.  Inserting at once made the whole warp walk the sift-down loop whenever any of its 32 queries had
//    a hit (almost every image while a split's threshold is still loose).
__global__ void __launch_bounds__(192, 1)
scan_topk_kernel(const __grid_constant__ CUtensorMap tmG, const ScanParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;
  u64* heap = reinterpret_cast<u64*>(smem + SSLOTS * SLOT_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SSLOTS * SLOT_BYTES + (p.k + PEND) * SQ * 8);
  uint64_t* full_bar = bars;                   // [SSLOTS]
  uint64_t* empty_bar = bars + SSLOTS;         // [SSLOTS]
  uint64_t* tfull_bar = bars + 2 * SSLOTS;     // [AS]
  uint64_t* tempty_bar = tfull_bar + AS;       // [AS]
  uint64_t* q_bar = tempty_bar + AS;           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % p.qtiles;
  const int sp = blockIdx.x / p.qtiles;
  const long long n_begin = static_cast<long long>(sp) * p.imgs_per_split;
  long long n_end = n_begin + p.imgs_per_split;
  if (n_end > p.N) n_end = p.N;
  const int pairs_total = n_end > n_begin ? static_cast<int>((n_end - n_begin + 3) / 4) : 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmG);
    for (int s = 0; s < SSLOTS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < AS; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    mbar_init(q_bar, 4);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int slot = 0;
      uint32_t phase = 0;
      // rows past the end of the gallery (odd tails) are zero-filled by the TMA and never scored
      for (int it = 0; it < 2 * pairs_total; ++it) {
        mbar_wait(&empty_bar[slot], phase ^ 1);
        mbar_expect_tx(&full_bar[slot], SLOT_BYTES);
        const long long row0 = (n_begin + 2LL * it) * 32;
        for (int kb = 0; kb < 4; ++kb)
          tma_load_2d(&tmG, &full_bar[slot], sB + kb * KB_STRIDE + slot * SLAB_BYTES, kb * 64,
                      static_cast<int>(row0), p.qtiles > 1 ? kEvictNormal : kEvictFirst);
        if (++slot == SSLOTS) {
          slot = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_16(SQ, PT, p.fp16);
    mbar_wait(q_bar, 0);   // the epilogue warps have written the query tile into TMEM
    tc_fence_after();
    int slot = 0, as = 0;
    uint32_t phase = 0, aphase = 0;
    for (int pr = 0; pr < pairs_total; ++pr) {
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      mbar_wait(&full_bar[slot], phase);
      mbar_wait(&full_bar[slot + 1], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(TM_COL_D + as * PT);
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
          // K step ks: dims [16 ks, 16 ks + 16) = TMEM columns [8 ks, 8 ks + 8) of the query tile; gallery K block
          // ks / 4 (rows of both slots contiguous), +32 B per step inside the 128 B swizzle row
          const uint64_t db =
              umma_desc_k_sw128(smem_u32(sB + (ks >> 2) * KB_STRIDE + slot * SLAB_BYTES)) + 2 * (ks & 3);
          umma_bf16_ts(d_tmem, tmem_base + TM_COL_A + ks * 8, db, idesc, ks != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[slot]);
        umma_commit(&empty_bar[slot + 1]);
        umma_commit(&tfull_bar[as]);
      }
      __syncwarp();
      slot += 2;
      if (slot == SSLOTS) {
        slot = 0;
        phase ^= 1;
      }
      if (++as == AS) {
        as = 0;
        aphase ^= 1;
      }
    }
  } else {
    // ===================== epilogue: one thread per query =====================
    const int qw = warp & 3;               // TMEM lane quarter this warp may access
    const int ql = qw * 32 + lane;         // query row inside the tile = TMEM lane
    const int q = qt * SQ + ql;
    const bool q_ok = q < p.Q;
    const int k = p.k;
    const uint32_t lane_addr = static_cast<uint32_t>(qw * 32) << 16;
    {
      // query row -> TMEM lane ql, 4 x 32 columns (two 16-bit elements per column, memory order)
      const uint4* src = p.queries + static_cast<size_t>(q_ok ? q : 0) * 32;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (q_ok) v = __ldg(src + c * 8 + j);
          r[4 * j] = v.x, r[4 * j + 1] = v.y, r[4 * j + 2] = v.z, r[4 * j + 3] = v.w;
        }
        tmem_st32(tmem_base + lane_addr + TM_COL_A + c * 32, r);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_bar);
    }
    constexpr uint32_t HS = SQ * 8;
    // slot j of this query's heap at byte offset j * HS (slot-major: conflict-free for any mix of heap positions)
    const uint32_t myheap = smem_u32(heap) + ql * 8;
    const uint32_t mypend = myheap + k * HS;    // PEND pending candidates
    for (int j = 0; j < k; ++j) sts64(myheap + j * HS, static_cast<u64>(j));  // distinct sub-minimal keys, valid min-heap
    u64 root = 0;
    int pcnt = 0;
    auto flush = [&]() {
      for (int j = 0; j < pcnt; ++j) {
        const u64 key = lds64(mypend + j * HS);
        if (key > root) {
          int i = 0;
          while (true) {
            const int l = 2 * i + 1;
            if (l >= k) break;
            const u64 kl = lds64(myheap + l * HS);
            const u64 kr = (l + 1 < k) ? lds64(myheap + (l + 1) * HS) : ~0ull;
            const int c = kr < kl ? l + 1 : l;
            const u64 kc = kr < kl ? kr : kl;
            if (kc >= key) break;
            sts64(myheap + i * HS, kc);
            i = c;
          }
          sts64(myheap + i * HS, key);
          root = lds64(myheap);
        }
      }
      pcnt = 0;
    };
    int as = 0;
    uint32_t aphase = 0;
    for (int pr = 0; pr < pairs_total; ++pr) {
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + lane_addr + static_cast<uint32_t>(TM_COL_D + as * PT);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r0[32], r1[32];
        tmem_ld32(taddr + half * 64, r0);
        tmem_ld32(taddr + half * 64 + 32, r1);
        tmem_ld_wait();
        if (half == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[as]);  // accumulator is in registers: release TMEM early
        }
        float s0 = __uint_as_float(r0[0]), s1 = __uint_as_float(r1[0]);
#pragma unroll
        for (int j = 1; j < 32; ++j) {
          s0 = fmaxf(s0, __uint_as_float(r0[j]));
          s1 = fmaxf(s1, __uint_as_float(r1[j]));
        }
        const long long n0 = n_begin + 4LL * pr + 2 * half;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const long long n = n0 + e;
          const float s = e == 0 ? s0 : s1;
          if (q_ok && n < n_end) {
            if (p.out_full) p.out_full[static_cast<size_t>(q) * p.N + n] = s;
            if (p.cand) {
              const u64 key = make_key(s, static_cast<uint32_t>(p.row_offset + n));
              if (key > root) {
                sts64(mypend + pcnt * HS, key);
                ++pcnt;
              }
            }
          }
        }
        if (__any_sync(0xffffffffu, pcnt > PEND - 2)) flush();
      }
      if (++as == AS) {
        as = 0;
        aphase ^= 1;
      }
    }
    flush();
    if (p.cand && q_ok) {
      const size_t qpad = static_cast<size_t>(p.qtiles) * SQ;
      for (int j = 0; j < k; ++j) p.cand[(static_cast<size_t>(sp) * k + j) * qpad + q] = lds64(myheap + j * HS);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// bitonic helpers (descending order of u64 keys) on a shared-memory array of n = power of two
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bitonic_sort_desc(u64* s, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const u64 a = s[lo], b = s[hi];
        if ((a < b) == desc) {
          s[lo] = b;
          s[hi] = a;
        }
      }
    }
  }
  __syncthreads();
}

// merge: for each query, P*k candidates -> top-k sorted.  Candidates are either packed keys laid out
// [P][k][qpad] (from scan_topk_kernel / row_topk_kernel) or (score, idx) pairs [P][Q][k] (C ABI), list p starting
// `pstride` elements after list p - 1 (Q * k when dense; 2 * Q * k for the exchange buffer [P][2][Q][k]).
__global__ void __launch_bounds__(1024)
topk_merge_kernel(const u64* __restrict__ cand_keys, size_t qpad, const float* __restrict__ cand_score,
                  const int32_t* __restrict__ cand_idx, size_t pstride, int P, int Q, int k, int npow2,
                  float* __restrict__ out_score, int32_t* __restrict__ out_idx, int group_rows, size_t group_stride) {
  extern __shared__ u64 skeys[];
  const int q = blockIdx.x;
  // grouped output: queries [g * group_rows, (g + 1) * group_rows) write their [group_rows, k] block at g * group_stride
  // (the multi-GPU exchange buffer [dest rank][scores | rows][Bq][k]); dense [Q, k] when group_rows == 0
  const size_t ob = group_rows > 0 ? static_cast<size_t>(q / group_rows) * group_stride +
                                         static_cast<size_t>(q % group_rows) * k
                                   : static_cast<size_t>(q) * k;
  const int total = P * k;
  for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
    u64 key = 0;
    if (i < total) {
      const int pp = i / k, j = i % k;
      if (cand_keys) {
        key = cand_keys[(static_cast<size_t>(pp) * k + j) * qpad + q];
      } else {
        const size_t o = static_cast<size_t>(pp) * pstride + static_cast<size_t>(q) * k + j;
        const int32_t idx = cand_idx[o];
        key = idx < 0 ? 0 : make_key(cand_score[o], static_cast<uint32_t>(idx));
      }
    }
    skeys[i] = key;
  }
  bitonic_sort_desc(skeys, npow2);
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    float sc;
    int32_t ix;
    decode_key(j < npow2 ? skeys[j] : 0, sc, ix);
    if (out_score) out_score[ob + j] = sc;
    if (out_idx) out_idx[ob + j] = ix;
  }
}

// large-k path: top-k of each 4096-wide segment of a full score row
__global__ void __launch_bounds__(1024)
row_topk_kernel(const float* __restrict__ full, long long N, long long row_offset, int k, size_t qpad,
                u64* __restrict__ cand) {
  __shared__ u64 skeys[SEG];
  const int seg = blockIdx.x, q = blockIdx.y;
  const long long n0 = static_cast<long long>(seg) * SEG;
  for (int i = threadIdx.x; i < SEG; i += blockDim.x) {
    const long long n = n0 + i;
    skeys[i] = n < N ? make_key(full[static_cast<size_t>(q) * N + n], static_cast<uint32_t>(row_offset + n)) : 0;
  }
  bitonic_sort_desc(skeys, SEG);
  for (int j = threadIdx.x; j < k; j += blockDim.x)
    cand[(static_cast<size_t>(seg) * k + j) * qpad + q] = j < SEG ? skeys[j] : 0;
}

static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

static int launch_merge(const u64* keys, size_t qpad, const float* cs, const int32_t* ci, int P, int Q, int k,
                        float* out_score, int32_t* out_idx, cudaStream_t st, size_t pstride = 0, int group_rows = 0,
                        size_t group_stride = 0) {
  if (pstride == 0) pstride = static_cast<size_t>(Q) * k;
  const int np2 = next_pow2(P * k < 2 ? 2 : P * k);
  const size_t smem = static_cast<size_t>(np2) * 8;
  SPRC_REQUIRE(smem <= 200 * 1024, "topk_merge: %d candidates per query exceed the merge capacity", P * k);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    SPRC_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = 200 * 1024;
  }
  prof_begin(st);
  topk_merge_kernel<<<Q, 1024, smem, st>>>(keys, qpad, cs, ci, pstride, P, Q, k, np2, out_score, out_idx, group_rows,
                                           group_stride);
  prof_end(PROF_MERGE, 0.0, (double)P * k * Q * 8, st);
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

int topk_merge(const float* cand_score, const int32_t* cand_idx, int P, int Q, int k, float* out_score,
               int32_t* out_idx, cudaStream_t st, size_t pstride) {
  SPRC_REQUIRE(P > 0 && Q > 0 && k > 0, "topk_merge: empty problem");
  return launch_merge(nullptr, 0, cand_score, cand_idx, P, Q, k, out_score, out_idx, st, pstride);
}

// workspace: candidate keys of the fused path, or [full matrix +] segment candidates of the large-k path
static void scan_plan(int Q, long long N, int k, int& qtiles, int& splits, long long& ips) {
  qtiles = (Q + SQ - 1) / SQ;
  const int sms = device_sm_count();
  splits = sms / qtiles;
  if (splits < 1) splits = 1;
  const long long max_splits = (N + 3) / 4;  // at least one slot pair (4 images) per split
  if (splits > max_splits) splits = static_cast<int>(max_splits > 0 ? max_splits : 1);
  // keep the merge within capacity (splits * k <= 16384)
  while (static_cast<long long>(splits) * k > 16384 && splits > 1) --splits;
  ips = (N + splits - 1) / splits;
  ips = (ips + 3) & ~3LL;
  if (ips < 4) ips = 4;
  splits = static_cast<int>((N + ips - 1) / ips);
  if (splits < 1) splits = 1;
}

size_t sim_topk_workspace_bytes(int Q, int64_t N, int k, bool caller_has_full) {
  const size_t qpad = static_cast<size_t>((Q + SQ - 1) / SQ) * SQ;
  if (k <= KCAP) {   // splits * k candidate keys per (padded) query
    int qtiles, splits;
    long long ips;
    scan_plan(Q, N, k, qtiles, splits, ips);
    return static_cast<size_t>(splits) * k * qpad * 8 + 256;
  }
  const size_t nseg = static_cast<size_t>((N + SEG - 1) / SEG);
  size_t need = ((nseg * k * qpad * 8 + 255) & ~size_t(255)) + 256;
  if (!caller_has_full) need += static_cast<size_t>(Q) * N * 4 + 256;
  return need;
}

int sim_topk(const bf16* queries, int Q, const bf16* gallery, int64_t N, int64_t row_offset, int k,
             float* out_score, int32_t* out_idx, float* out_full, void* workspace, size_t workspace_bytes,
             cudaStream_t st, int out_group_rows, size_t out_group_stride) {
  SPRC_REQUIRE(out_group_rows >= 0 && (out_group_rows == 0 || out_group_stride >= (size_t)out_group_rows * k),
               "sim_topk: grouped output needs group_stride >= group_rows * k");
  SPRC_REQUIRE(Q > 0 && N > 0, "sim_topk: empty problem (Q=%d N=%lld)", Q, (long long)N);
  SPRC_REQUIRE(N * 32 < (1LL << 31), "sim_topk: gallery shard of %lld images exceeds the 2^31-token TMA coordinate "
               "range; shard it", (long long)N);
  const bool want_topk = (out_score || out_idx);
  SPRC_REQUIRE(!want_topk || (k > 0 && k <= 1024), "sim_topk: k=%d outside [1, 1024]", k);
  SPRC_REQUIRE(want_topk || out_full, "sim_topk: no output requested");
  const bool fused = want_topk && k <= KCAP;
  int qtiles, splits;
  long long ips;
  scan_plan(Q, N, fused ? k : 1, qtiles, splits, ips);
  const size_t qpad = static_cast<size_t>(qtiles) * SQ;

  CUtensorMap tmG;
  SPRC_REQUIRE((reinterpret_cast<uintptr_t>(queries) & 15) == 0, "sim_topk: queries must be 16-byte aligned");
  SPRC_TRY(make_tmap_bf16(&tmG, gallery, 256, (uint64_t)N * 32, 1, 256, 0, 64, ST, 1, 2));

  ScanParams p;
  p.queries = reinterpret_cast<const uint4*>(queries);
  p.Q = Q;
  p.N = N;
  p.row_offset = row_offset;
  p.k = fused ? k : 0;
  p.qtiles = qtiles;
  p.splits = splits;
  const int smem_bytes = SSLOTS * SLOT_BYTES + (p.k + PEND) * SQ * 8 + 256 + 1024;
  p.imgs_per_split = ips;
  p.fp16 = act_fp16();
  p.out_full = out_full;
  p.cand = nullptr;
  u64* cand = static_cast<u64*>(workspace);
  float* full = out_full;
  if (fused) {
    const size_t need = static_cast<size_t>(splits) * k * qpad * 8;
    SPRC_REQUIRE(workspace && workspace_bytes >= need, "sim_topk: workspace too small (%zu < %zu)", workspace_bytes,
                 need);
    p.cand = cand;
  } else if (want_topk) {
    // large k: materialise the score matrix (in the caller's buffer if given), then segment top-k + merge
    const int nseg = static_cast<int>((N + SEG - 1) / SEG);
    const size_t need_c = static_cast<size_t>(nseg) * k * qpad * 8;
    size_t need = need_c;
    if (!full) need += static_cast<size_t>(Q) * N * 4 + 256;
    SPRC_REQUIRE(workspace && workspace_bytes >= need, "sim_topk: workspace too small (%zu < %zu)", workspace_bytes,
                 need);
    if (!full) {
      full = reinterpret_cast<float*>(static_cast<char*>(workspace) + ((need_c + 255) & ~size_t(255)));
      p.out_full = full;
    }
  }

  static bool attr_set = false;
  if (!attr_set) {
    SPRC_CUDA(cudaFuncSetAttribute(scan_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   SSLOTS * SLOT_BYTES + (KCAP + PEND) * SQ * 8 + 256 + 1024));
    attr_set = true;
  }
  prof_begin(st);
  scan_topk_kernel<<<qtiles * splits, 192, smem_bytes, st>>>(tmG, p);
  prof_end(PROF_SCAN, 2.0 * Q * (double)N * 32 * 256, (double)N * 32 * 256 * 2 + (double)Q * 512, st);
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  if (fused)
    return launch_merge(cand, qpad, nullptr, nullptr, splits, Q, k, out_score, out_idx, st, 0, out_group_rows,
                        out_group_stride);
  if (want_topk) {
    const int nseg = static_cast<int>((N + SEG - 1) / SEG);
    SPRC_REQUIRE(static_cast<long long>(nseg) * k <= 16384 * 1, "sim_topk: N=%lld with k=%d exceeds the merge capacity",
                 (long long)N, k);
    dim3 grid(nseg, Q);
    row_topk_kernel<<<grid, 1024, 0, st>>>(full, N, row_offset, k, qpad, cand);
    count_launch();
    SPRC_CUDA(cudaGetLastError());
    return launch_merge(cand, qpad, nullptr, nullptr, nseg, Q, k, out_score, out_idx, st, 0, out_group_rows,
                        out_group_stride);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// gather_scores: sim of explicit (query, row) pairs, one warp per pair (lane = gallery token)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_scores_kernel(const bf16* __restrict__ queries, int Q, const bf16* __restrict__ gallery, long long N,
                     const int32_t* __restrict__ rows, int m, float* __restrict__ out, int fp16) {
  const int pair = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (pair >= Q * m) return;
  const int lane = threadIdx.x & 31;
  const int q = pair / m;
  const int32_t n = rows[pair];
  if (n < 0 || n >= N) {
    if (lane == 0) out[pair] = -INFINITY;
    return;
  }
  const uint4* g = reinterpret_cast<const uint4*>(gallery + (static_cast<size_t>(n) * 32 + lane) * 256);
  const uint4* qv = reinterpret_cast<const uint4*>(queries + static_cast<size_t>(q) * 256);
  float acc = 0.f;
#pragma unroll 4
  for (int i = 0; i < 32; ++i) {
    const uint4 a = g[i], b = __ldg(qv + i);
    const unsigned short* a2 = reinterpret_cast<const unsigned short*>(&a);
    const unsigned short* b2 = reinterpret_cast<const unsigned short*>(&b);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(from_act(a2[j], fp16), from_act(b2[j], fp16), acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc = fmaxf(acc, __shfl_xor_sync(0xffffffffu, acc, o));
  if (lane == 0) out[pair] = acc;
}

int gather_scores(const bf16* queries, int Q, const bf16* gallery, int64_t N, const int32_t* rows, int m,
                  float* out, cudaStream_t st) {
  if (Q <= 0 || m <= 0) return 0;
  const int pairs = Q * m;
  gather_scores_kernel<<<(pairs + 7) / 8, 256, 0, st>>>(queries, Q, gallery, N, rows, m, out, act_fp16());
  count_launch();
  SPRC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sprc
