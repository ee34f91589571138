// tcgen05.mma issue-rate microbenchmark (sm_100a): cycles per 128 x N x 16 MMA for operands already resident
// in shared memory / TMEM, as a function of N, operand source of A (smem descriptor vs TMEM) and whether
// consecutive MMAs accumulate into the same TMEM columns.  Build: make -C tools/microbench ; run: ./mma_rate
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../sprc_b200/csrc/ptx.cuh"

using namespace sprc;

struct Cfg { int N; int ts; int nacc; int iters; int ctas; };

__global__ void __launch_bounds__(64, 1) mma_rate_kernel(Cfg c, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;               // 128 rows x 64 bf16 (16 KB), K-major SW128
  uint8_t* sB = smem + 16384;       // 256 rows x 64 bf16 (32 KB)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  const int warp = threadIdx.x >> 5;
  if (warp == 1) tmem_alloc(slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_16(128, c.N, 0);
    const uint64_t da = umma_desc_k_sw128(smem_u32(sA));
    const uint64_t db = umma_desc_k_sw128(smem_u32(sB));
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < c.iters; ++it) {
        const uint32_t d = tm + 128 + (it % c.nacc) * 128;   // accumulators after the A columns
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (c.ts) umma_bf16_ts(d, tm + k * 8, db + 2 * k, idesc, 1u);
          else umma_bf16(d, da + 2 * k, db + 2 * k, idesc, 1u);
        }
      }
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    if (elect_one()) {
      t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int Ns[] = {32, 64, 128, 256};
  printf("%6s %4s %5s %6s %12s %10s %8s\n", "N", "A", "nacc", "ctas", "clk/MMA", "floor", "ratio");
  for (int ctas : {1, 148})
    for (int ts = 0; ts < 2; ++ts)
      for (int nacc : {1, 2, 3})
        for (int N : Ns) {
          if (nacc > 1 && N > 128) continue;
          Cfg c{N, ts, nacc, 2000, ctas};
          mma_rate_kernel<<<ctas, 64, 64 * 1024>>>(c, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long clk;
          cudaMemcpy(&clk, d_out, 8, cudaMemcpyDeviceToHost);
          const double per = (double)clk / (c.iters * 4.0);
          const double floor_ = 128.0 * N / 256.0;
          printf("%6d %4s %5d %6d %12.1f %10.1f %8.2f\n", N, ts ? "tmem" : "smem", nacc, ctas, per, floor_, per / floor_);
        }
  return 0;
}
