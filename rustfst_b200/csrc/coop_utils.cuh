// coop_utils.cuh — building blocks of the persistent cooperative kernels: CTA-wide scans, per-CTA partial prefixes kept
// in shared memory (every CTA recomputes the same prefix from the same global partial array, so control flow stays
// uniform across the grid without broadcasts), shared-memory segment search for load-balanced tiles.
#pragma once
#include <cooperative_groups.h>

#include "device_common.cuh"

namespace b200 {
namespace coop {

constexpr int kCoopThreads = 256;

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Grid-wide barrier on a monotonically increasing arrival counter (no reset, so no second phase): the k-th barrier
// completes when the counter reaches k * gridDim.x.  Requires all CTAs to be co-resident (cooperative launch).
// Lighter than cg::grid_group::sync(): one release atomic and acquire-load polling by one thread per CTA, no L1
// invalidation inside the spin.  Cross-CTA data read after the barrier must bypass L1 (__ldcg / volatile / atomics).
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& epoch) {
  __syncthreads();
  epoch++;
  if (threadIdx.x == 0) {
    const unsigned int target = epoch * gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while ((int)(v - target) < 0);
  }
  __syncthreads();
}

// Exclusive scan of one value per thread over the CTA (256 threads); returns the exclusive prefix, total in `total`.
__device__ __forceinline__ uint32_t cta_exclusive_scan(uint32_t v, uint32_t* s_warp /*8*/, uint32_t& total) {
  const uint32_t wl = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o); if ((int)wl >= o) incl += u; }
  __syncthreads();  // protect s_warp reuse
  if (wl == 31) s_warp[wid] = incl;
  __syncthreads();
  uint32_t before = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kCoopThreads / 32; w++) { uint32_t c = s_warp[w]; if (w < (int)wid) before += c; tot += c; }
  total = tot;
  return before + incl - v;
}
// Two independent exclusive scans sharing the shuffles and barriers of one (s_warp2 holds 2 x 8 words).
__device__ __forceinline__ void cta_exclusive_scan2(uint32_t a, uint32_t b, uint32_t* s_warp2, uint32_t& ex_a,
                                                    uint32_t& ex_b, uint32_t& tot_a, uint32_t& tot_b) {
  const uint32_t wl = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t ia = a, ib = b;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t ua = __shfl_up_sync(0xFFFFFFFFu, ia, o), ub = __shfl_up_sync(0xFFFFFFFFu, ib, o);
    if ((int)wl >= o) { ia += ua; ib += ub; }
  }
  __syncthreads();
  if (wl == 31) { s_warp2[wid] = ia; s_warp2[kCoopThreads / 32 + wid] = ib; }
  __syncthreads();
  uint32_t ba = 0, bb = 0, ta = 0, tb = 0;
#pragma unroll
  for (int w = 0; w < kCoopThreads / 32; w++) {
    uint32_t ca = s_warp2[w], cb = s_warp2[kCoopThreads / 32 + w];
    if (w < (int)wid) { ba += ca; bb += cb; }
    ta += ca; tb += cb;
  }
  ex_a = ba + ia - a; ex_b = bb + ib - b; tot_a = ta; tot_b = tb;
}
// s_out[0..n] = exclusive prefix of src[0..n) (n <= 2048), computed by the whole CTA; s_out[n] = total.
__device__ __forceinline__ void cta_prefix_to_smem(const uint32_t* __restrict__ src, uint32_t n, uint32_t* s_out,
                                                   uint32_t* s_warp) {
  uint32_t v[8], sum = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) { uint32_t i = threadIdx.x * 8 + k; v[k] = i < n ? __ldcg(&src[i]) : 0u; sum += v[k]; }
  uint32_t total;
  uint32_t run = cta_exclusive_scan(sum, s_warp, total);
#pragma unroll
  for (int k = 0; k < 8; k++) { uint32_t i = threadIdx.x * 8 + k; if (i < n) s_out[i] = run; run += v[k]; }
  if (threadIdx.x == 0) s_out[n] = total;
  __syncthreads();
}
// largest k in [0, m) with s[k] <= x   (s non-decreasing, s[0] <= x)
__device__ __forceinline__ uint32_t smem_segment(const uint32_t* s, uint32_t m, uint32_t x) {
  uint32_t lo = 0, hi = m;
  while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (s[mid] <= x) lo = mid; else hi = mid; }
  return lo;
}


}  // namespace coop
}  // namespace b200
