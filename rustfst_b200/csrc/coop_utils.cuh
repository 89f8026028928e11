// coop_utils.cuh — building blocks of the persistent cooperative kernels: CTA-wide scans, per-CTA partial prefixes kept
// in shared memory (every CTA recomputes the same prefix from the same global partial array, so control flow stays
// uniform across the grid without broadcasts), shared-memory segment search for load-balanced tiles.
#pragma once
#include <cooperative_groups.h>

#include "device_common.cuh"

namespace b200 {
namespace coop {

// Threads per CTA of the persistent kernels.  One 768-thread CTA per SM (148 CTAs) measured best for the compose
// kernel on C3: 3 x 256 threads per SM 5.08 ms, 2 x 512 5.11, 1 x 512 5.19, 1 x 768 4.78 (same warps per SM as
// 3 x 256, but a third of the words in every count exchange and no scheduling skew between co-resident CTAs).
// -DB200_COOP_THREADS=256|512|640 rebuilds the other shapes.
#ifndef B200_COOP_THREADS
#define B200_COOP_THREADS 768
#endif
constexpr int kCoopThreads = B200_COOP_THREADS;


__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Last-resort watchdog of the spin loops below: a grid-wide wait that lasts longer than 20 s can only be a bug (a CTA
// that never arrives); the kernel then traps — the call fails with a CUDA error — instead of hanging the device.
// (k_compose_ws has its own, recoverable watchdog; see compose_ws.cu.)
__device__ __forceinline__ void spin_watchdog(unsigned long long& t_start) {
  const unsigned long long now = globaltimer_ns();
  if (!t_start) t_start = now;
  else if (now - t_start > 20ull * 1000 * 1000 * 1000) __trap();
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
    unsigned long long t_start = 0;
    for (unsigned int it = 1;; it++) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if ((int)(v - target) >= 0) break;
      if ((it & 0xFFFFu) == 0) spin_watchdog(t_start);
    }
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
// Exclusive scans of a 0/1 flag (ballot + popc) and of a count (shuffles) over the CTA, sharing the two barriers.
__device__ __forceinline__ void cta_exclusive_scan_flag_count(bool flag, uint32_t b, uint32_t* s_warp2, uint32_t& ex_a,
                                                              uint32_t& ex_b, uint32_t& tot_a, uint32_t& tot_b) {
  const uint32_t wl = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t bal = __ballot_sync(0xFFFFFFFFu, flag);
  const uint32_t ia_ex = __popc(bal & ((1u << wl) - 1u));
  uint32_t ib = b;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t ub = __shfl_up_sync(0xFFFFFFFFu, ib, o); if ((int)wl >= o) ib += ub; }
  __syncthreads();
  if (wl == 31) { s_warp2[wid] = __popc(bal); s_warp2[kCoopThreads / 32 + wid] = ib; }
  __syncthreads();
  uint32_t ba = 0, bb = 0, ta = 0, tb = 0;
#pragma unroll
  for (int w = 0; w < kCoopThreads / 32; w++) {
    uint32_t ca = s_warp2[w], cb = s_warp2[kCoopThreads / 32 + w];
    if (w < (int)wid) { ba += ca; bb += cb; }
    ta += ca; tb += cb;
  }
  ex_a = ba + ia_ex; ex_b = bb + ib - b; tot_a = ta; tot_b = tb;
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
// ---- epoch-tagged count exchange: a grid barrier and a broadcast in one round trip.
// Each CTA publishes one word (tag << 32 | count) after its writes of the phase; every CTA then spins until all
// gridDim words carry the current tag and builds the exclusive prefix in shared memory.  Because every CTA waits for
// every other CTA's word, passing the wait is a full grid barrier (release: fence + store by thread 0 after a CTA
// barrier; acquire: relaxed polling + fence + CTA barrier).  Requires co-resident CTAs (cooperative launch) and a
// zero-initialised array; tags start at 1 and only grow.
__device__ __forceinline__ void publish_count(unsigned long long* part, uint32_t c, uint32_t tag, uint32_t count) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long v = ((unsigned long long)tag << 32) | count;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(part + c), "l"(v) : "memory");
  }
}
__device__ __forceinline__ void cta_prefix_wait_to_smem(const unsigned long long* part, uint32_t n, uint32_t tag,
                                                        uint32_t* s_out, uint32_t* s_warp, uint32_t poll_ns = 0) {
  for (uint32_t i = threadIdx.x; i < n; i += kCoopThreads) {
    unsigned long long v;
    unsigned long long t_start = 0;
    for (unsigned int it = 1;; it++) {
      asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(part + i) : "memory");
      if ((uint32_t)(v >> 32) == tag) break;
      if (poll_ns) __nanosleep(poll_ns);
      if ((it & 0xFFFFu) == 0) spin_watchdog(t_start);
    }
    s_out[i] = (uint32_t)v;
  }
  __threadfence();
  __syncthreads();
  if (n <= 2 * kCoopThreads) {  // the usual grid (<= 512 CTAs): two entries per thread instead of eight
    const uint32_t i0 = threadIdx.x * 2;
    const uint32_t a = i0 < n ? s_out[i0] : 0u, b = i0 + 1 < n ? s_out[i0 + 1] : 0u;
    uint32_t total;
    const uint32_t run = cta_exclusive_scan(a + b, s_warp, total);
    if (i0 < n) s_out[i0] = run;
    if (i0 + 1 < n) s_out[i0 + 1] = run + a;
    if (threadIdx.x == 0) s_out[n] = total;
    __syncthreads();
    return;
  }
  uint32_t v[8], sum = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) { uint32_t i = threadIdx.x * 8 + k; v[k] = i < n ? s_out[i] : 0u; sum += v[k]; }
  uint32_t total;
  uint32_t run = cta_exclusive_scan(sum, s_warp, total);
#pragma unroll
  for (int k = 0; k < 8; k++) { uint32_t i = threadIdx.x * 8 + k; if (i < n) s_out[i] = run; run += v[k]; }
  if (threadIdx.x == 0) s_out[n] = total;
  __syncthreads();
}
// 128-bit load of a 16-byte record that other CTAs update concurrently (bypasses L1; one transaction)
__device__ __forceinline__ uint4 ld_volatile_u4(const void* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// largest k in [0, m) with s[k] <= x   (s non-decreasing, s[0] <= x)
__device__ __forceinline__ uint32_t smem_segment(const uint32_t* s, uint32_t m, uint32_t x) {
  uint32_t lo = 0, hi = m;
  while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (s[mid] <= x) lo = mid; else hi = mid; }
  return lo;
}


}  // namespace coop
}  // namespace b200
