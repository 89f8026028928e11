// bulk_async.cuh — sm_100a asynchronous-copy primitives used by the persistent kernels: mbarrier transaction barriers,
// 1-D bulk copies through the TMA unit (cp.async.bulk, SASS UBLKCP) in both directions, the proxy fences that order
// them against ordinary loads/stores, and the 128-bit compare-and-swap (ATOMG.CAS.128) of the state table.
#pragma once
#include <cstdint>

namespace b200 {
namespace bulk {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier (shared memory, CTA scope).  `parity` of a wait is the phase bit of the use being waited for: 0 for
// the first completion after init, then alternating.
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// one arrival that also announces `bytes` of asynchronous-copy traffic to complete on the barrier
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " WAIT_%=:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra DONE_%=;\n"
      " bra WAIT_%=;\n"
      " DONE_%=:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ---- 1-D bulk copies.  Addresses 16-byte aligned, size a multiple of 16 bytes.
// global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes) : "memory");
}
__device__ __forceinline__ void commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest N groups of this thread have finished READING their shared-memory source
template <int N>
__device__ __forceinline__ void wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void wait_group() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// Ordinary (generic-proxy) writes must be made visible to the asynchronous proxy before a bulk copy reads them.
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// ---- 128-bit compare-and-swap on a 16-byte aligned global location; returns the previous contents.
struct U128 { unsigned long long lo, hi; };
__device__ __forceinline__ U128 cas128(void* p, U128 cmp, U128 swp) {
  U128 old;
  asm volatile(
      "{\n"
      " .reg .b128 c, s, o;\n"
      " mov.b128 c, {%2, %3};\n"
      " mov.b128 s, {%4, %5};\n"
      " atom.global.relaxed.gpu.cas.b128 o, [%6], c, s;\n"
      " mov.b128 {%0, %1}, o;\n"
      "}"
      : "=l"(old.lo), "=l"(old.hi) : "l"(cmp.lo), "l"(cmp.hi), "l"(swp.lo), "l"(swp.hi), "l"(p) : "memory");
  return old;
}

}  // namespace bulk
}  // namespace b200
