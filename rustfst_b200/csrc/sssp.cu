// sssp.cu — single shortest path and forward shortest distances over the tropical semiring on B200.
//
// Replaces (paths relative to /root/reference):
//   rustfst/src/algorithms/shortest_path.rs:173-239  single_shortest_path (queue-ordered label-correcting relaxation)
//   rustfst/src/algorithms/shortest_path.rs:241-282  single_shortest_path_backtrace
//   rustfst/src/algorithms/shortest_distance.rs:153-237  shortest_distance (forward; feeds the n-best search, nshortest.cu)
//   rustfst/src/algorithms/queues/{state_order,top_order,lifo,fifo,trivial,scc}_queue.rs  (serial kernels only)
//
// Three device paths, selected per call (SsspStats::path):
//
//  (0) PARALLEL path — taken when the reference would process states in a topological order (StateOrderQueue on a
//      TOP_SORTED input, or TopOrderQueue).  Distances are exact minima computed by frontier relaxation waves with
//      atomicMin on order-preserving integer images of the f32 distances and warp-aggregated frontier
//      compaction (k_relax_coop: every wave inside one cooperative launch).  One more edge scan (k_parents) then
//      (a) selects for every state the parent the reference would end up with: the first candidate, in its
//      processing order (order[src], arc position), that attains the minimum, and (b) CERTIFIES that the
//      reference's approximate relaxation test (`d != min(d, c)` with KDELTA = 1/1024 tolerance,
//      shortest_path.rs:225; `|d - min(d, c)| <= delta` for shortest_distance.rs:217) cannot have kept a non-minimal
//      candidate: no candidate value c of any state lies in (m, m + tolerance] where m is that state's minimum.
//      When the certificate holds the reference's sequential fold provably ends with the same (distance, parent)
//      at every state.
//
//  (2) ORDER-FAITHFUL PARALLEL FOLD — same queue kinds, taken when the certificate fails (near-ties): a reverse CSR
//      with in-arcs sorted by the reference's processing order, states scheduled level by level (Kahn), one thread
//      per state replays the reference's fold verbatim (k_of_fold).
//
//  (1) ORDER-FAITHFUL SERIAL kernels — one device thread replays the reference's loop verbatim (queue discipline
//      included) for everything else: cyclic inputs (SccQueue / LIFO) and graphs the fold finds cyclic
//      (k_serial_sssp, k_serial_sdist).
//
// The backtrace runs on the device as well; only the shortest path itself (a few hundred bytes) returns to the host.
#include <cooperative_groups.h>
#include <cuda_pipeline.h>

#include <cstdlib>
#include <vector>

#include "algos.h"
#include "bulk_async.cuh"
#include "coop_utils.cuh"

namespace cg = cooperative_groups;

namespace b200 {
namespace {

constexpr uint32_t kEncInf = 0xFF800000u;  // enc(+inf)
constexpr unsigned long long kNoParent = ~0ull;

__host__ __device__ __forceinline__ uint32_t enc_f32(float f) {  // monotone float -> uint32
#if defined(__CUDA_ARCH__)
  uint32_t b = __float_as_uint(f);
#else
  uint32_t b; memcpy(&b, &f, 4);
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_f32(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// warp-aggregated append of `v` for lanes with pred set
__device__ __forceinline__ void warp_push(bool pred, uint32_t v, uint32_t* __restrict__ list,
                                          uint32_t* __restrict__ count) {
  uint32_t active = __activemask();
  uint32_t m = __ballot_sync(active, pred);
  if (!m) return;
  uint32_t lane = threadIdx.x & 31, leader = __ffs(m) - 1, base = 0;
  if (lane == leader) base = atomicAdd(count, __popc(m));
  base = __shfl_sync(active, base, leader);
  if (pred) list[base + __popc(m & ((1u << lane) - 1u))] = v;
}

__global__ void k_fill_u32(uint32_t* p, uint32_t v, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void k_fill_u64(unsigned long long* p, unsigned long long v, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// The whole relaxation (all waves) in one cooperative launch: every frontier state pushes d[s] (x) w over its arcs
// with atomicMin.  kG lanes share a frontier state
// and stride over its arcs (128-bit loads).  Newly improved states are first collected in a per-CTA shared-memory
// queue and flushed to the next frontier with ONE global atomic per CTA and wave (a single global counter hit by
// every warp serialises in the L2 atomic unit: ~30k same-address atomics per wave on C4).  Three frontier counters
// rotate so that no reset races with a read: wave w consumes cnt[w % 3] (filled during wave w-1), fills
// cnt[(w+1) % 3] and clears cnt[(w+2) % 3].
// out[0] = waves, out[1] = non-convergence flag, out64[0] = arcs relaxed, out64[1] = states settled.
constexpr uint32_t kQueueCap = 2048;
template <int kG, int kT, bool kPreTest, int kS>
__global__ void __launch_bounds__(kT, 2048 / kT)  // the kernel scales with resident threads: stay at 2048 per SM
k_relax_coop(const uint32_t* __restrict__ off, const Tr* __restrict__ arcs, uint32_t n, uint32_t* __restrict__ dist,
             uint32_t* __restrict__ stamp, uint32_t* __restrict__ fr_a, uint32_t* __restrict__ fr_b,
             uint32_t* __restrict__ cnt /*3*/, uint32_t* __restrict__ out, unsigned long long* __restrict__ out64,
             unsigned long long budget) {
  unsigned int bar_epoch = 0;  // out[2] = arrival counter of the grid barrier (zero-initialised)
  uint32_t visits = 0;  // frontier entries so far, saturating (every thread reads the same counters: uniform)
  uint32_t fresh = 0;   // states this thread reached for the first time (summed into out64[2] wave by wave)
  bool over_budget = false;
  const uint32_t budget32 = budget > 0xFFFFFFFEull ? 0xFFFFFFFFu : (uint32_t)budget;  // 0xFFFFFFFF = no budget
  constexpr uint32_t kQCap = kQueueCap * (kT / 256);
  __shared__ uint32_t s_q[kQCap];
  __shared__ uint32_t s_qn, s_gbase, s_fresh;
  const uint32_t lane = threadIdx.x % kG;
  const uint32_t groups = gridDim.x * (kT / kG);
  const uint32_t gid = blockIdx.x * (kT / kG) + threadIdx.x / kG;
  uint32_t* cur = fr_a;
  uint32_t* nxt = fr_b;
  unsigned long long relaxed = 0, settled = 0;
  uint32_t wave = 0;
  while (true) {
    const uint32_t nf = __ldcg(&cnt[wave % 3]);
    if (nf == 0 || wave > n) break;
    visits = visits + nf < visits ? 0xFFFFFFFEu : min(visits + nf, 0xFFFFFFFEu);
    // States are revisited over and over (a deep DAG with skip arcs): more than `budget` visits in all, or — early, so
    // that little is thrown away — eight times as many visits as states reached so far.
    if (budget32 != 0xFFFFFFFFu) {
      if (visits > budget32) { over_budget = true; break; }
      if (wave >= 16 && (unsigned long long)visits > 8ull * __ldcg(&out64[2]) + 65536ull) { over_budget = true; break; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) cnt[(wave + 2) % 3] = 0;
    if (threadIdx.x == 0) { s_qn = 0; if (wave == 0) s_fresh = 0; }
    __syncthreads();
    uint32_t* next_count = &cnt[(wave + 1) % 3];
    // kS frontier states per group and round: their loads, and then their atomics, are issued together (the kernel is
    // bound by the chain frontier entry -> distance / offsets -> arcs -> atomicMin -> stamp, not by bandwidth)
    for (uint32_t i = gid; i < nf; i += groups * kS) {
      uint32_t b[kS], e[kS];
      float ds[kS];
      {
        uint32_t st[kS];
#pragma unroll
        for (int q = 0; q < kS; q++) st[q] = (i + q * groups < nf) ? __ldcg(&cur[i + q * groups]) : 0xFFFFFFFFu;
#pragma unroll
        for (int q = 0; q < kS; q++) {
          b[q] = 0; e[q] = 0; ds[q] = 0.0f;
          if (st[q] != 0xFFFFFFFFu) { ds[q] = dec_f32(__ldcg(&dist[st[q]])); b[q] = off[st[q]]; e[q] = off[st[q] + 1]; }
        }
#pragma unroll
        for (int q = 0; q < kS; q++)
          if (lane == 0 && st[q] != 0xFFFFFFFFu) { relaxed += e[q] - b[q]; settled++; }
      }
      for (uint32_t r = 0;; r += kG) {
        bool any = false;
        int4 v[kS];
        bool has[kS];
#pragma unroll
        for (int q = 0; q < kS; q++) {
          any |= b[q] + r < e[q];
          has[q] = b[q] + r + lane < e[q];
          if (has[q]) v[q] = __ldg(reinterpret_cast<const int4*>(&arcs[b[q] + r + lane]));
        }
        if (!any) break;
        uint32_t ec[kS], old[kS];
#pragma unroll
        for (int q = 0; q < kS; q++) {
          old[q] = 0; ec[q] = kEncInf;
          if (has[q]) {
            const float c = w_times(ds[q], __int_as_float(v[q].z));
            if (c != w_zero()) {
              ec[q] = enc_f32(c);
              // kPreTest: a plain load filters most candidates before the atomic
              if (!kPreTest || ec[q] < __ldcg(&dist[(uint32_t)v[q].w])) old[q] = atomicMin(&dist[(uint32_t)v[q].w], ec[q]);
              if (old[q] == kEncInf && ec[q] < old[q]) fresh++;
            }
          }
        }
#pragma unroll
        for (int q = 0; q < kS; q++) {
          const uint32_t t = (uint32_t)v[q].w;
          bool push = false;
          if (has[q] && ec[q] < old[q]) push = atomicExch(&stamp[t], wave) != wave;
          // warp-aggregated append to the CTA queue; spill to the global list when the queue is full
          const uint32_t active = __activemask();
          const uint32_t m = __ballot_sync(active, push);
          if (m) {
            const uint32_t wl = threadIdx.x & 31, leader = __ffs(m) - 1;
            uint32_t qb = 0;
            if (wl == leader) qb = atomicAdd(&s_qn, __popc(m));
            qb = __shfl_sync(active, qb, leader);
            if (push) {
              const uint32_t pos = qb + __popc(m & ((1u << wl) - 1u));
              if (pos < kQCap) s_q[pos] = t;
              else nxt[atomicAdd(next_count, 1u)] = t;
            }
          }
        }
      }
    }
    {  // newly reached states of this wave: warp sums into shared memory, ONE global atomic per CTA below
      uint32_t f = fresh;
      fresh = 0;
      for (int o = 16; o > 0; o >>= 1) f += __shfl_down_sync(0xFFFFFFFFu, f, o);
      if ((threadIdx.x & 31) == 0 && f) atomicAdd(&s_fresh, f);
    }
    __syncthreads();
    const uint32_t qn = min(s_qn, kQCap);
    if (threadIdx.x == 0) {
      if (qn) s_gbase = atomicAdd(next_count, qn);
      if (s_fresh) { atomicAdd(&out64[2], (unsigned long long)s_fresh); s_fresh = 0; }  // read by everybody after the barrier
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < qn; i += kT) nxt[s_gbase + i] = s_q[i];
    wave++;
    uint32_t* tmp = cur; cur = nxt; nxt = tmp;
    coop::grid_barrier(&out[2], bar_epoch);  // lighter than cg::grid_group::sync() (measured 3.4 us less per barrier)
  }
  // statistics: one atomic per warp
  for (int o = 16; o > 0; o >>= 1) {
    relaxed += __shfl_down_sync(0xFFFFFFFFu, relaxed, o);
    settled += __shfl_down_sync(0xFFFFFFFFu, settled, o);
  }
  if ((threadIdx.x & 31) == 0) { if (relaxed) atomicAdd(&out64[0], relaxed); if (settled) atomicAdd(&out64[1], settled); }
  if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = wave; out[1] = over_budget ? 2u : (wave > n) ? 1u : 0u; }
}

// ---- IN-ORDER SWEEP for deep top-sorted DAGs ---------------------------------------------------------------------------
// Label-correcting waves revisit a state every time a shorter route reaches it; on a DAG whose arcs skip levels (targets
// up to 1000 ids ahead, longest path tens of thousands of hops) that is hundreds of visits per state and thousands of
// grid-wide waves.  What such a machine needs is the reference's own plan — states in id order, every arc once — with
// the hop latency of shared memory instead of L2.  (Tried first and dropped: a Kahn-style dataflow kernel over the whole
// grid, one visit per state, ready list + tickets: 415 ms on the 5 M-state window DAG — its critical path is the LONGEST
// path, ~55 000 hops of ~7 us through L2 atomics — against 387 ms for the waves.)
// ONE CTA walks the states of a TOP_SORTED machine in blocks of kSwB ids.  The distances of the ids
// [base, base + kSwRing) live in a shared-memory ring; a block's offsets and arcs arrive through cp.async while the
// previous block is worked on.  Inside a block, sub-blocks of kSub consecutive states are taken in id order with one
// thread per ARC: when a sub-block starts, every arc into it from earlier states has been relaxed with a final distance,
// so only arcs inside the sub-block can ask for another round (shared-memory atomicMin + one CTA barrier per round).
// Then the block retires: final distances go to global memory, the ring advances, and the slice that enters the ring
// is the minimum of what global memory held for it when the block started (candidates for ids beyond the ring are
// pushed with global atomicMin) and of the candidates this block collected for it in shared memory.
// (One thread per STATE iterating the whole block to a fixpoint took ~46 rounds per block: 122 ms instead of 89.)
// ctl[0] = 1 if an arc points backwards (the property word lied), out64 as in k_relax_coop.
constexpr uint32_t kSwRing = 8192;
// arcs of one block staged in shared memory: 11 per state (2 buffers x 88 KB at 512 states); larger blocks read global memory
template <uint32_t kSwB> constexpr uint32_t sw_arc_cap() { return 11u * kSwB; }
template <uint32_t kSwB> constexpr size_t sw_smem() { return (size_t)kSwRing * 4 + (size_t)kSwB * 4 + 2 * ((size_t)sw_arc_cap<kSwB>() * 16 + (kSwB + 1) * 4) + 24; }
template <uint32_t kSwB, uint32_t kSub>
__global__ void __launch_bounds__(kSwB)
k_relax_sweep(const uint32_t* __restrict__ off, const Tr* __restrict__ arcs, uint32_t n, uint32_t* __restrict__ dist,
              uint32_t* __restrict__ ctl, unsigned long long* __restrict__ out64) {
  constexpr uint32_t kSwThreads = kSwB, kSwArcCap = sw_arc_cap<kSwB>();
  extern __shared__ __align__(16) unsigned char sw_smem[];
  int4* const s_arcs0 = reinterpret_cast<int4*>(sw_smem);
  int4* const s_arcs1 = s_arcs0 + kSwArcCap;
  uint32_t* const s_off0 = reinterpret_cast<uint32_t*>(s_arcs1 + kSwArcCap);
  uint32_t* const s_off1 = s_off0 + (kSwB + 1);
  uint32_t* const ring = s_off1 + (kSwB + 1);
  uint32_t* const pending = ring + kSwRing;  // candidates for the slice that enters the ring when this block retires
  // two transaction barriers (8-byte aligned: every array before them is a multiple of 8 bytes... kSwB + 1 words twice = even)
  unsigned long long* const bars = reinterpret_cast<unsigned long long*>(pending + kSwB);
  const uint32_t tid = threadIdx.x;
  constexpr uint32_t kMask = kSwRing - 1u;
  for (uint32_t i = tid; i < kSwRing; i += kSwThreads) ring[i] = i < n ? __ldcg(&dist[i]) : kEncInf;
  pending[tid] = kEncInf;
  if (tid == 0) { bulk::mbar_init(&bars[0], 1); bulk::mbar_init(&bars[1], 1); bulk::fence_mbar_init(); }
  __syncthreads();
  unsigned long long relaxed = 0, settled = 0;
  bool backwards = false;
  // arc range [lo, hi) of the block that starts at state `base` (every thread reads the same two words)
  auto bounds = [&](uint32_t base, uint32_t& lo, uint32_t& hi) {
    lo = hi = 0;
    if (base < n) { lo = __ldg(&off[base]); hi = __ldg(&off[min(base + kSwB, n)]); }
  };
  // Asynchronous copy of a block's offsets (cp.async, one word per thread) and — when they fit — of its arcs into
  // buffer `which`: ONE bulk copy through the TMA unit (the arcs of consecutive states are one contiguous range of
  // 16-byte records), completion counted on the buffer's transaction barrier.  The buffer was last touched by ordinary
  // shared-memory accesses (the source tags), hence the proxy fence before the copy engine writes it again.
  auto prefetch = [&](uint32_t base, uint32_t lo, uint32_t hi, uint32_t which) {
    if (base < n) {
      uint32_t* so = which ? s_off1 : s_off0;
      int4* sa = which ? s_arcs1 : s_arcs0;
      if (base + tid <= n && tid <= kSwB) __pipeline_memcpy_async(so + tid, off + base + tid, 4);
      if (tid == 0 && base + kSwB <= n) __pipeline_memcpy_async(so + kSwB, off + base + kSwB, 4);
      if (tid == 0 && hi > lo && hi - lo <= kSwArcCap) {
        bulk::fence_async_smem();
        bulk::mbar_expect_tx(&bars[which], (hi - lo) * 16u);
        bulk::g2s(sa, arcs + lo, (hi - lo) * 16u, &bars[which]);
      }
    }
    __pipeline_commit();
  };
  uint32_t lo0, hi0, lo1, hi1;  // bounds of the current and of the next block
  uint32_t bar_phase = 0;       // bit b = parity of the next completion of bars[b] (a buffer without a bulk copy skips its turn)
  bounds(0, lo0, hi0);
  bounds(kSwB, lo1, hi1);
  prefetch(0, lo0, hi0, 0);
  for (uint32_t base = 0, it = 0; base < n; base += kSwB, it++) {
    const uint32_t cur = it & 1u;
    prefetch(base + kSwB, lo1, hi1, cur ^ 1u);      // the next block travels while this one is worked on
    uint32_t lo2, hi2;
    bounds(base + 2 * kSwB, lo2, hi2);               // consumed one block later
    // The slice [base + kSwRing, base + kSwRing + kSwB) enters the ring when this block retires.  Its distances are
    // requested now (a DRAM round trip that would otherwise sit between two blocks); candidates this block produces for
    // it are collected in `pending` and merged at the end.  Earlier blocks pushed theirs to global memory before the
    // barrier that ended them.
    const uint32_t enter_id = base + kSwRing + tid;
    const uint32_t enter_pre = enter_id < n ? __ldcg(&dist[enter_id]) : kEncInf;
    __pipeline_wait_prior(1);                        // this block's offsets have landed
    if (hi0 > lo0 && hi0 - lo0 <= kSwArcCap) { bulk::mbar_wait(&bars[cur], (bar_phase >> cur) & 1u); bar_phase ^= 1u << cur; }  // ... and its arcs
    __syncthreads();
    const uint32_t* so = cur ? s_off1 : s_off0;
    const int4* sa = cur ? s_arcs1 : s_arcs0;
    const bool staged = hi0 - lo0 <= kSwArcCap;
    const uint32_t live_states = min(kSwB, n - base);
    // Tag every staged arc with its source (the label fields are not needed here), one thread per state; the relaxation
    // below runs one thread per ARC.
    if (staged && tid < live_states) {
      int4* const sa_w = cur ? s_arcs1 : s_arcs0;
      for (uint32_t k = so[tid]; k < so[tid + 1]; k++) sa_w[k - lo0].x = (int)tid;
    }
    __syncthreads();
    // Sub-blocks of kSub consecutive states, in id order: when a sub-block starts, every arc into it from earlier states
    // has been relaxed with a final distance, so only arcs INSIDE the sub-block can ask for another round.
    for (uint32_t j0 = 0; j0 < live_states; j0 += kSub) {
      const uint32_t j1 = min(j0 + kSub, live_states);
      const uint32_t a_lo = so[j0], a_hi = so[j1];
      while (true) {
        bool changed = false;
        for (uint32_t k = a_lo + tid; k < a_hi; k += kSwThreads) {
          int4 v;
          uint32_t src;
          if (staged) {
            v = sa[k - lo0];
            src = (uint32_t)v.x;
          } else {  // a block with more arcs than the staging buffer holds: global reads, source by binary search
            v = __ldg(reinterpret_cast<const int4*>(&arcs[k]));
            uint32_t l = j0, h = j1;
            while (h - l > 1) { const uint32_t m = (l + h) >> 1; if (so[m] <= k) l = m; else h = m; }
            src = l;
          }
          const uint32_t s = base + src;
          const uint32_t d = ring[s & kMask];
          if (d == kEncInf) continue;
          relaxed++;
          const float c = w_times(dec_f32(d), __int_as_float(v.z));
          if (c == w_zero()) continue;
          const uint32_t ec = enc_f32(c), t = (uint32_t)v.w;
          if (t <= s) { backwards = true; continue; }
          if (t - base < kSwRing) {  // in the ring (t > s >= base)
            const uint32_t old = atomicMin(&ring[t & kMask], ec);
            if (ec < old && t < base + j1) changed = true;  // a state of this sub-block improved: one more round
          } else if (t - base < kSwRing + kSwB) {
            atomicMin(&pending[t - base - kSwRing], ec);
          } else {
            atomicMin(&dist[t], ec);
          }
        }
        if (!__syncthreads_or(changed)) break;
      }
    }
    if (tid < live_states) {
      const uint32_t d = ring[(base + tid) & kMask];
      dist[base + tid] = d;
      if (d != kEncInf) settled++;
    }
    bulk::fence_async_smem();  // my ordinary accesses to this block's arc buffer precede the next bulk copy into it
    __syncthreads();  // everybody is done with the slice that is overwritten next
    // ids [base + kSwRing, base + kSwRing + kSwB) enter the ring in the slots the retired block leaves
    ring[enter_id & kMask] = min(enter_pre, pending[tid]);
    pending[tid] = kEncInf;
    lo0 = lo1; hi0 = hi1; lo1 = lo2; hi1 = hi2;
    // the next iteration's barrier (after its wait) orders these ring writes before anybody reads them
  }
  __pipeline_wait_prior(0);
  if (backwards) atomicExch(&ctl[0], 1u);
  for (int o = 16; o > 0; o >>= 1) {
    relaxed += __shfl_down_sync(0xFFFFFFFFu, relaxed, o);
    settled += __shfl_down_sync(0xFFFFFFFFu, settled, o);
  }
  if ((tid & 31u) == 0) { if (relaxed) atomicAdd(&out64[0], relaxed); if (settled) atomicAdd(&out64[1], settled); }
}

// Parent selection + certificate over all arcs of reached states.  flags[0] = certificate violations.
// kAbs selects the approximate test the caller's reference loop uses: false = the KDELTA `==` of
// semiring.rs:159-168 (single_shortest_path), true = approx_equal(.., delta) = |a - b| <= delta
// (utils_float.rs:1-3, shortest_distance.rs:217).  pkey may be null (distances only).
template <bool kAbs>
__global__ void __launch_bounds__(kThreads)
k_parents(const uint32_t* __restrict__ off, const Tr* __restrict__ arcs, uint32_t n,
          const uint32_t* __restrict__ dist, const uint32_t* __restrict__ order,
          unsigned long long* __restrict__ pkey, uint32_t* __restrict__ flags, float delta) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  uint32_t es = dist[s];
  if (es == kEncInf) return;
  float ds = dec_f32(es);
  unsigned long long hi = (unsigned long long)(order ? order[s] : s) << 32;
  uint32_t b = off[s], e = off[s + 1];
  bool bad = false;
  for (uint32_t k = b; k < e; k++) {
    int4 v = __ldg(reinterpret_cast<const int4*>(&arcs[k]));
    float c = w_times(ds, __int_as_float(v.z));
    if (c == w_zero()) continue;
    uint32_t t = (uint32_t)v.w;
    float m = dec_f32(dist[t]);
    if (c == m) { if (pkey) atomicMin(&pkey[t], hi | (k - b)); }
    else if (kAbs ? (fabsf(c - m) <= delta) : !(c > m + delta)) bad = true;
  }
  if (bad) atomicAdd(&flags[0], 1u);
}
__global__ void k_decode_dist(const uint32_t* __restrict__ enc, uint32_t n, float* __restrict__ out) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) out[s] = dec_f32(enc[s]);
}

// Best final state: key = (enc(d[s] (x) rho(s)) << 32) | order[s]   (shortest_path.rs:214-220)
__global__ void k_final_min(const float* __restrict__ fin, uint32_t n, const uint32_t* __restrict__ dist,
                            const uint32_t* __restrict__ order, unsigned long long* __restrict__ fkey) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long key = kNoParent;
  if (s < n) {
    float rho = fin[s];
    if (rho != w_zero() && dist[s] != kEncInf)
      key = ((unsigned long long)enc_f32(w_times(dec_f32(dist[s]), rho)) << 32) | (order ? order[s] : s);
  }
  // one atomic per warp: a lattice has ~1e5 final states and they all hit the same 8 bytes
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, key, o);
    key = other < key ? other : key;
  }
  if ((threadIdx.x & 31) == 0 && key != kNoParent) atomicMin(fkey, key);
}
__global__ void k_final_check(const float* __restrict__ fin, uint32_t n, const uint32_t* __restrict__ dist,
                              const unsigned long long* __restrict__ fkey, uint32_t* __restrict__ flags) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  float rho = fin[s];
  if (rho == w_zero() || dist[s] == kEncInf || *fkey == kNoParent) return;
  float v = w_times(dec_f32(dist[s]), rho);
  float m = dec_f32((uint32_t)(*fkey >> 32));
  if (v != m && !(v > m + kDelta)) atomicAdd(&flags[0], 1u);
}
__global__ void k_invert_order(const uint32_t* __restrict__ order, uint32_t n, uint32_t* __restrict__ inv) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) inv[order[s]] = s;
}

// Walk the parent chain from the best final state (shortest_path.rs:241-282).  out_arcs[k-1] is the single arc of
// output state k (k >= 1), already pointing at output state k-1; meta = {found, length L (arcs), f_parent, overflow}.
__global__ void k_backtrace_keys(const uint32_t* __restrict__ off, const Tr* __restrict__ arcs, uint32_t n,
                                 const unsigned long long* __restrict__ pkey,
                                 const unsigned long long* __restrict__ fkey, const uint32_t* __restrict__ inv_order,
                                 Tr* __restrict__ out_arcs, uint32_t cap, uint32_t* __restrict__ meta) {
  if (blockIdx.x || threadIdx.x) return;
  meta[0] = meta[1] = meta[2] = meta[3] = 0;
  if (*fkey == kNoParent) return;
  uint32_t ford = (uint32_t)(*fkey & 0xFFFFFFFFull);
  uint32_t state = inv_order ? inv_order[ford] : ford;
  meta[0] = 1; meta[2] = state;
  uint32_t L = 0;
  while (true) {
    unsigned long long key = pkey[state];
    if (key == kNoParent) break;
    uint32_t pord = (uint32_t)(key >> 32), pos = (uint32_t)(key & 0xFFFFFFFFull);
    uint32_t src = inv_order ? inv_order[pord] : pord;
    if (L >= cap || L >= n) { meta[3] = 1; break; }
    // only (source, position) here: the walk is one dependent load per hop; k_backtrace_fill fetches the arcs in parallel
    Tr tr; tr.ilabel = src; tr.olabel = pos; tr.weight = 0.0f; tr.nextstate = L;
    out_arcs[L] = tr;
    L++;
    state = src;
  }
  meta[1] = L;
}
// out_arcs[k] = arc `position` of state `source` (as left by k_backtrace_keys), pointing at output state k
__global__ void __launch_bounds__(kThreads)
k_backtrace_fill(const uint32_t* __restrict__ off, const Tr* __restrict__ arcs, Tr* __restrict__ out_arcs,
                 const uint32_t* __restrict__ meta) {
  const uint32_t L = meta[1];
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < L; k += gridDim.x * blockDim.x) {
    const Tr ref = out_arcs[k];
    Tr tr = arcs[off[ref.ilabel] + ref.olabel];
    tr.nextstate = k;  // previous output state
    out_arcs[k] = tr;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Order-faithful serial kernel: the reference loop, one thread.
// ---------------------------------------------------------------------------------------------------------------
struct SerialArgs {
  const uint32_t* off; const Tr* arcs; const float* fin; uint32_t n; uint32_t start;
  int kind;
  const uint32_t* order;      // kTopOrderQueue
  const uint32_t* scc;        // kSccQueue
  const uint8_t* scc_fifo;    // kSccQueue
  const uint32_t* scc_base;   // kSccQueue: ring buffer base per component (size nscc + 1)
  // scratch
  float* dist; uint32_t* pstate; uint32_t* ppos; uint8_t* enq;
  int32_t* slot;      // StateOrder: presence flag per state; TopOrder: state stored at order position (-1 none)
  uint32_t* stack;    // Lifo storage / Scc ring storage
  uint32_t* qhead; uint32_t* qlen;  // per component ring head / length (Trivial uses length 0/1)
  // shortest_distance only (queues may hold duplicates): radder, FIFO components as linked lists over a node pool
  // (stack[node] = state, pool_next[node] = next node, qhead/qtail = first/last node of the component)
  float* radder; uint32_t* pool_next; uint32_t* qtail; uint32_t pool_cap;
  // out
  Tr* out_arcs; uint32_t cap; uint32_t* meta; unsigned long long* counters;
};

template <bool kDup>
struct SerialQueue {
  const SerialArgs& a;
  // kDup: node pool bookkeeping + overflow flag
  uint32_t pool_used = 0, free_head = 0xFFFFFFFFu;
  bool overflow = false;
  // StateOrder / TopOrder
  uint32_t front = 0, back = 0; bool has_back = false;
  // Lifo
  uint32_t sp = 0;
  // Scc
  long long sfront = 0, sback = -1;
  __device__ explicit SerialQueue(const SerialArgs& args) : a(args) {}

  __device__ bool scc_q_empty(uint32_t c) const { return a.qlen[c] == 0; }
  __device__ void enqueue(uint32_t s) {
    switch (a.kind) {
      case kStateOrderQueue:
      case kTopOrderQueue: {  // state_order_queue.rs:16-31, top_order_queue.rs:45-56
        uint32_t o = a.kind == kTopOrderQueue ? a.order[s] : s;
        if (!has_back || front > back) { front = o; back = o; has_back = true; }
        else if (o > back) back = o;
        else if (o < front) front = o;
        a.slot[o] = a.kind == kTopOrderQueue ? (int32_t)s : 1;
        break;
      }
      case kLifoQueue:  // lifo_queue.rs
        if (kDup && sp >= a.pool_cap) { overflow = true; break; }
        a.stack[sp++] = s;
        break;
      default: {  // scc_queue.rs:34-45
        long long c = a.scc[s];
        if (sfront > sback) { sfront = c; sback = c; }
        else if (c > sback) sback = c;
        else if (c < sfront) sfront = c;
        if (kDup) {
          if (a.scc_fifo[c]) {  // FifoQueue with duplicates: append a pool node to the component's list
            uint32_t node;
            if (free_head != 0xFFFFFFFFu) { node = free_head; free_head = a.pool_next[node]; }
            else if (pool_used < a.pool_cap) node = pool_used++;
            else { overflow = true; break; }
            a.stack[node] = s; a.pool_next[node] = 0xFFFFFFFFu;
            if (a.qlen[c] == 0) a.qhead[c] = node; else a.pool_next[a.qtail[c]] = node;
            a.qtail[c] = node;
            a.qlen[c]++;
          } else {              // TrivialQueue: enqueue overwrites (trivial_queue.rs); the state lives in qhead
            a.qhead[c] = s;
            a.qlen[c] = 1;
          }
          break;
        }
        uint32_t size = a.scc_base[c + 1] - a.scc_base[c];
        if (a.scc_fifo[c]) {  // FifoQueue (ring sized to the component: a state is enqueued at most once at a time)
          a.stack[a.scc_base[c] + (a.qhead[c] + a.qlen[c]) % size] = s;
          a.qlen[c]++;
        } else {              // TrivialQueue: enqueue overwrites (trivial_queue.rs)
          a.stack[a.scc_base[c]] = s;
          a.qlen[c] = 1;
        }
      }
    }
  }
  __device__ bool dequeue(uint32_t* out) {
    switch (a.kind) {
      case kStateOrderQueue: {  // state_order_queue.rs:32-45
        if (!has_back || front > back) return false;
        *out = front;
        a.slot[front] = 0;
        while (front <= back && !a.slot[front]) front++;
        return true;
      }
      case kTopOrderQueue: {  // top_order_queue.rs:57-68
        if (!has_back || front > back) return false;
        int32_t head = a.slot[front];
        a.slot[front] = -1;
        while (front <= back && a.slot[front] < 0) front++;
        if (head < 0) return false;
        *out = (uint32_t)head;
        return true;
      }
      case kLifoQueue:
        if (sp == 0) return false;
        *out = a.stack[--sp];
        return true;
      default: {  // scc_queue.rs:46-62
        bool empty = sfront < sback ? false : (sfront > sback ? true : scc_q_empty((uint32_t)sfront));
        if (empty) return false;
        while (sfront <= sback && scc_q_empty((uint32_t)sfront)) sfront++;
        uint32_t c = (uint32_t)sfront;
        if (a.qlen[c] == 0) return false;
        if (kDup) {
          if (a.scc_fifo[c]) {
            uint32_t node = a.qhead[c];
            *out = a.stack[node];
            a.qhead[c] = a.pool_next[node];
            a.pool_next[node] = free_head; free_head = node;
          } else {
            *out = a.qhead[c];
          }
          a.qlen[c]--;
          return true;
        }
        uint32_t size = a.scc_base[c + 1] - a.scc_base[c];
        *out = a.stack[a.scc_base[c] + a.qhead[c]];
        if (a.scc_fifo[c]) a.qhead[c] = (a.qhead[c] + 1) % size;
        a.qlen[c]--;
        return true;
      }
    }
  }
};

__global__ void k_serial_sssp(SerialArgs a) {
  if (blockIdx.x || threadIdx.x) return;
  const float inf = w_zero();
  for (uint32_t s = 0; s < a.n; s++) { a.dist[s] = inf; a.pstate[s] = kNoState; a.ppos[s] = 0; a.enq[s] = 0; }
  if (a.kind == kStateOrderQueue) for (uint32_t s = 0; s < a.n; s++) a.slot[s] = 0;
  if (a.kind == kTopOrderQueue) for (uint32_t s = 0; s < a.n; s++) a.slot[s] = -1;
  SerialQueue<false> q(a);
  float f_distance = inf;
  bool has_f_parent = false;
  uint32_t f_parent = 0;
  unsigned long long relaxed = 0, dequeued = 0;
  a.dist[a.start] = 0.0f;
  a.enq[a.start] = 1;
  q.enqueue(a.start);
  uint32_t s;
  while (q.dequeue(&s)) {
    a.enq[s] = 0;
    float sd = a.dist[s];
    dequeued++;
    float rho = a.fin[s];
    if (rho != inf) {  // shortest_path.rs:214-220
      float plus = w_plus(f_distance, w_times(sd, rho));
      if (!w_approx_eq(f_distance, plus)) { f_distance = plus; f_parent = s; has_f_parent = true; }
    }
    uint32_t b = a.off[s], e = a.off[s + 1];
    relaxed += e - b;
    for (uint32_t k = b; k < e; k++) {  // shortest_path.rs:222-236
      Tr tr = a.arcs[k];
      float nd = a.dist[tr.nextstate];
      float p = w_plus(nd, w_times(sd, tr.weight));
      if (!w_approx_eq(nd, p)) {
        a.dist[tr.nextstate] = p;
        a.pstate[tr.nextstate] = s;
        a.ppos[tr.nextstate] = k - b;
        if (!a.enq[tr.nextstate]) { q.enqueue(tr.nextstate); a.enq[tr.nextstate] = 1; }
      }
    }
  }
  a.counters[0] = relaxed; a.counters[1] = dequeued;
  // backtrace (shortest_path.rs:241-282)
  a.meta[0] = has_f_parent ? 1u : 0u; a.meta[1] = 0; a.meta[2] = f_parent; a.meta[3] = 0;
  if (!has_f_parent) return;
  uint32_t state = f_parent, L = 0;
  while (a.pstate[state] != kNoState) {
    uint32_t src = a.pstate[state];
    if (L >= a.cap) { a.meta[3] = 1; break; }
    Tr tr = a.arcs[a.off[src] + a.ppos[state]];
    tr.nextstate = L;
    a.out_arcs[L++] = tr;
    state = src;
  }
  a.meta[1] = L;
}

// The loop of shortest_distance.rs:176-233, one thread, verbatim (tropical: adder == distance at all times, so
// only distance and radder are kept).  meta[3] = queue pool overflow (the host retries with a larger pool).
__global__ void k_serial_sdist(SerialArgs a, float delta) {
  if (blockIdx.x || threadIdx.x) return;
  const float inf = w_zero();
  for (uint32_t s = 0; s < a.n; s++) { a.dist[s] = inf; a.radder[s] = inf; a.enq[s] = 0; }
  if (a.kind == kStateOrderQueue) for (uint32_t s = 0; s < a.n; s++) a.slot[s] = 0;
  if (a.kind == kTopOrderQueue) for (uint32_t s = 0; s < a.n; s++) a.slot[s] = -1;
  SerialQueue<true> q(a);
  unsigned long long relaxed = 0, dequeued = 0;
  a.dist[a.start] = 0.0f;
  a.radder[a.start] = 0.0f;
  a.enq[a.start] = 1;
  q.enqueue(a.start);
  uint32_t s;
  while (!q.overflow && q.dequeue(&s)) {
    a.enq[s] = 0;
    const float r = a.radder[s];
    a.radder[s] = inf;
    dequeued++;
    const uint32_t b = a.off[s], e = a.off[s + 1];
    relaxed += e - b;
    for (uint32_t k = b; k < e; k++) {
      const Tr tr = a.arcs[k];
      const float nd = a.dist[tr.nextstate];
      const float w = w_times(r, tr.weight);
      const float p = w_plus(nd, w);
      if (!(fabsf(nd - p) <= delta)) {  // :217 — NaN (inf - inf) counts as "not equal", as in the reference
        a.dist[tr.nextstate] = p;
        a.radder[tr.nextstate] = w_plus(a.radder[tr.nextstate], w);
        if (!a.enq[s]) { q.enqueue(tr.nextstate); a.enq[tr.nextstate] = 1; }  // :224 tests `state` (sic)
      }
    }
  }
  a.counters[0] = relaxed; a.counters[1] = dequeued;
  a.meta[0] = a.meta[1] = a.meta[2] = 0; a.meta[3] = q.overflow ? 1u : 0u;
}

// ---------------------------------------------------------------------------------------------------------------
// Order-faithful PARALLEL path for DAGs processed in a topological order (used when the certificate fails, i.e. for
// weights with near-ties).  The reference dequeues states in `order`; when state t is dequeued every predecessor is
// final, so the value t ends up with is a pure function of its in-arcs: the sequential fold, in (order[src], arc
// position) order, of  d <- (d != min(d, c)) ? min(d, c) : d  with the approximate != of semiring.rs:159-168.
// That fold is replayed verbatim by one thread per state over a reverse CSR whose in-arc lists are sorted by
// (order[src], position) (one stable radix sort); states are scheduled level by level (Kahn), one grid barrier per
// level.  Exact for any weights; falls back to the serial kernel if the graph turns out to be cyclic.
__global__ void k_of_keys(const uint32_t* __restrict__ off, const Tr* __restrict__ arcs, uint32_t n,
                          const uint32_t* __restrict__ order, unsigned long long* __restrict__ keys,
                          uint32_t* __restrict__ vals, uint32_t* __restrict__ src_of, uint32_t* __restrict__ indeg) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const unsigned long long o = order ? order[s] : s;
  for (uint32_t e = off[s]; e < off[s + 1]; e++) {
    const uint32_t t = __ldg(&arcs[e].nextstate);
    keys[e] = ((unsigned long long)t << 32) | o;
    vals[e] = e;
    src_of[e] = s;
    atomicAdd(&indeg[t], 1u);
  }
}
__global__ void k_of_roots(const uint32_t* __restrict__ indeg, uint32_t n, uint32_t* __restrict__ topo,
                           uint32_t* __restrict__ cursor) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  warp_push(s < n && indeg[s] == 0, s, topo, cursor);
}
// ctl: [0] append cursor into topo (starts at #roots), [1] levels, [2] states processed, [3] grid barrier,
// [4..6] rotating per-level append counters
template <bool kAbs>
__global__ void __launch_bounds__(kThreads)
k_of_fold(const uint32_t* __restrict__ off, const Tr* __restrict__ arcs, uint32_t n, uint32_t source,
          const uint32_t* __restrict__ roff, const uint32_t* __restrict__ in_arc /* sorted arc ids */,
          const uint32_t* __restrict__ src_of, uint32_t* __restrict__ indeg, uint32_t* __restrict__ topo,
          float* __restrict__ dist, uint32_t* __restrict__ pstate, uint32_t* __restrict__ ppos,
          uint32_t* __restrict__ ctl, uint32_t n_roots, float delta) {
  unsigned int bar_epoch = 0;  // ctl[3] = arrival counter of the grid barrier (zero-initialised)
  __shared__ uint32_t s_q[kQueueCap];
  __shared__ uint32_t s_qn, s_gbase;
  const uint32_t gsize = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t lo = 0, hi = n_roots, levels = 0;
  while (lo < hi) {
    // ctl[4 + (levels + 1) % 3] counts the states appended during this level (= the next level), ctl[4 + (levels + 2) % 3]
    // is cleared for the level after it (nobody reads or writes it now), ctl[4 + levels % 3] is this level's own size
    uint32_t* const next_level = &ctl[4 + (levels + 1) % 3];
    if (gtid == 0) ctl[4 + (levels + 2) % 3] = 0;
    if (threadIdx.x == 0) s_qn = 0;
    __syncthreads();
    for (uint32_t i = lo + gtid; i < hi; i += gsize) {
      const uint32_t t = __ldcg(&topo[i]);
      // ---- the reference's relaxation of all in-arcs of t, in its processing order (shortest_path.rs:222-236)
      float d = (t == source) ? 0.0f : w_zero();
      uint32_t ps = kNoState, pp = 0;
      for (uint32_t k = roff[t]; k < roff[t + 1]; k++) {
        const uint32_t e = in_arc[k], src = src_of[e];
        const float c = w_times(__ldcg(&dist[src]), __ldg(&arcs[e].weight));
        const float p = w_plus(d, c);
        // kAbs: |inf - inf| is NaN, i.e. "not approx_equal": the update then rewrites inf with inf (harmless)
        const bool same = kAbs ? (fabsf(d - p) <= delta) : w_approx_eq(d, p);
        if (!same) { d = p; ps = src; pp = e - off[src]; }
      }
      dist[t] = d; pstate[t] = ps; ppos[t] = pp;
      // ---- Kahn: release the successors
      for (uint32_t e = off[t]; e < off[t + 1]; e++) {
        const uint32_t u = __ldg(&arcs[e].nextstate);
        if (atomicSub(&indeg[u], 1u) == 1u) {
          const uint32_t pos = atomicAdd(&s_qn, 1u);
          if (pos < kQueueCap) s_q[pos] = u;
          else { topo[atomicAdd(&ctl[0], 1u)] = u; atomicAdd(next_level, 1u); }
        }
      }
    }
    __syncthreads();
    const uint32_t qn = min(s_qn, kQueueCap);
    if (threadIdx.x == 0 && qn) { s_gbase = atomicAdd(&ctl[0], qn); atomicAdd(next_level, qn); }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < qn; i += kThreads) topo[s_gbase + i] = s_q[i];
    levels++;
    coop::grid_barrier(&ctl[3], bar_epoch);
    // The size of the next level comes from a counter nobody touches during that level (a CTA that races ahead appends
    // to the cursor ctl[0] — and to the NEXT rotating counter — while slower CTAs are still reading): every CTA sees
    // the same [lo, hi) and leaves the loop in the same iteration.
    lo = hi;
    hi = lo + __ldcg(next_level);
  }
  if (gtid == 0) { ctl[1] = levels; ctl[2] = hi; }
}
// Final-state fold in processing order + backtrace (one thread; the candidates are contiguous and sorted).
__global__ void k_of_finish(const uint32_t* __restrict__ off, const Tr* __restrict__ arcs, const float* __restrict__ fin,
                            const float* __restrict__ dist, const uint32_t* __restrict__ finals_sorted, uint32_t n_fin,
                            const uint32_t* __restrict__ pstate, const uint32_t* __restrict__ ppos,
                            Tr* __restrict__ out_arcs, uint32_t cap, uint32_t* __restrict__ meta) {
  if (blockIdx.x || threadIdx.x) return;
  float f = w_zero();
  bool has = false;
  uint32_t fp = 0;
  for (uint32_t i = 0; i < n_fin; i++) {  // shortest_path.rs:214-220
    const uint32_t s = finals_sorted[i];
    const float p = w_plus(f, w_times(dist[s], fin[s]));
    if (!w_approx_eq(f, p)) { f = p; fp = s; has = true; }
  }
  meta[0] = has ? 1u : 0u; meta[1] = 0; meta[2] = fp; meta[3] = 0;
  if (!has) return;
  uint32_t state = fp, L = 0;
  while (pstate[state] != kNoState) {
    const uint32_t src = pstate[state];
    if (L >= cap) { meta[3] = 1; break; }
    Tr tr = arcs[off[src] + ppos[state]];
    tr.nextstate = L;
    out_arcs[L++] = tr;
    state = src;
  }
  meta[1] = L;
}
__global__ void k_of_final_keys(const float* __restrict__ fin, uint32_t n, const uint32_t* __restrict__ order,
                                unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals,
                                uint32_t* __restrict__ count) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  const bool f = s < n && fin[s] != w_zero();
  uint32_t active = __activemask();
  uint32_t m = __ballot_sync(active, f);
  if (!m) return;
  uint32_t lane = threadIdx.x & 31, leader = __ffs(m) - 1, base = 0;
  if (lane == leader) base = atomicAdd(count, __popc(m));
  base = __shfl_sync(active, base, leader);
  if (f) {
    const uint32_t p = base + __popc(m & ((1u << lane) - 1u));
    keys[p] = order ? order[s] : s;
    vals[p] = s;
  }
}

// Rebuild the output FST on the host from the path arcs, replaying the reference's mutation sequence so the
// property word comes out identical (add_state; add_tr | set_final; ...; set_start; shortest_path_properties).
CsrFst build_path_fst(bool found, const std::vector<Tr>& path, float final_w) {
  CsrFst o;
  uint64_t p = props::kNull;
  if (found) {
    size_t L = path.size();
    o.offsets.assign(L + 2, 0);
    o.finals.assign(L + 1, w_zero());
    o.arcs.assign(path.begin(), path.end());
    for (size_t k = 0; k <= L; k++) {
      p = props::on_add_state(p);
      if (k == 0) {
        p = props::on_set_final(p, nullptr, &final_w);
        o.finals[0] = final_w;
        if (final_w == w_zero()) o.inf_finals.push_back(0);
      } else {
        p = props::on_add_tr(p, (StateId)k, path[k - 1], nullptr);
      }
      o.offsets[k + 1] = (uint32_t)k;
    }
    o.offsets[0] = 0;
    o.has_start = true; o.start = (StateId)L;
    p = props::on_set_start(p);
  }
  o.props = props::of_shortest_path(p, true) & props::kTrinary;
  return o;
}


using EventPairs = std::vector<std::pair<cudaEvent_t, cudaEvent_t>>;

// Exact minima on a TOP_SORTED machine by the in-order sweep (k_relax_sweep); dist as in run_relax_coop.
void run_relax_sweep(const DevFst& f, DevBuf<uint32_t>& dist, SsspStats& st, EventPairs& relax_events, cudaStream_t s) {
  const uint32_t n = f.num_states;
  DevBuf<uint32_t> ctl(s, 4);
  DevBuf<unsigned long long> out64(s, 2);
  dist.reserve_discard(n);
  k_fill_u32<<<blocks_for(n), kThreads, 0, s>>>(dist.p, kEncInf, n);
  const uint32_t zero_enc = enc_f32(0.0f);
  B200_CUDA(cudaMemcpyAsync(dist.p + f.start, &zero_enc, 4, cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemsetAsync(ctl.p, 0, 16, s));
  B200_CUDA(cudaMemsetAsync(out64.p, 0, 16, s));
  cudaEvent_t ea, eb;
  B200_CUDA(cudaEventCreate(&ea)); B200_CUDA(cudaEventCreate(&eb));
  B200_CUDA(cudaEventRecord(ea, s));
  int block = 512, sub = 128;  // states per block (= threads of the CTA) and per sub-block
  if (const char* e = std::getenv("B200_SWEEP_BLOCK")) block = std::atoi(e);
  if (const char* e = std::getenv("B200_SWEEP_SUB")) sub = std::atoi(e);
#define B200_SWEEP_CASE(B, S)                                                                                              \
  if (block == B && sub == S) {                                                                                            \
    B200_CUDA(cudaFuncSetAttribute((const void*)k_relax_sweep<B, S>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                   (int)sw_smem<B>()));                                                                    \
    k_relax_sweep<B, S><<<1, B, sw_smem<B>(), s>>>(f.offsets.p, f.arcs.p, n, dist.p, ctl.p, out64.p);                      \
    launched = true;                                                                                                       \
  }
  bool launched = false;
  B200_SWEEP_CASE(512, 16) B200_SWEEP_CASE(512, 32) B200_SWEEP_CASE(512, 64) B200_SWEEP_CASE(512, 128) B200_SWEEP_CASE(256, 32)
  if (!launched) { block = 512; sub = 128; B200_SWEEP_CASE(512, 128) }
#undef B200_SWEEP_CASE
  B200_CUDA(cudaEventRecord(eb, s));
  relax_events.emplace_back(ea, eb);
  st.relax_launches++; st.kernel_launches += 2;
  uint32_t hctl[4]; unsigned long long h64[2];
  B200_CUDA(cudaMemcpyAsync(hctl, ctl.p, 16, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaMemcpyAsync(h64, out64.p, 16, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  if (hctl[0]) throw FstError("shortest_path: the machine's property word says it is topologically sorted, but an arc points backwards");
  st.arcs_relaxed += h64[0]; st.states_settled += h64[1];
  st.sweep = true;
}

// Exact minima from f.start by frontier relaxation waves: one cooperative launch of k_relax_coop.
// dist receives the order-preserving integer images (kEncInf = unreached).
void run_relax_coop(const DevFst& f, DevBuf<uint32_t>& dist, SsspStats& st, EventPairs& relax_events, cudaStream_t s,
                    bool top_sorted) {
  const uint32_t n = f.num_states;
  DevBuf<uint32_t> stamp(s, n), fr_a(s, n), fr_b(s, n), cnt(s, 3), outw(s, 3);
  DevBuf<unsigned long long> out64(s, 3);
  dist.reserve_discard(n);
  k_fill_u32<<<blocks_for(n), kThreads, 0, s>>>(dist.p, kEncInf, n);
  B200_CUDA(cudaMemsetAsync(stamp.p, 0xFF, (size_t)n * 4, s));
  uint32_t zero_enc = enc_f32(0.0f), src = f.start;
  uint32_t init_cnt[3] = {1, 0, 0};
  B200_CUDA(cudaMemcpyAsync(dist.p + src, &zero_enc, 4, cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemcpyAsync(fr_a.p, &src, 4, cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemcpyAsync(cnt.p, init_cnt, 12, cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemsetAsync(outw.p, 0, 12, s));
  B200_CUDA(cudaMemsetAsync(out64.p, 0, 24, s));
  B200_CUDA(cudaStreamSynchronize(s));  // the sources of the small copies are host temporaries
  st.kernel_launches += 1;
  // lanes per frontier state: few lanes = more states in flight (the relaxation is latency-bound)
  // shape of the persistent relaxation kernel (measured on C4, ms of the kernel): 256 threads x 6 CTAs/SM with the
  // pre-test 1.25; 1024 x 2 1.19; without the pre-test 1.09 / 1.02; see BENCH_NOTES.md for the sweep
  int lanes = 8, threads = 1024, pre = 0, ilp = 1;
  if (const char* e = std::getenv("B200_RELAX_LANES")) lanes = std::atoi(e);
  if (const char* e = std::getenv("B200_RELAX_THREADS")) threads = std::atoi(e);
  if (const char* e = std::getenv("B200_RELAX_PRETEST")) pre = std::atoi(e);
  if (const char* e = std::getenv("B200_RELAX_ILP")) ilp = std::atoi(e);
  void* kern = nullptr;
#define B200_RELAX_PICK(G, T, P, S) if (!kern && lanes == G && threads == T && pre == P && ilp == S) kern = (void*)k_relax_coop<G, T, P != 0, S>;
  B200_RELAX_PICK(8, 1024, 0, 1) B200_RELAX_PICK(8, 1024, 0, 2) B200_RELAX_PICK(8, 1024, 0, 4)
  B200_RELAX_PICK(8, 1024, 1, 1) B200_RELAX_PICK(8, 1024, 1, 2)
  B200_RELAX_PICK(8, 512, 0, 1) B200_RELAX_PICK(8, 512, 0, 2) B200_RELAX_PICK(8, 512, 0, 4)
  B200_RELAX_PICK(8, 256, 0, 1) B200_RELAX_PICK(8, 256, 1, 1) B200_RELAX_PICK(8, 256, 0, 2)
  B200_RELAX_PICK(4, 1024, 0, 1) B200_RELAX_PICK(4, 1024, 0, 2) B200_RELAX_PICK(16, 1024, 0, 1) B200_RELAX_PICK(16, 1024, 0, 2)
  B200_RELAX_PICK(1, 256, 1, 1) B200_RELAX_PICK(2, 256, 1, 1) B200_RELAX_PICK(4, 256, 1, 1) B200_RELAX_PICK(16, 256, 1, 1)
#undef B200_RELAX_PICK
  if (!kern) { kern = (void*)k_relax_coop<8, 1024, false, 1>; threads = 1024; }
  int per_sm = 0;
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0));
  if (per_sm < 1) throw FstError("cooperative relaxation kernel does not fit on the device");
  if (const char* e = std::getenv("B200_RELAX_CTAS_PER_SM")) per_sm = std::min(per_sm, std::max(1, std::atoi(e)));
  int grid = sm_count() * per_sm;
  const uint32_t* a_off = f.offsets.p; const Tr* a_arcs = f.arcs.p;
  uint32_t nn = n;
  uint32_t* a_dist = dist.p; uint32_t* a_stamp = stamp.p; uint32_t* a_fa = fr_a.p; uint32_t* a_fb = fr_b.p;
  uint32_t* a_cnt = cnt.p; uint32_t* a_out = outw.p; unsigned long long* a_out64 = out64.p;
  // Visit budget: a state of a layered lattice is visited a handful of times; hundreds of visits per state mean a deep
  // DAG with skip arcs.  A TOP_SORTED machine (StateOrderQueue) is then handed to the in-order sweep.
  unsigned long long budget = top_sorted ? 4ull * n + 65536ull : ~0ull;
  if (const char* e = std::getenv("B200_RELAX_VISIT_BUDGET")) { if (top_sorted) budget = std::strtoull(e, nullptr, 10); }
  void* args[] = {&a_off, &a_arcs, &nn, &a_dist, &a_stamp, &a_fa, &a_fb, &a_cnt, &a_out, &a_out64, &budget};
  cudaEvent_t ea, eb;
  B200_CUDA(cudaEventCreate(&ea)); B200_CUDA(cudaEventCreate(&eb));
  B200_CUDA(cudaEventRecord(ea, s));
  B200_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(threads), args, 0, s));
  B200_CUDA(cudaEventRecord(eb, s));
  relax_events.emplace_back(ea, eb);
  st.relax_launches++; st.kernel_launches++;
  uint32_t hw[2]; unsigned long long h64[2];
  B200_CUDA(cudaMemcpyAsync(hw, outw.p, 8, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaMemcpyAsync(h64, out64.p, 16, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  st.waves = hw[0]; st.arcs_relaxed = h64[0]; st.states_settled = h64[1];
  if (hw[1] == 2) { run_relax_sweep(f, dist, st, relax_events, s); return; }  // over the visit budget
  if (hw[1]) throw FstError("shortest_path: relaxation did not converge (negative cycle?)");
}

// Order-faithful parallel fold (see k_of_fold).  Returns false when the graph turned out to be cyclic.
struct FoldOut {
  DevBuf<float> dist;
  DevBuf<uint32_t> pstate, ppos;
  explicit FoldOut(cudaStream_t s) : dist(s), pstate(s), ppos(s) {}
};
bool run_order_faithful_fold(const DevFst& f, const uint32_t* order_p, bool abs_form, float delta, FoldOut& out,
                             SsspStats& st, cudaStream_t s) {
  const uint32_t n = f.num_states, A = f.num_arcs;
  DevBuf<unsigned long long> k_in(s, A ? A : 1), k_out(s, A ? A : 1);
  DevBuf<uint32_t> v_in(s, A ? A : 1), in_arc(s, A ? A : 1), src_of(s, A ? A : 1), indeg(s, n), roff(s, (size_t)n + 1),
      topo(s, n), ctl(s, 8);
  DevBuf<uint8_t> tmp(s);
  out.dist.reserve_discard(n); out.pstate.reserve_discard(n); out.ppos.reserve_discard(n);
  B200_CUDA(cudaMemsetAsync(indeg.p, 0, (size_t)n * 4, s));
  B200_CUDA(cudaMemsetAsync(ctl.p, 0, 32, s));
  k_of_keys<<<blocks_for(n), kThreads, 0, s>>>(f.offsets.p, f.arcs.p, n, order_p, k_in.p, v_in.p, src_of.p, indeg.p);
  B200_CUDA(cudaMemcpyAsync(roff.p, indeg.p, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
  B200_CUDA(cudaMemsetAsync(roff.p + n, 0, 4, s));
  exclusive_sum_u32(roff.p, roff.p, (size_t)n + 1, tmp, s);
  sort_pairs_u64_u32(k_in.p, k_out.p, v_in.p, in_arc.p, A, 64, tmp, s);
  k_of_roots<<<blocks_for(n), kThreads, 0, s>>>(indeg.p, n, topo.p, ctl.p);
  uint32_t n_roots = read_u32(ctl.p, s);
  st.kernel_launches += 2;
  void* kern = abs_form ? (void*)k_of_fold<true> : (void*)k_of_fold<false>;
  int per_sm = 0;
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0));
  if (per_sm < 1) throw FstError("cooperative fold kernel does not fit on the device");
  int grid = sm_count() * per_sm;
  const uint32_t* a_off = f.offsets.p; const Tr* a_arcs = f.arcs.p; uint32_t nn = n, src0 = f.start;
  const uint32_t* a_roff = roff.p; const uint32_t* a_in = in_arc.p; const uint32_t* a_src = src_of.p;
  uint32_t* a_indeg = indeg.p; uint32_t* a_topo = topo.p; float* a_dist = out.dist.p;
  uint32_t* a_ps = out.pstate.p; uint32_t* a_pp = out.ppos.p; uint32_t* a_ctl = ctl.p;
  float a_delta = delta;
  void* args[] = {&a_off, &a_arcs, &nn, &src0, &a_roff, &a_in, &a_src, &a_indeg, &a_topo, &a_dist, &a_ps, &a_pp,
                  &a_ctl, &n_roots, &a_delta};
  B200_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kThreads), args, 0, s));
  st.kernel_launches++;
  uint32_t hctl[4];
  B200_CUDA(cudaMemcpyAsync(hctl, ctl.p, 16, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  if (hctl[2] != n) return false;  // not every state was scheduled: the graph has a cycle
  st.waves = hctl[1];
  st.arcs_relaxed = A; st.states_settled = n;
  return true;
}

// Device copies of the queue plan for the serial kernels.
struct SerialPlanBufs {
  DevBuf<uint32_t> scc, base, qhead, qlen, qtail;
  DevBuf<uint8_t> fifo;
  explicit SerialPlanBufs(cudaStream_t s) : scc(s), base(s), qhead(s), qlen(s), qtail(s), fifo(s) {}
};
void upload_scc_plan(const QueuePlan& plan, uint32_t n, SerialPlanBufs& b, SerialArgs& a, cudaStream_t s) {
  size_t nscc = plan.scc_is_fifo.size();
  std::vector<uint32_t> base(nscc + 1, 0);
  for (uint32_t c : plan.scc) base[c + 1]++;
  for (size_t c = 0; c < nscc; c++) base[c + 1] += base[c];
  b.scc.reserve_discard(n); b.base.reserve_discard(nscc + 1); b.fifo.reserve_discard(nscc);
  b.qhead.reserve_discard(nscc); b.qlen.reserve_discard(nscc); b.qtail.reserve_discard(nscc);
  B200_CUDA(cudaMemcpyAsync(b.scc.p, plan.scc.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemcpyAsync(b.base.p, base.data(), (nscc + 1) * 4, cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemcpyAsync(b.fifo.p, plan.scc_is_fifo.data(), nscc, cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemsetAsync(b.qhead.p, 0, nscc * 4, s));
  B200_CUDA(cudaMemsetAsync(b.qlen.p, 0, nscc * 4, s));
  B200_CUDA(cudaMemsetAsync(b.qtail.p, 0, nscc * 4, s));
  B200_CUDA(cudaStreamSynchronize(s));  // base is a host temporary
  a.scc = b.scc.p; a.scc_base = b.base.p; a.scc_fifo = b.fifo.p; a.qhead = b.qhead.p; a.qlen = b.qlen.p;
  a.qtail = b.qtail.p;
}

}  // namespace

// The serial replay walks the machine with one device thread (dependent accesses, ~10 M arcs per second); beyond this
// size a call would look like a hang, so it is refused with an explanation instead.  B200_SERIAL_SSSP_MAX_ARCS overrides.
void check_serial_size(const DevFst& f, const char* what) {
  size_t limit = (size_t)1 << 26;
  if (const char* e = std::getenv("B200_SERIAL_SSSP_MAX_ARCS")) limit = (size_t)std::strtoull(e, nullptr, 10);
  if (f.num_arcs > limit)
    throw FstError(std::string(what) + ": this machine is cyclic (or forces the order-faithful serial replay) and has " +
                   std::to_string(f.num_arcs) + " transitions; the serial replay is limited to " + std::to_string(limit) +
                   " (set B200_SERIAL_SSSP_MAX_ARCS to override)");
}

CsrFst shortest_path_device(const DevFst& f, const QueuePlan& plan, SsspStats* stats, cudaStream_t s,
                            bool force_serial) {
  SsspStats local;
  SsspStats& st = stats ? *stats : local;
  st = SsspStats();
  st.plan_host_ms = plan.host_ms;
  const uint32_t n = f.num_states;
  if (!f.has_start || n == 0) return build_path_fst(false, {}, 0.0f);  // shortest_path.rs:186-189
  DeviceExclusive excl(device_exclusive());  // the persistent kernels want every SM (device_common.cu)

  cudaEvent_t ev0, ev1;
  B200_CUDA(cudaEventCreate(&ev0)); B200_CUDA(cudaEventCreate(&ev1));
  B200_CUDA(cudaEventRecord(ev0, s));
  EventPairs relax_events;

  const uint32_t cap = n;  // a shortest path visits each state at most once
  DevBuf<Tr> out_arcs(s, cap);
  DevBuf<uint32_t> meta(s, 4);
  DevBuf<uint32_t> d_order(s);
  const uint32_t* order_p = nullptr;
  if (plan.kind == kTopOrderQueue && plan.d_order) {
    order_p = plan.d_order;  // computed on the device (dag_order.cu)
    st.order_device_ms = plan.device_ms; st.order_on_device = true;
  } else if (plan.kind == kTopOrderQueue) {
    if (plan.order.size() != n) throw FstError("shortest path: the TopOrderQueue order has not been computed");
    d_order.reserve_discard(n);
    B200_CUDA(cudaMemcpyAsync(d_order.p, plan.order.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
    order_p = d_order.p;
  }

  bool parallel_ok = !force_serial && (plan.kind == kStateOrderQueue || plan.kind == kTopOrderQueue);
  bool done = false;
  uint32_t hmeta[4] = {0, 0, 0, 0};

  if (parallel_ok) {
    DevBuf<uint32_t> dist(s), flags(s, 1), inv(s);
    DevBuf<unsigned long long> pkey(s, n), fkey(s, 1);
    k_fill_u64<<<blocks_for(n), kThreads, 0, s>>>(pkey.p, kNoParent, n);
    B200_CUDA(cudaMemsetAsync(fkey.p, 0xFF, 8, s));
    B200_CUDA(cudaMemsetAsync(flags.p, 0, 4, s));
    st.kernel_launches += 1;
    run_relax_coop(f, dist, st, relax_events, s, plan.kind == kStateOrderQueue);  // one cooperative launch runs every wave
    k_parents<false><<<blocks_for(n), kThreads, 0, s>>>(f.offsets.p, f.arcs.p, n, dist.p, order_p, pkey.p, flags.p,
                                                        kDelta);
    k_final_min<<<blocks_for(n), kThreads, 0, s>>>(f.finals.p, n, dist.p, order_p, fkey.p);
    k_final_check<<<blocks_for(n), kThreads, 0, s>>>(f.finals.p, n, dist.p, fkey.p, flags.p);
    st.kernel_launches += 3;
    uint32_t violations = read_u32(flags.p, s);
    if (violations == 0) {
      const uint32_t* inv_p = nullptr;
      if (order_p) {
        inv.reserve_discard(n);
        k_invert_order<<<blocks_for(n), kThreads, 0, s>>>(order_p, n, inv.p);
        inv_p = inv.p; st.kernel_launches++;
      }
      k_backtrace_keys<<<1, 32, 0, s>>>(f.offsets.p, f.arcs.p, n, pkey.p, fkey.p, inv_p, out_arcs.p, cap, meta.p);
      k_backtrace_fill<<<64, kThreads, 0, s>>>(f.offsets.p, f.arcs.p, out_arcs.p, meta.p);
      st.kernel_launches += 2;
      B200_CUDA(cudaMemcpyAsync(hmeta, meta.p, 16, cudaMemcpyDeviceToHost, s));
      B200_CUDA(cudaStreamSynchronize(s));
      st.path = 0;
      done = true;
    }
  }

  if (!done && parallel_ok) {  // certificate failed: order-faithful parallel fold over a sorted reverse CSR
    FoldOut fo(s);
    if (run_order_faithful_fold(f, order_p, false, kDelta, fo, st, s)) {  // every state scheduled: a DAG
      // final states in processing order
      DevBuf<unsigned long long> fk_in(s, n), fk_out(s, n);
      DevBuf<uint32_t> fv_in(s, n), fv_out(s, n), cnt(s, 1);
      DevBuf<uint8_t> tmp(s);
      B200_CUDA(cudaMemsetAsync(cnt.p, 0, 4, s));
      k_of_final_keys<<<blocks_for(n), kThreads, 0, s>>>(f.finals.p, n, order_p, fk_in.p, fv_in.p, cnt.p);
      uint32_t n_fin = read_u32(cnt.p, s);
      sort_pairs_u64_u32(fk_in.p, fk_out.p, fv_in.p, fv_out.p, n_fin, 32, tmp, s);
      k_of_finish<<<1, 32, 0, s>>>(f.offsets.p, f.arcs.p, f.finals.p, fo.dist.p, fv_out.p, n_fin, fo.pstate.p,
                                   fo.ppos.p, out_arcs.p, cap, meta.p);
      st.kernel_launches += 2;
      B200_CUDA(cudaMemcpyAsync(hmeta, meta.p, 16, cudaMemcpyDeviceToHost, s));
      B200_CUDA(cudaStreamSynchronize(s));
      st.path = 2;
      done = true;
    }
  }

  if (!done) {  // order-faithful serial replay
    check_serial_size(f, "shortest_path");
    st.path = 1;
    DevBuf<float> dist(s, n);
    DevBuf<uint32_t> pstate(s, n), ppos(s, n), stack(s, (size_t)n + 1);
    DevBuf<uint8_t> enq(s, n);
    DevBuf<int32_t> slot(s, n);
    DevBuf<unsigned long long> counters(s, 2);
    SerialPlanBufs pb(s);
    SerialArgs a{};
    a.off = f.offsets.p; a.arcs = f.arcs.p; a.fin = f.finals.p; a.n = n; a.start = f.start;
    a.kind = plan.kind; a.order = order_p;
    if (plan.kind == kSccQueue) upload_scc_plan(plan, n, pb, a, s);
    a.dist = dist.p; a.pstate = pstate.p; a.ppos = ppos.p; a.enq = enq.p; a.slot = slot.p; a.stack = stack.p;
    a.out_arcs = out_arcs.p; a.cap = cap; a.meta = meta.p; a.counters = counters.p;
    k_serial_sssp<<<1, 32, 0, s>>>(a);
    st.kernel_launches++;
    unsigned long long hc[2] = {0, 0};
    B200_CUDA(cudaMemcpyAsync(hmeta, meta.p, 16, cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaMemcpyAsync(hc, counters.p, 16, cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaStreamSynchronize(s));
    st.arcs_relaxed = hc[0]; st.states_settled = hc[1]; st.waves = 0;
  }

  if (hmeta[3]) throw FstError("shortest_path: parent chain does not terminate (zero/negative cycle)");
  std::vector<Tr> path(hmeta[1]);
  float final_w = 0.0f;
  if (hmeta[0]) {
    if (hmeta[1]) B200_CUDA(cudaMemcpyAsync(path.data(), out_arcs.p, (size_t)hmeta[1] * 16, cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaMemcpyAsync(&final_w, f.finals.p + hmeta[2], 4, cudaMemcpyDeviceToHost, s));
  }
  B200_CUDA(cudaEventRecord(ev1, s));
  B200_CUDA(cudaStreamSynchronize(s));
  B200_CUDA(cudaEventElapsedTime(&st.ms_device, ev0, ev1));
  for (auto& pr : relax_events) {
    float ms = 0;
    B200_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
    st.ms_relax_kernel += ms;
    cudaEventDestroy(pr.first); cudaEventDestroy(pr.second);
  }
  cudaEventDestroy(ev0); cudaEventDestroy(ev1);
  return build_path_fst(hmeta[0] != 0, path, final_w);
}

// Forward shortest distances with the semantics of shortest_distance.rs:153-237 (see the header of this function in
// algos.h).  Same three device paths as shortest_path_device, with approx_equal(.., delta) as the update test.
void shortest_distance_device(const DevFst& f, const QueuePlan& plan, float delta, DevBuf<float>& out,
                              SsspStats* stats, cudaStream_t s, bool force_serial) {
  SsspStats local;
  SsspStats& st = stats ? *stats : local;
  st = SsspStats();
  st.plan_host_ms = plan.host_ms;
  const uint32_t n = f.num_states;
  out.reserve_discard(n ? n : 1);
  if (!f.has_start || n == 0) return;
  DeviceExclusive excl(device_exclusive());  // the persistent kernels want every SM (device_common.cu)
  EventPairs relax_events;
  DevBuf<uint32_t> d_order(s);
  const uint32_t* order_p = nullptr;
  if (plan.kind == kTopOrderQueue && plan.d_order) {
    order_p = plan.d_order;  // computed on the device (dag_order.cu)
    st.order_device_ms = plan.device_ms; st.order_on_device = true;
  } else if (plan.kind == kTopOrderQueue) {
    if (plan.order.size() != n) throw FstError("shortest path: the TopOrderQueue order has not been computed");
    d_order.reserve_discard(n);
    B200_CUDA(cudaMemcpyAsync(d_order.p, plan.order.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
    order_p = d_order.p;
  }
  const bool parallel_ok = !force_serial && (plan.kind == kStateOrderQueue || plan.kind == kTopOrderQueue);
  bool done = false;
  if (parallel_ok) {
    // On a DAG processed in topological order every state is dequeued once, after all its predecessors, with
    // r = radder[state] = distance[state] (both fold the same accepted candidates), so the reference computes the
    // same per-state fold as single_shortest_path, only with |d - min(d, c)| <= delta as the "unchanged" test.
    // Exact minima + the certificate "no candidate in (m, m + delta]" therefore reproduce it (see k_parents).
    DevBuf<uint32_t> dist(s), flags(s, 1);
    B200_CUDA(cudaMemsetAsync(flags.p, 0, 4, s));
    run_relax_coop(f, dist, st, relax_events, s, plan.kind == kStateOrderQueue);
    k_parents<true><<<blocks_for(n), kThreads, 0, s>>>(f.offsets.p, f.arcs.p, n, dist.p, order_p, nullptr, flags.p,
                                                       delta);
    st.kernel_launches++;
    if (read_u32(flags.p, s) == 0) {
      k_decode_dist<<<blocks_for(n), kThreads, 0, s>>>(dist.p, n, out.p);
      st.kernel_launches++;
      st.path = 0;
      done = true;
    }
  }
  if (!done && parallel_ok) {
    FoldOut fo(s);
    if (run_order_faithful_fold(f, order_p, true, delta, fo, st, s)) {
      B200_CUDA(cudaMemcpyAsync(out.p, fo.dist.p, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
      st.path = 2;
      done = true;
    }
  }
  if (!done) {  // order-faithful serial replay with the queue discipline AutoQueue picks
    check_serial_size(f, "shortest_distance");
    st.path = 1;
    // Unlike single_shortest_path, shortest_distance re-enqueues a state that is already queued
    // (shortest_distance.rs:224 tests enqueued[state], not enqueued[nextstate]), so LIFO / FIFO queues hold
    // duplicates and have no a-priori bound: pools are sized generously and grown on overflow.
    size_t pool = 4 * ((size_t)n + f.num_arcs) + 1024;
    for (int attempt = 0;; attempt++) {
      DevBuf<float> radder(s, n);
      DevBuf<uint32_t> stack(s, pool), pool_next(s, pool), meta(s, 4);
      DevBuf<uint8_t> enq(s, n);
      DevBuf<int32_t> slot(s, n);
      DevBuf<unsigned long long> counters(s, 2);
      SerialPlanBufs pb(s);
      SerialArgs a{};
      a.off = f.offsets.p; a.arcs = f.arcs.p; a.fin = f.finals.p; a.n = n; a.start = f.start;
      a.kind = plan.kind; a.order = order_p;
      if (plan.kind == kSccQueue) upload_scc_plan(plan, n, pb, a, s);
      a.dist = out.p; a.radder = radder.p; a.enq = enq.p; a.slot = slot.p; a.stack = stack.p;
      a.pool_next = pool_next.p; a.pool_cap = (uint32_t)std::min<size_t>(pool, 0xFFFFFFF0u);
      a.meta = meta.p; a.counters = counters.p;
      k_serial_sdist<<<1, 32, 0, s>>>(a, delta);
      st.kernel_launches++;
      uint32_t hmeta[4];
      unsigned long long hc[2] = {0, 0};
      B200_CUDA(cudaMemcpyAsync(hmeta, meta.p, 16, cudaMemcpyDeviceToHost, s));
      B200_CUDA(cudaMemcpyAsync(hc, counters.p, 16, cudaMemcpyDeviceToHost, s));
      B200_CUDA(cudaStreamSynchronize(s));
      st.arcs_relaxed = hc[0]; st.states_settled = hc[1]; st.waves = 0;
      if (!hmeta[3]) break;
      if (attempt >= 4 || pool >= 0xFFFFFFF0u) throw FstError("shortest_distance: queue pool exhausted");
      pool *= 8;
    }
  }
  for (auto& pr : relax_events) {
    float ms = 0;
    B200_CUDA(cudaStreamSynchronize(s));
    B200_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
    st.ms_relax_kernel += ms;
    cudaEventDestroy(pr.first); cudaEventDestroy(pr.second);
  }
}

}  // namespace b200
