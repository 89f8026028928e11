// compose_ws.cu — the composition BFS as one persistent kernel of AUTONOMOUS WARP STREAMS (B200: 148 SMs x 24 warps).
//
// Same algorithm, same canonical numbering and the same reference citations as compose.cu / compose_coop.cu; what
// changes is how a BFS wave is synchronised.  compose_coop.cu needs four grid-wide exchanges per wave (items, arcs,
// barrier, new states); here a wave is
//
//   exchange  (new states, items, record region) of every CTA, epoch-tagged words: the only all-to-all of the wave
//   stream    every warp owns a contiguous run of the wave's items and works through it alone, with no CTA barrier:
//               match   32 items per tile against the sorted label lists (state records of the tile arrive in shared
//                       memory by one bulk-async copy, cp.async.bulk + mbarrier, issued one tile ahead)
//               reserve one atomicAdd on the arc cursor gives the warp a contiguous PROVISIONAL region for its arcs
//               emit    one lane per arc: gather the two component arcs, ONE 128-bit compare-and-swap on the state
//                       table (insert, or fetch key + id + first-emission word), write the arc
//   barrier   the only grid barrier of the wave: all first-emission minima have landed
//   rank      every warp looks at its OWN arcs: which of them is the first emission of a new tuple?  CTA-local ranks,
//             per-state setup of the next frontier into a record region reserved with one atomicAdd per CTA
//
// What makes this possible without changing the result:
//   * The reference numbers states by first emission under a FIFO BFS (lazy_fst.rs:226-269, state_table.rs:49-59).
//     Item runs are handed out in global warp order, so (warp << 20 | warp-local arc index) is order-isomorphic to the
//     canonical emission index of the wave: atomicMin on that key finds the first emitter without knowing how many
//     arcs the other warps emit.
//   * A state's canonical id is lo + (states found by lower CTAs) + CTA-local rank; it becomes known with the next
//     exchange and is published (table slot, tuple, final weight) by the lane that matches item 0 of the state.
//     Arcs that reach a tuple whose id is not published yet carry the slot index and are patched by the final pass.
//   * Arcs sit in provisional (warp, wave) runs; one flat pass after the BFS moves every run to its canonical place
//     (runs are ordered by (wave, warp) = canonical order, a prefix sum over the run lengths gives the targets) and
//     resolves the pending targets.  Each run is one contiguous bulk copy in, patch in shared memory, bulk copy out.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

#include "bulk_async.cuh"
#include "compose_match.cuh"
#include "coop_utils.cuh"

namespace b200 {
namespace {
using namespace composeimpl;
using namespace coop;

constexpr uint32_t kClaimed = 0xFFFFFFFEu;  // slot.id from the rank phase of the discovery wave until the id is published
constexpr uint32_t kKeyBits = kCoopThreads > 864 ? 19 : 20;  // emission key = global warp << kKeyBits | warp-local arc index
constexpr uint32_t kWarps = kCoopThreads / 32;
constexpr uint32_t kMaxProbes = 1u << 14;
constexpr uint32_t kNoSlot = 0xFFFFFFFFu;

// Per-state record of the current frontier (32 bytes: one tile of 33 records is one bulk copy).
struct __align__(32) StRec {
  uint32_t alo, ahi, blo, bhi;  // arc ranges of the two component states
  uint32_t item_loc;            // CTA-slice-local exclusive item offset | side bit
  uint32_t meta;                // alleps1 | noeps1<<1 | alleps2<<2 | noeps2<<3 | fs<<4 | searched side has sigma<<6 | owner CTA<<8
  uint32_t key_lo, key_hi;      // packed tuple
};
struct __align__(8) StCold { uint32_t slot; float fin; };  // read once, by the lane that matches item 0 of the state

struct WsParams {
  FstView a, b;
  const uint32_t* lab1; const uint32_t* lab2;
  int kind, side;
  unsigned long long* tuples; uint32_t states_cap;
  float* out_finals;
  uint2* st_first;          // per canonical state: (run index, warp-local index) of its first arc
  Tr* prov_arcs; uint32_t arcs_cap; unsigned long long* arc_cursor;
  uint32_t* prov_next;  // the next-state word of every provisional arc once more, dense: the rank phase and the final
                        // next-state pass read 4 bytes per arc instead of a 32-byte sector per two arcs
  uint32_t* run_src; uint32_t* run_cnt; uint32_t runs_cap;
  Slot* slots; uint32_t mask; uint32_t table_cap;
  StRec* st_rec; StCold* st_cold; uint32_t* st_cursor;   // record region of the current frontier
  uint4* recs; uint32_t* arc_loc; uint32_t items_cap;    // per item: x = first match, y = count|flags, z = record index, w = iterated arc or ~0
  unsigned long long* part_new; unsigned long long* part_sbase;  // epoch-tagged exchange words, one per CTA
  uint32_t* witems;         // items of the states each warp set up for the coming wave (read after the tagged words)
  uint32_t* ctl;            // [1], [7] overflow / error flags, [2] #states, [3] #runs, [4] #waves, [5] watchdog abort, [6] barrier
  uint32_t* wave_lo; uint32_t wave_cap;
  unsigned int* barrier;
  SigmaDev sig1, sig2;
  uint32_t n_starts;
  unsigned long long* stats;  // [0] states, [1] arcs iterated, [2] arcs emitted, [3] waves; CTA 0 / warp 0 timeline in ns:
                              // [4] match, [5] emit, [6] rank + setup, [7] exchange wait, [8] barrier wait, [9] run reservation
};

// Arcs per lane and emit round, arcs per lane and rank round.  Measured on C3 (kernel ms): two items per lane in the
// match phase 5.55 (spills: 80 registers at 768 threads) against 4.16 with one; two arcs per lane in the emit phase
// 4.16 against 4.01 with one; the match phase is software-pipelined over its tiles instead.
#ifndef B200_WS_EA
#define B200_WS_EA 1
#endif
#ifndef B200_WS_RA
#define B200_WS_RA 4
#endif
constexpr uint32_t kEA = B200_WS_EA, kRA = B200_WS_RA;
static_assert(kEA == 1, "the emit phase locates records with a 32-bit start mask: one arc per lane and round");
constexpr uint32_t kMT = 32;        // items per match tile (one per lane)
constexpr uint32_t kET = 32 * kEA;  // arcs per emit round

struct __align__(128) WarpSmem {
  StRec win[2][kMT + 2];      // match: state records of the current / next tile (kMT + 1 used)
  uint4 brec[2][kET];         // emit: match records of the current / next round
  uint32_t bloc[2][kET + 8];  // emit: their warp-local arc offsets (a 16-byte aligned window)
  unsigned long long mbar[4]; // [0,1] match windows, [2,3] emit windows
};
// the rank phase keeps the slots of the first emissions it found in the (then idle) match windows
constexpr uint32_t kFirstCap = (uint32_t)(sizeof(StRec) * 2 * (kMT + 2) / sizeof(uint32_t));

__device__ __forceinline__ void store_rec(StRec* p, const StRec& r) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(r.alo, r.ahi, r.blo, r.bhi);
  q[1] = make_uint4(r.item_loc, r.meta, r.key_lo, r.key_hi);
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Per-state setup of a product state (compose_fst_op.rs:199-219 match side, :420-449 final weight; filter flags as in
// compose_common.cuh).  Returns (#items | side bit); r.item_loc is left to the caller.  owner = global warp that sets
// the state up (its item offsets are relative to that warp's share of the wave).
__device__ __forceinline__ uint32_t setup_state_ws(const WsParams& P, unsigned long long key, uint32_t owner, StRec& r,
                                                   float& fin) {
  uint32_t fs, s1, s2;
  unpack_key(key, fs, s1, s2);
  const uint32_t alo = __ldg(&P.a.off[s1]), ahi = __ldg(&P.a.off[s1 + 1]);
  const uint32_t blo = __ldg(&P.b.off[s2]), bhi = __ldg(&P.b.off[s2 + 1]);
  const float f1 = __ldg(&P.a.fin[s1]), f2 = __ldg(&P.b.fin[s2]);
  const uint32_t ne1 = P.a.neps ? __ldg(&P.a.neps[s1]) : 0u, ne2 = P.b.neps ? __ldg(&P.b.neps[s2]) : 0u;
  const uint32_t d1 = ahi - alo, d2 = bhi - blo;
  bool mi = P.side == kMatchInput || (P.side == kMatchBoth && d1 <= d2);
  bool hs1 = false, hs2 = false;
  if (P.sig1.enabled) hs1 = dev_has_sigma<true>(P.sig1, P.a.arcs, alo, ahi);
  if (P.sig2.enabled) hs2 = dev_has_sigma<false>(P.sig2, P.b.arcs, blo, bhi);
  if (P.side == kMatchBoth && (hs1 || hs2)) {  // SigmaMatcher::priority = REQUIRE_PRIORITY (compose_fst_op.rs:199-219)
    if (hs1 && hs2) atomicOr(&P.ctl[7], (uint32_t)kErrBothRequire);
    mi = hs2;
  }
  const bool hs_searched = mi ? hs2 : hs1;
  const uint32_t fl = ((d1 == ne1 && f1 == w_zero()) ? 1u : 0u) | ((ne1 == 0) ? 2u : 0u) |
                      ((d2 == ne2 && f2 == w_zero()) ? 4u : 0u) | ((ne2 == 0) ? 8u : 0u) | (fs << 4) |
                      (hs_searched ? 64u : 0u);
  r.alo = alo; r.ahi = ahi; r.blo = blo; r.bhi = bhi;
  r.meta = fl | (owner << 8);
  r.key_lo = (uint32_t)key; r.key_hi = (uint32_t)(key >> 32);
  const float fw = w_times(f1, f2);
  fin = w_is_zero(fw) ? w_zero() : fw;
  return (1u + (mi ? d1 : d2)) | (mi ? kSideBit : 0u);
}

// Watchdog of the spin loops: a wait that lasts longer than kSpinLimitNs (or that sees another CTA's abort flag) raises
// kErrWatchdog and makes the whole grid leave the kernel instead of hanging the device.
constexpr unsigned long long kSpinLimitNs = 10ull * 1000 * 1000 * 1000;
__device__ __forceinline__ bool spin_check(const WsParams& P, unsigned long long t_start) {
  if (__ldcg(&P.ctl[5]) != 0u) return true;
  if (globaltimer_ns() - t_start > kSpinLimitNs) {
    atomicOr(&P.ctl[1], (uint32_t)kErrWatchdog);
    st_relaxed_u32(&P.ctl[5], 1u);
    return true;
  }
  return false;
}
__device__ __forceinline__ bool poll_tagged(const WsParams& P, const unsigned long long* p, uint32_t tag, uint32_t& out) {
  unsigned long long v;
  unsigned long long t_start = 0;
  for (uint32_t it = 1;; it++) {
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if ((uint32_t)(v >> 32) == tag) break;
    if ((it & 1023u) == 0) {
      if (!t_start) t_start = globaltimer_ns();
      if (spin_check(P, t_start)) return false;
    }
  }
  out = (uint32_t)v;
  return true;
}

// The exchange of a wave.  Every warp stores the item count of the states it set up, every CTA then publishes two
// epoch-tagged words (new states, base of its record region) after its writes; every CTA waits for all tagged words
// and builds, in shared memory, the exclusive prefix of the new states per CTA and of the items per WARP.  Passing the
// wait is a full grid barrier (see coop_utils.cuh).  false = the watchdog fired (uniform over the CTA).
__device__ __forceinline__ void publish2(const WsParams& P, uint32_t c, uint32_t tag, uint32_t n_new, uint32_t sbase) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long t = (unsigned long long)tag << 32;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(P.part_new + c), "l"(t | n_new) : "memory");
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(P.part_sbase + c), "l"(t | sbase) : "memory");
  }
}
constexpr uint32_t kWarpsPerThreadMax = 8;  // grid * kWarps <= 8 * kCoopThreads (checked by the host)
__device__ __forceinline__ bool wait2(const WsParams& P, uint32_t G, uint32_t tag, uint32_t* s_new, uint32_t* s_sbase,
                                      uint32_t* s_wpref, uint32_t* s_warp2, uint32_t* s_abort) {
  for (uint32_t i = threadIdx.x; i < G; i += kCoopThreads) {
    uint32_t a = 0, d = 0;
    if (!poll_tagged(P, P.part_new + i, tag, a) || !poll_tagged(P, P.part_sbase + i, tag, d)) *s_abort = 1u;
    s_new[i] = a; s_sbase[i] = d;
  }
  __threadfence();
  __syncthreads();
  if (*s_abort) return false;
  const uint32_t n_warps = G * kWarps;
  const uint32_t per = (n_warps + kCoopThreads - 1) / kCoopThreads;
  uint32_t wv[kWarpsPerThreadMax], sumw = 0;
#pragma unroll
  for (uint32_t k = 0; k < kWarpsPerThreadMax; k++) {
    const uint32_t idx = threadIdx.x * per + k;
    wv[k] = (k < per && idx < n_warps) ? __ldcg(&P.witems[idx]) : 0u;
    sumw += wv[k];
  }
  const uint32_t i0 = threadIdx.x * 2;  // G <= 2 * kCoopThreads (checked by the host)
  const uint32_t a0 = i0 < G ? s_new[i0] : 0u, a1 = i0 + 1 < G ? s_new[i0 + 1] : 0u;
  uint32_t ea, eb, ta, tb;
  cta_exclusive_scan2(a0 + a1, sumw, s_warp2, ea, eb, ta, tb);
  if (i0 < G) s_new[i0] = ea;
  if (i0 + 1 < G) s_new[i0 + 1] = ea + a0;
#pragma unroll
  for (uint32_t k = 0; k < kWarpsPerThreadMax; k++) {
    const uint32_t idx = threadIdx.x * per + k;
    if (k < per && idx < n_warps) s_wpref[idx] = eb;
    eb += wv[k];
  }
  if (threadIdx.x == 0) { s_new[G] = ta; s_wpref[n_warps] = tb; }
  __syncthreads();
  return true;
}
// arrival-counter grid barrier of coop_utils.cuh with the watchdog; false = aborted (uniform over the CTA)
__device__ __forceinline__ bool grid_barrier_wd(const WsParams& P, unsigned int& epoch, uint32_t* s_abort) {
  __syncthreads();
  epoch++;
  if (threadIdx.x == 0) {
    const unsigned int target = epoch * gridDim.x;
    __threadfence();
    atomicAdd(P.barrier, 1u);
    unsigned int v;
    unsigned long long t_start = 0;
    for (uint32_t it = 1;; it++) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(P.barrier) : "memory");
      if ((int)(v - target) >= 0) break;
      if ((it & 1023u) == 0) {
        if (!t_start) t_start = globaltimer_ns();
        if (spin_check(P, t_start)) { *s_abort = 1u; break; }
      }
    }
  }
  __syncthreads();
  return *s_abort == 0u;
}

__global__ void __launch_bounds__(kCoopThreads, 1)
k_compose_ws(WsParams P) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ uint32_t s_warp[2 * kWarps];
  __shared__ uint32_t s_misc[2];  // [0] record region of this CTA, [1] watchdog abort
  const uint32_t G = gridDim.x, c = blockIdx.x, tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
  WarpSmem& W = reinterpret_cast<WarpSmem*>(s_raw)[wid];
  uint32_t* const s_pref_new = reinterpret_cast<uint32_t*>(s_raw + kWarps * sizeof(WarpSmem));  // G + 1
  uint32_t* const s_sbase = s_pref_new + (G + 1);                                               // G
  uint32_t* const s_wpref = s_sbase + G;  // G * kWarps + 1: wave-global item offset of every producer warp's share
  const uint32_t gw = c * kWarps + wid;   // global warp index: item runs, arc runs and emission keys follow its order
  const uint32_t n_warps = G * kWarps;
  const uint32_t lt_mask = (1u << lane) - 1u;

  if (tid == 0) s_misc[1] = 0;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 4; k++) bulk::mbar_init(&W.mbar[k], 1);
    bulk::fence_mbar_init();
  }
  __syncwarp();
  uint32_t par = 0;  // phase parity of the four mbarriers

  unsigned int bar_epoch = 0;
  uint32_t tag = 1, lo = 0, run_base = 0, overflow = 0;
  unsigned long long n_states_exp = 0, n_items = 0, n_waves = 0;
  unsigned long long t_match = 0, t_emit = 0, t_rank = 0, t_wait = 0, t_bar = 0, t_res = 0;
  unsigned long long my_arcs_total = 0;

  // ---------------------------------------------------------------------- records of the initial frontier [0, n_starts)
  // CTA c takes a contiguous slice, every warp a contiguous part of it (item offsets are relative to the warp's part)
  {
    const uint32_t n0 = P.n_starts;
    const uint32_t sc0 = (n0 + G - 1) / G;
    const uint32_t s_begin = min(n0, c * sc0), s_end = min(n0, s_begin + sc0);
    const uint32_t pw = (s_end - s_begin + kWarps - 1) / kWarps;
    const uint32_t w_begin = min(s_end, s_begin + wid * pw), w_end = min(s_end, w_begin + pw);
    uint32_t run = 0;
    for (uint32_t i0 = w_begin; i0 < w_end; i0 += 32) {
      const uint32_t i = i0 + lane;
      uint32_t nit = 0;
      StRec r{};
      float fin = 0.f;
      if (i < w_end) nit = setup_state_ws(P, __ldcg(&P.tuples[i]), gw, r, fin);
      uint32_t inc = nit & ~kSideBit;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, inc, o); if ((int)lane >= o) inc += u; }
      if (i < w_end) {
        r.item_loc = (run + inc - (nit & ~kSideBit)) | (nit & kSideBit);
        store_rec(&P.st_rec[i], r);
        P.st_cold[i] = StCold{kNoSlot, fin};  // the start tuples already carry their ids (k_ws_init)
      }
      run += __shfl_sync(0xFFFFFFFFu, inc, 31);
    }
    if (lane == 0) P.witems[gw] = run;
    publish2(P, c, tag, s_end - s_begin, s_begin);
  }

  while (true) {
    const unsigned long long tw0 = globaltimer_ns();
    if (!wait2(P, G, tag, s_pref_new, s_sbase, s_wpref, s_warp, &s_misc[1])) break;  // watchdog
    bulk::fence_async_global();  // records written by other CTAs are read through bulk copies below
    const unsigned long long tw1 = globaltimer_ns();
    const uint32_t F = s_pref_new[G], T = s_wpref[n_warps];
    if (F == 0) break;  // uniform
    // ctl[1]: flags raised while matching / emitting (complete at the barrier); ctl[7]: flags raised while setting up the
    // next frontier (complete at the exchange).  Each word is only read where it is stable (ctl[1] after the barrier,
    // ctl[7] here), so every CTA takes the same decision.
    overflow = __ldcg(&P.ctl[7]);
    if ((unsigned long long)lo + F > P.states_cap || (unsigned long long)lo + F >= 0x7FFFFFFFull) overflow |= kOvStates;
    if (((unsigned long long)lo + F) * 2ull > P.table_cap) overflow |= kOvTable;
    if (T > P.items_cap) overflow |= kOvScratch;
    if (n_waves + 1 >= P.wave_cap) overflow |= kOvWaves;
    // every warp owns a contiguous run of wc items (at least one tile); warps beyond the last item sit the wave out
    const uint32_t wc = max(kMT, (T + n_warps - 1) / n_warps);
    const uint32_t w_act = (T + wc - 1) / wc;
    if ((unsigned long long)run_base + w_act > P.runs_cap) overflow |= kOvRuns;
    if (overflow) break;  // uniform
    if (c == 0 && tid == 0) {
      P.wave_lo[n_waves] = lo;
      *P.st_cursor = 0;  // the regions of this wave are all reserved; the next reservations follow this wave's barrier
    }

    const uint32_t wb = min(T, gw * wc), we = min(T, wb + wc);
    const uint32_t run_idx = run_base + gw;
    uint32_t w_active = 0, w_arcs = 0;  // active records / emitted arcs of this warp (lane-uniform)
    // ------------------------------------------------------------------ match
    if (wb < we) {
      // state containing my first item: producing warp from the shared prefix, then a 32-ary search (one probe per lane
      // and round) over the record region of its CTA
      uint32_t i_cur, p_cur;
      {
        const uint32_t p = smem_segment(s_wpref, n_warps, wb) / kWarps;
        const StRec* __restrict__ reg = P.st_rec + s_sbase[p];
        uint32_t l = 0, h = s_pref_new[p + 1] - s_pref_new[p];
        while (h - l > 1) {  // invariant: first item of record l <= wb < first item of record h (or h = end of the region)
          const uint32_t step = (h - l + 31u) >> 5, idx = l + lane * step;
          bool ok = false;
          if (idx < h) {
            const uint2 im = __ldcg(reinterpret_cast<const uint2*>(&reg[idx].item_loc));
            ok = s_wpref[im.y >> 8] + (im.x & ~kSideBit) <= wb;
          }
          const uint32_t n_ok = __popc(__ballot_sync(0xFFFFFFFFu, ok));  // ok lanes form a prefix, lane 0 is always ok
          l += (n_ok - 1u) * step;
          h = min(h, l + step);
        }
        i_cur = s_pref_new[p] + l; p_cur = p;
      }
      // window of a tile = records of the states i .. i + kMT (frontier order), fetched by lane 0 as one bulk copy per
      // producer region it touches, one tile ahead of the matching
      auto issue_awin = [&](uint32_t buf, uint32_t i_base, uint32_t p_start) {
        if (lane == 0) {
          uint32_t remaining = min(kMT + 1u, F - i_base), i = i_base, p = p_start, dst = 0;
          bulk::mbar_expect_tx(&W.mbar[buf], remaining * (uint32_t)sizeof(StRec));
          while (remaining) {
            while (s_pref_new[p + 1] <= i) p++;
            const uint32_t n = min(remaining, s_pref_new[p + 1] - i);
            bulk::g2s(&W.win[buf][dst], P.st_rec + s_sbase[p] + (i - s_pref_new[p]), n * (uint32_t)sizeof(StRec),
                      &W.mbar[buf]);
            dst += n; i += n; remaining -= n;
          }
        }
      };
      // Software pipeline over the tiles of the run.  Stage A of tile t + 1 (unpack its window, locate the lane's item,
      // request the item's label) runs before stage B of tile t (search, filter, scans, stores), so the label request
      // of one tile and the label-window requests of the other are in flight together, and the window of tile t + 2 is
      // on its way as a bulk copy.  A lane carries the few words of its item from stage A to stage B in registers.
      struct Item {
        uint32_t se_lo, se_hi, it_idx, flags, sidx;  // flags = record flags | side bit
        Label label;
        uint32_t id, slot, fin_bits, key_lo, key_hi;  // item 0 only: canonical facts of the state to publish
        bool valid;
      };
      uint32_t i_next = 0;
      auto stage_a = [&](uint32_t buf, uint32_t t0, uint32_t i_base) -> Item {
        bulk::mbar_wait(&W.mbar[buf], (par >> buf) & 1u);
        par ^= 1u << buf;
        const StRec* __restrict__ win = W.win[buf];
        const uint32_t n_win = min(kMT + 1u, F - i_base);
        // lane x holds the first item of state i_base + x (and every lane that of state i_base + 32).  A state starts at a
        // lane of the tile iff its first item falls inside the tile: one warp-wide OR of those bits, and the state of a
        // lane's item is the number of starts at or before the lane (no per-lane binary search).
        const uint32_t sv = lane < n_win ? s_wpref[win[lane].meta >> 8] + (win[lane].item_loc & ~kSideBit) : T;
        const uint32_t sv32 = 32u < n_win ? s_wpref[win[32].meta >> 8] + (win[32].item_loc & ~kSideBit) : T;
        const uint32_t starts = __reduce_or_sync(0xFFFFFFFFu, (lane >= 1 && sv > t0 && sv - t0 < 32u) ? 1u << (sv - t0) : 0u);
        Item it;
        it.se_lo = 0; it.se_hi = 0; it.it_idx = 0xFFFFFFFFu; it.flags = 0; it.sidx = 0; it.label = kNoLabel;
        it.id = 0; it.slot = kNoSlot; it.fin_bits = 0; it.key_lo = 0; it.key_hi = 0;
        const uint32_t t = t0 + lane;
        it.valid = t < we;
        const uint32_t k = __popc(starts & (lt_mask | (1u << lane)));
        const uint32_t seg_k = __shfl_sync(0xFFFFFFFFu, sv, k & 31u);
        const uint32_t seg_k1s = __shfl_sync(0xFFFFFFFFu, sv, (k + 1u) & 31u);
        const uint32_t seg_k1 = k + 1u < 32u ? seg_k1s : sv32;
        // the lane on the tile's last item knows which state holds the first item of the next tile
        const uint32_t next_note = it.valid ? i_base + k + (seg_k1 <= t + 1 ? 1u : 0u) : 0u;
        i_next = __shfl_sync(0xFFFFFFFFu, next_note, 31);
        if (t0 + kMT < we) issue_awin(buf ^ 1u, i_next, (win[min(i_next - i_base, n_win - 1u)].meta >> 8) / kWarps);
        if (it.valid) {
          const uint4 so = *reinterpret_cast<const uint4*>(&win[k]);
          const uint4 sm = *(reinterpret_cast<const uint4*>(&win[k]) + 1);  // item_loc, meta, key
          const uint32_t i = i_base + k, j = t - seg_k;
          const bool match_input = (sm.x & kSideBit) != 0;
          const uint32_t owner = (sm.y >> 8) / kWarps;
          it.sidx = s_sbase[owner] + (i - s_pref_new[owner]);
          it.flags = (sm.y & 0xFFu) | (sm.x & kSideBit);
          it.se_lo = match_input ? so.z : so.x; it.se_hi = match_input ? so.w : so.y;
          if (j != 0) {  // absolute index of the iterated arc; all ones = implicit epsilon loop
            it.it_idx = (match_input ? so.x : so.z) + j - 1;
            it.label = __ldg(&(match_input ? P.lab1 : P.lab2)[it.it_idx]);
          } else {
            const uint2 cold = __ldcg(reinterpret_cast<const uint2*>(&P.st_cold[it.sidx]));
            it.id = lo + i; it.slot = cold.x; it.fin_bits = cold.y; it.key_lo = sm.z; it.key_hi = sm.w;
          }
        }
        __syncwarp();  // the other window buffer is rewritten by the next stage A
        return it;
      };
      issue_awin(0, i_cur, p_cur);
      Item cur = stage_a(0, wb, i_cur);
      uint32_t buf = 1;
      for (uint32_t t0 = wb; t0 < we; t0 += kMT) {
        Item nxt;
        nxt.valid = false;
        if (t0 + kMT < we) { nxt = stage_a(buf, t0 + kMT, i_next); buf ^= 1u; }
        // ---- stage B
        uint32_t cnt_out = 0;
        uint4 rec = make_uint4(0, 0, 0, 0);
        if (cur.valid) {
          const uint32_t fl = cur.flags;
          const bool match_input = (fl & kSideBit) != 0;
          const uint32_t fs = (fl >> 4) & 3u;
          const bool hs_searched = (fl & 64) != 0;
          FsFlags ff;
          ff.alleps1 = fl & 1; ff.noeps1 = fl & 2; ff.alleps2 = fl & 4; ff.noeps2 = fl & 8;
          const uint32_t* __restrict__ se_lab = match_input ? P.lab2 : P.lab1;
          const Label lab = cur.label;
          const bool has_loop = (lab == kEps);
          const Label key = (lab == kNoLabel) ? kEps : lab;
          uint32_t pos, end;
          match_range_eq(se_lab, cur.se_lo, cur.se_hi, key, has_loop, pos, end);
          uint32_t cnt = end - pos;
          // filter_tr sees (arc1.olabel, arc2.ilabel): the iterated arc's label on its own side, the match on the other
          const uint32_t fs_loop = !has_loop ? kNoFs
                                   : filter_eval(P.kind, fs, ff, match_input ? lab : kNoLabel, match_input ? kNoLabel : lab);
          const uint32_t fs_real = filter_eval(P.kind, fs, ff, match_input ? lab : key, match_input ? key : lab);
          bool sigma_mode = false;
          if (P.sig1.enabled | P.sig2.enabled) {
            const SigmaDev& sg = match_input ? P.sig2 : P.sig1;  // matcher of the searched side
            if (sg.enabled) {  // IteratorSigmaMatcher::new (sigma_matcher.rs:196-246)
              if (lab == sg.label && sg.label != kNoLabel) atomicOr(&P.ctl[1], (uint32_t)kErrBadSigmaLabel);
              if (!has_loop && cnt == 0 && hs_searched && lab != kEps && lab != kNoLabel && dev_sigma_allowed(sg, lab)) {
                pos = lower_bound_lab(se_lab, cur.se_lo, cur.se_hi, sg.label);
                end = run_end_lab(se_lab, pos, cur.se_hi, sg.label);
                cnt = end - pos;
                sigma_mode = true;  // the filter sees the relabelled arc: (label, label), i.e. fs_real as computed
              }
            }
          }
          const bool loop_ok = has_loop && fs_loop != kNoFs;
          const bool real_ok = fs_real != kNoFs && cnt > 0;
          cnt_out = (loop_ok ? 1u : 0u) + (real_ok ? cnt : 0u);
          // y: emitted real matches (25 bits) | sigma | loop_ok | fs_loop(2) | fs_real(2) | match_input
          rec = make_uint4(pos, (real_ok ? (cnt & 0x01FFFFFFu) : 0u) | (sigma_mode ? 1u << 25 : 0u) |
                                    (loop_ok ? 1u << 26 : 0u) | ((fs_loop & 3u) << 27) |
                                    ((fs_real & 3u) << 29) | (match_input ? 1u << 31 : 0u), cur.sidx, cur.it_idx);
        }
        // warp scans: active records by ballot, arcs by shuffles
        const bool act = cnt_out != 0;
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, act);
        uint32_t inc = cnt_out;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, inc, o); if ((int)lane >= o) inc += u; }
        const uint32_t first_arc = w_arcs + inc - cnt_out;
        if (cur.valid && cur.it_idx == 0xFFFFFFFFu) {
          // this lane is the only one in the grid that sees item 0 of the state: publish its canonical id (table slot),
          // tuple and final weight, and where its arcs begin
          if (cur.slot != kNoSlot) st_relaxed_u32(&P.slots[cur.slot].id, cur.id);
          P.tuples[cur.id] = (unsigned long long)cur.key_lo | ((unsigned long long)cur.key_hi << 32);
          P.out_finals[cur.id] = __uint_as_float(cur.fin_bits);
          P.st_first[cur.id] = make_uint2(run_idx, first_arc);
        }
        if (act) {
          const uint32_t r = wb + w_active + __popc(bal & lt_mask);  // in-place compaction inside the warp's own item run
          P.recs[r] = rec;
          P.arc_loc[r] = first_arc;
        }
        w_active += __popc(bal);
        w_arcs += __shfl_sync(0xFFFFFFFFu, inc, 31);
        cur = nxt;
      }
    }
    const unsigned long long tm1 = globaltimer_ns();

    // ------------------------------------------------------------------ reserve the run, emit
    // the first emit window is requested before the run is reserved: the bulk copy flies during the atomic's round trip
    auto issue_bwin = [&](uint32_t buf, uint32_t cur) {
      if (lane == 0) {
        const uint32_t n = min(kET, w_active - cur);
        bulk::mbar_expect_tx(&W.mbar[2 + buf], n * 16u + (kET + 8u) * 4u);
        bulk::g2s(W.brec[buf], P.recs + wb + cur, n * 16u, &W.mbar[2 + buf]);
        bulk::g2s(W.bloc[buf], P.arc_loc + ((wb + cur) & ~3u), (kET + 8u) * 4u, &W.mbar[2 + buf]);
      }
    };
    if (w_arcs) {
      bulk::fence_async_global();  // the records written above come back through bulk copies
      __syncwarp();
      issue_bwin(0, 0);
    }
    uint32_t prov = 0;
    bool emit_ok = w_arcs != 0;
    if (gw < w_act) {
      if (lane == 0) {
        if (w_arcs >= (1u << kKeyBits)) { atomicOr(&P.ctl[1], (uint32_t)kOvChunk); emit_ok = false; }
        if (emit_ok) {
          const unsigned long long old = atomicAdd(P.arc_cursor, (unsigned long long)w_arcs);
          if (old + w_arcs > P.arcs_cap) { atomicOr(&P.ctl[1], (uint32_t)kOvArcs); emit_ok = false; }
          prov = (uint32_t)old;
        }
        P.run_src[run_idx] = prov;
        P.run_cnt[run_idx] = w_arcs;
      }
      prov = __shfl_sync(0xFFFFFFFFu, prov, 0);
      emit_ok = __shfl_sync(0xFFFFFFFFu, emit_ok ? 1u : 0u, 0) != 0;
    }
    const unsigned long long tm2 = globaltimer_ns();
    const uint32_t ekey = gw << kKeyBits;
    Tr* __restrict__ run_arcs = P.prov_arcs + prov;
    uint32_t* __restrict__ run_next = P.prov_next + prov;
    if (emit_ok) {
      // Tried and dropped: keeping the warp's records (and the next-state words for the rank phase) in shared memory so
      // that the emit phase needs no fence / copy / wait.  With 160 records + 256 words per warp (201 KB per CTA) the
      // kernel went from 4.01 to 4.68 ms — the L1 cache shrinks to what the carve-out leaves and the label windows of
      // the match phase stop hitting it; with 64 + 128 (144 KB) it was a wash (3.89 vs 3.86 ms).
      uint32_t buf = 0, cursor = 0;  // first record (warp-local) that can contain the round's first arc
      for (uint32_t e0 = 0; e0 < w_arcs; e0 += kET) {
        // lane x holds the first arc of record cursor + x; which record an arc belongs to follows from one warp-wide OR
        // of the record starts inside the round (as in the match phase)
        bulk::mbar_wait(&W.mbar[2 + buf], (par >> (2 + buf)) & 1u);
        par ^= 4u << buf;
        const uint32_t off4 = (wb + cursor) & 3u;
        const uint32_t lv = cursor + lane < w_active ? W.bloc[buf][off4 + lane] : w_arcs;
        const uint32_t lv32 = cursor + 32u < w_active ? W.bloc[buf][off4 + 32u] : w_arcs;
        const uint4* __restrict__ rec_win = W.brec[buf];
        const uint32_t starts = __reduce_or_sync(0xFFFFFFFFu, (lane >= 1 && lv > e0 && lv - e0 < 32u) ? 1u << (lv - e0) : 0u);
        uint32_t el[kEA], k[kEA], seg_k[kEA];
        bool valid[kEA];
        el[0] = e0 + lane;
        valid[0] = el[0] < w_arcs;
        k[0] = __popc(starts & (lt_mask | (1u << lane)));
        seg_k[0] = __shfl_sync(0xFFFFFFFFu, lv, k[0] & 31u);
        const uint32_t seg_k1s = __shfl_sync(0xFFFFFFFFu, lv, (k[0] + 1u) & 31u);
        const uint32_t seg_k1 = k[0] + 1u < 32u ? seg_k1s : lv32;
        // the lane on the round's last arc knows which record holds the first arc of the next round
        const uint32_t next_note = valid[0] ? cursor + k[0] + (seg_k1 <= el[0] + 1 ? 1u : 0u) : 0u;
        const uint32_t cursor_next = __shfl_sync(0xFFFFFFFFu, next_note, 31);
        if (e0 + kET < w_arcs) issue_bwin(buf ^ 1u, cursor_next);
        // step 1: gather the two component arcs of every arc of the lane
        Tr it[kEA], cand[kEA];
        uint32_t fsn[kEA];
        bool mi[kEA];
#pragma unroll
        for (uint32_t q = 0; q < kEA; q++) {
          mi[q] = false; fsn[q] = 0;
          if (valid[q]) {
            const uint4 rec = rec_win[k[q]];
            const uint32_t kk = el[q] - seg_k[q];
            const bool loop_ok = (rec.y >> 26) & 1u;
            const bool match_input = rec.y >> 31;
            const bool it_is_loop = rec.w == 0xFFFFFFFFu, cand_is_loop = loop_ok && kk == 0;
            const Tr* __restrict__ it_arcs = match_input ? P.a.arcs : P.b.arcs;
            const Tr* __restrict__ cd_arcs = match_input ? P.b.arcs : P.a.arcs;
            uint32_t s1 = 0, s2 = 0;
            if (it_is_loop || cand_is_loop) {
              const uint2 kq = __ldcg(reinterpret_cast<const uint2*>(&P.st_rec[rec.z].key_lo));
              uint32_t fs;
              unpack_key((unsigned long long)kq.x | ((unsigned long long)kq.y << 32), fs, s1, s2);
            }
            // implicit epsilon loops (matcher.rs: eps_loop): (0, NO_LABEL) / (NO_LABEL, 0) staying in the same state
            it[q] = match_input ? Tr{kEps, kNoLabel, 0.0f, s1} : Tr{kNoLabel, kEps, 0.0f, s2};
            cand[q] = match_input ? Tr{kNoLabel, kEps, 0.0f, s2} : Tr{kEps, kNoLabel, 0.0f, s1};
            if (!it_is_loop) it[q] = load_tr(&it_arcs[rec.w]);
            if (!cand_is_loop) cand[q] = load_tr(&cd_arcs[rec.x + kk - (loop_ok ? 1u : 0u)]);
            fsn[q] = (rec.y >> 27) & 3u;
            if (!cand_is_loop) {
              fsn[q] = (rec.y >> 29) & 3u;
              if ((rec.y >> 25) & 1u) {  // sigma match: relabel (value_openfst, sigma_matcher.rs:249-276)
                const SigmaDev& sg = match_input ? P.sig2 : P.sig1;
                const Label l = match_input ? it[q].olabel : it[q].ilabel;
                if (sg.rewrite_both) { if (cand[q].ilabel == sg.label) cand[q].ilabel = l; if (cand[q].olabel == sg.label) cand[q].olabel = l; }
                else if (match_input) cand[q].ilabel = l;
                else cand[q].olabel = l;
              }
            }
            mi[q] = match_input;
          }
        }
        // step 2: the first probe of every arc.  One 128-bit compare-and-swap either inserts {key, unassigned, my emission
        // key} or returns the slot's key, id and first-emission word.
        Tr out[kEA];
        unsigned long long key[kEA];
        uint32_t h[kEA];
        bulk::U128 old[kEA];
#pragma unroll
        for (uint32_t q = 0; q < kEA; q++) {
          if (valid[q]) {
            out[q].ilabel = mi[q] ? it[q].ilabel : cand[q].ilabel;   // arc1 = the fst1 arc, arc2 = the fst2 arc
            out[q].olabel = mi[q] ? cand[q].olabel : it[q].olabel;
            out[q].weight = w_times(it[q].weight, cand[q].weight);
            key[q] = pack_key(fsn[q], mi[q] ? it[q].nextstate : cand[q].nextstate, mi[q] ? cand[q].nextstate : it[q].nextstate);
            h[q] = hash_key(key[q]) & P.mask;
            old[q] = bulk::cas128(&P.slots[h[q]], bulk::U128{kEmptyKey, ~0ull},
                                  bulk::U128{key[q], ((unsigned long long)(ekey | el[q]) << 32) | kUnassigned});
          }
        }
        // step 3: resolve (further probes on a collision), write the arc
#pragma unroll
        for (uint32_t q = 0; q < kEA; q++) {
          if (valid[q]) {
            const uint32_t e = ekey | el[q];
            out[q].nextstate = 0;
            for (uint32_t probes = 0;; probes++) {
              if (old[q].lo == kEmptyKey) { out[q].nextstate = kPendingBit | h[q]; break; }  // fresh insert, emin = e
              if (old[q].lo == key[q]) {
                const uint32_t id = (uint32_t)old[q].hi;
                if (id == kUnassigned) { atomicMin(&P.slots[h[q]].emin, e); out[q].nextstate = kPendingBit | h[q]; }
                else if (id == kClaimed) out[q].nextstate = kPendingBit | h[q];  // found in the previous wave, id not published yet
                else out[q].nextstate = id;
                break;
              }
              if (probes > kMaxProbes) { atomicOr(&P.ctl[1], (uint32_t)kOvTable); break; }
              h[q] = (h[q] + 1) & P.mask;
              old[q] = bulk::cas128(&P.slots[h[q]], bulk::U128{kEmptyKey, ~0ull},
                                    bulk::U128{key[q], ((unsigned long long)e << 32) | kUnassigned});
            }
            store_tr(&run_arcs[el[q]], out[q]);
            run_next[el[q]] = out[q].nextstate;
          }
        }
        cursor = cursor_next;
        buf ^= 1u;
        __syncwarp();
      }
      my_arcs_total += w_arcs;
    }
    const unsigned long long te1 = globaltimer_ns();
    if (!grid_barrier_wd(P, bar_epoch, &s_misc[1])) break;  // watchdog
    const unsigned long long tb1 = globaltimer_ns();
    overflow = __ldcg(&P.ctl[1]);
    if (overflow) break;  // uniform: every flag of this wave was raised before the barrier

    // ------------------------------------------------------------------ rank first emissions, set up the next frontier
    // Every warp looks at its own arcs (kRA per lane and round, loads in flight together) and lists the table slots of
    // the tuples it emitted first, in emission order; the list lives in the idle match windows.
    uint32_t* const firsts = reinterpret_cast<uint32_t*>(&W.win[0][0]);
    auto list_firsts = [&](uint32_t e_begin, uint32_t e_end, bool fill) -> uint32_t {
      uint32_t n = 0;
      for (uint32_t e0 = e_begin; e0 < e_end; e0 += 32u * kRA) {
        uint32_t ns[kRA];
        uint4 sv[kRA];
#pragma unroll
        for (uint32_t q = 0; q < kRA; q++) {
          const uint32_t el = e0 + 32u * q + lane;
          ns[q] = el < e_end ? __ldcg(&run_next[el]) : 0u;
        }
#pragma unroll
        for (uint32_t q = 0; q < kRA; q++)
          sv[q] = (ns[q] & kPendingBit) ? ld_volatile_u4(&P.slots[ns[q] & ~kPendingBit]) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (uint32_t q = 0; q < kRA; q++) {
          const uint32_t el = e0 + 32u * q + lane;
          const bool first = (ns[q] & kPendingBit) && sv[q].z == kUnassigned && sv[q].w == (ekey | el);
          const uint32_t bal = __ballot_sync(0xFFFFFFFFu, first);
          if (first && fill) firsts[n + __popc(bal & lt_mask)] = ns[q] & ~kPendingBit;
          n += __popc(bal);
        }
      }
      return n;
    };
    // Tried and dropped: resolving the pending targets of the warp's PREVIOUS run here (their ids were published while
    // this wave was matched, the slots are still in L2) instead of in the final pass.  The final pass got 0.22 ms
    // shorter, this phase 0.24 ms longer (C3): a wash, and one more pass over the arcs inside the wave loop.
    const bool listed = emit_ok && w_arcs <= kFirstCap;  // otherwise (rare) count now, list chunk by chunk later
    uint32_t n_w = 0;
    if (emit_ok) n_w = list_firsts(0, w_arcs, listed);
    if (lane == 0) s_warp[wid] = n_w;
    __syncthreads();
    uint32_t w_new_base = 0, cta_new = 0;
#pragma unroll
    for (uint32_t w = 0; w < kWarps; w++) { const uint32_t v = s_warp[w]; if (w < wid) w_new_base += v; cta_new += v; }
    if (tid == 0) {  // reserve the record region of this CTA's new states (its latency hides behind the setup loads below)
      uint32_t sb = 0;
      if (cta_new) {
        sb = atomicAdd(P.st_cursor, cta_new);
        if ((unsigned long long)sb + cta_new > P.states_cap) { atomicOr(&P.ctl[7], (uint32_t)kOvStates); sb = 0xFFFFFFFFu; }
      }
      s_misc[0] = sb;
    }
    // set the listed states up: lane k of a round takes list entry k (perfectly balanced, ranks = list positions).  The
    // loads of the first round are issued before the CTA barrier that hands out the region base.
    uint32_t run_items = 0, sbase = 0;
    auto setup_load = [&](uint32_t n, uint32_t k0, uint32_t& nit, uint32_t& h, StRec& r, float& fin) {
      nit = 0; h = 0;
      const uint32_t kx = k0 + lane;
      if (kx < n) {
        h = firsts[kx];
        const uint4 sv = ld_volatile_u4(&P.slots[h]);
        st_relaxed_u32(&P.slots[h].id, kClaimed);
        nit = setup_state_ws(P, (unsigned long long)sv.x | ((unsigned long long)sv.y << 32), gw, r, fin);
      }
    };
    auto setup_store = [&](uint32_t n, uint32_t k0, uint32_t rank0, uint32_t nit, uint32_t h, StRec& r, float fin) {
      const uint32_t kx = k0 + lane, items = nit & ~kSideBit;
      uint32_t inc = items;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, inc, o); if ((int)lane >= o) inc += u; }
      if (kx < n && sbase != 0xFFFFFFFFu) {
        const uint32_t ridx = sbase + w_new_base + rank0 + kx;
        r.item_loc = (run_items + inc - items) | (nit & kSideBit);
        store_rec(&P.st_rec[ridx], r);
        P.st_cold[ridx] = StCold{h, fin};
      }
      run_items += __shfl_sync(0xFFFFFFFFu, inc, 31);
    };
    {
      uint32_t nit = 0, h = 0;
      StRec r{};
      float fin = 0.f;
      const uint32_t n_first = listed ? n_w : 0u;
      if (n_first) setup_load(n_first, 0, nit, h, r, fin);
      __syncthreads();
      sbase = s_misc[0];
      if (n_first) {
        setup_store(n_first, 0, 0, nit, h, r, fin);
        for (uint32_t k0 = 32; k0 < n_first; k0 += 32) {
          setup_load(n_first, k0, nit, h, r, fin);
          setup_store(n_first, k0, 0, nit, h, r, fin);
        }
      } else if (emit_ok && n_w) {  // more arcs than the list holds: list and set up chunk by chunk
        uint32_t rank0 = 0;
        for (uint32_t eb = 0; eb < w_arcs; eb += kFirstCap) {
          const uint32_t n = list_firsts(eb, min(w_arcs, eb + kFirstCap), true);
          __syncwarp();
          for (uint32_t k0 = 0; k0 < n; k0 += 32) {
            setup_load(n, k0, nit, h, r, fin);
            setup_store(n, k0, rank0, nit, h, r, fin);
          }
          rank0 += n;
          __syncwarp();
        }
      }
    }
    if (lane == 0) P.witems[gw] = run_items;
    if (run_items >= 0x7FFFFFFFu && lane == 0) atomicOr(&P.ctl[7], (uint32_t)kOvScratch);  // bit 31 is the side bit
    publish2(P, c, tag + 1, cta_new, s_misc[0] != 0xFFFFFFFFu ? s_misc[0] : 0u);
    const unsigned long long tr1 = globaltimer_ns();
    n_states_exp += F; n_items += T; n_waves++;
    lo += F; run_base += w_act; tag++;
    t_wait += tw1 - tw0; t_match += tm1 - tw1; t_res += tm2 - tm1; t_emit += te1 - tm2; t_bar += tb1 - te1; t_rank += tr1 - tb1;
  }

  if (lane == 0 && my_arcs_total) atomicAdd(&P.stats[2], my_arcs_total);
  if (c == 0 && tid == 0) {
    atomicOr(&P.ctl[1], overflow);
    P.ctl[2] = lo;        // number of product states
    P.ctl[3] = run_base;  // number of arc runs
    P.ctl[4] = (uint32_t)n_waves;
    P.wave_lo[n_waves] = lo;
    P.stats[0] = n_states_exp; P.stats[1] = n_items - n_states_exp; P.stats[3] = n_waves;
    P.stats[4] = t_match; P.stats[5] = t_emit; P.stats[6] = t_rank; P.stats[7] = t_wait; P.stats[8] = t_bar; P.stats[9] = t_res;
  }
}

// Seeds the state table with the start tuples: id i <- (start_fs, starts1[i], start2).  The s1 components are
// pairwise distinct (different acceptors of a batch live in disjoint state ranges), so are the keys.
__global__ void k_ws_init(Slot* slots, uint32_t mask, unsigned long long* tuples, uint32_t start_fs,
                          const uint32_t* __restrict__ starts1, uint32_t single_start1, uint32_t start2,
                          uint32_t n_starts, uint32_t* ctl) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { for (int k = 0; k < 12; k++) ctl[k] = 0; }
  if (i >= n_starts) return;
  const unsigned long long key = pack_key(start_fs, starts1 ? starts1[i] : single_start1, start2);
  uint32_t h = hash_key(key) & mask;
  while (atomicCAS(&slots[h].key, kEmptyKey, key) != kEmptyKey) h = (h + 1) & mask;
  slots[h].id = i; slots[h].emin = 0;
  tuples[i] = key;
}

// CSR offsets of the result: state s begins at (canonical base of the run that holds its first arc) + (its index in
// that run); run_dst = exclusive prefix of the run lengths in (wave, warp) order = canonical emission order.
__global__ void k_ws_offsets(const uint2* __restrict__ st_first, const uint32_t* __restrict__ run_dst, uint32_t n,
                             uint32_t n_runs, uint32_t* __restrict__ off) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) { const uint2 f = st_first[s]; off[s] = __ldg(&run_dst[f.x]) + f.y; }
  else if (s == n) off[n] = __ldg(&run_dst[n_runs]);
}

// Canonical order of the arcs = runs in (wave, warp) order; run_dst = exclusive prefix of the run lengths.  One warp per
// run, four arcs per lane in flight.  Pending targets (slot index) are resolved to the published state id on the way.
// k_ws_move_next only produces the dense array of (resolved) next states — all the trim needs before its final gather,
// which then reads the arcs straight from their provisional runs; k_ws_move_arcs moves whole arcs (connect = false).
constexpr uint32_t kMoveWarps = 8;
template <bool kWholeArcs>
__global__ void __launch_bounds__(kMoveWarps * 32)
k_ws_move(const Tr* __restrict__ prov, const uint32_t* __restrict__ prov_next, const uint32_t* __restrict__ run_src,
          const uint32_t* __restrict__ run_cnt, const uint32_t* __restrict__ run_dst, uint32_t n_runs,
          const Slot* __restrict__ slots, Tr* __restrict__ out_arcs, uint32_t* __restrict__ out_next) {
  const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
  const uint32_t warps = gridDim.x * kMoveWarps;
  for (uint32_t r = blockIdx.x * kMoveWarps + wid; r < n_runs; r += warps) {
    const uint32_t cnt = __ldg(&run_cnt[r]);
    if (!cnt) continue;
    const uint32_t first = __ldg(&run_src[r]);
    const Tr* __restrict__ src = prov + first;
    const uint32_t* __restrict__ src_next = prov_next + first;
    const uint32_t dst = __ldg(&run_dst[r]);
    for (uint32_t e0 = 0; e0 < cnt; e0 += 128) {
      int4 v[4];
      uint32_t ns[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const uint32_t e = e0 + 32u * q + lane;
        if (kWholeArcs) { v[q] = e < cnt ? __ldg(reinterpret_cast<const int4*>(&src[e])) : make_int4(0, 0, 0, 0); ns[q] = (uint32_t)v[q].w; }
        else ns[q] = e < cnt ? __ldg(&src_next[e]) : 0u;
      }
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (ns[q] & kPendingBit) ns[q] = __ldg(&slots[ns[q] & ~kPendingBit].id);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const uint32_t e = e0 + 32u * q + lane;
        if (e < cnt) {
          if (kWholeArcs) { v[q].w = (int)ns[q]; *reinterpret_cast<int4*>(&out_arcs[dst + e]) = v[q]; }
          else out_next[dst + e] = ns[q];
        }
      }
    }
  }
}

static int ws_grid_used = 1;
float run_ws(const WsParams& P0, int sms, cudaStream_t s) {
  WsParams P = P0;
  void* kern = (void*)k_compose_ws;
  int grid = sms;  // one 768-thread CTA per SM
  if (const char* e = std::getenv("B200_WS_GRID")) grid = std::max(1, std::atoi(e));
  if (grid > 2 * kCoopThreads || grid * (int)kWarps >= (1 << (32 - kKeyBits)))
    throw FstError("compose kernel: grid too large for the emission key");
  if (grid * (int)kWarps > (int)(kWarpsPerThreadMax * kCoopThreads)) throw FstError("compose kernel: grid too large for the exchange");
  const size_t dyn = kWarps * sizeof(WarpSmem) + (2 * (size_t)grid + 1 + (size_t)grid * kWarps + 1 + 4) * sizeof(uint32_t);
  B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  int per_sm = 0;
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kCoopThreads, dyn));
  if (per_sm < 1) throw FstError("compose kernel does not fit on the device");
  if (grid > sms * per_sm) grid = sms * per_sm;
  ws_grid_used = grid;
  void* args[] = {(void*)&P};
  cudaEvent_t e0, e1;
  float ms = 0;
  B200_CUDA(cudaEventCreate(&e0)); B200_CUDA(cudaEventCreate(&e1));
  B200_CUDA(cudaEventRecord(e0, s));
  B200_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kCoopThreads), args, dyn, s));
  B200_CUDA(cudaEventRecord(e1, s));
  B200_CUDA(cudaStreamSynchronize(s));
  B200_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return ms;
}

struct Events3 {  // RAII: the three timing events of a call
  cudaEvent_t e[3] = {nullptr, nullptr, nullptr};
  Events3() { for (auto& x : e) B200_CUDA(cudaEventCreate(&x)); }
  ~Events3() { for (auto& x : e) if (x) cudaEventDestroy(x); }
};

}  // namespace

// Capacities of one attempt; compose_device() doubles whatever overflowed and runs again.
int compose_device_ws(const DevFst& fa, const DevFst& fb, const ComposeOptions& opt, ComposeStats* stats,
                      cudaStream_t s, DevFst* result, const BatchStarts* batch, WsCaps* caps) {
  int kind = opt.filter == kAutoFilter ? kSequenceFilter : opt.filter;
  if (kind < kNullFilter || kind > kNoMatchFilter) throw FstError("EnumConversionError");
  // ---- sigma matcher configs: construction and REQUIRE_MATCH checks (sigma_matcher.rs:55-84,126-132,
  // compose_fst_op.rs:170-179, compose_static.rs:219-223)
  if (opt.filter == kAutoFilter && (opt.sigma1.enabled || opt.sigma2.enabled))
    throw FstError("Custom MatcherConfig not supported with AutoFilter");
  for (const SigmaSpec* sp : {&opt.sigma1, &opt.sigma2}) {
    if (!sp->enabled) continue;
    if (sp->rewrite_mode < 0 || sp->rewrite_mode > 2) throw FstError("EnumConversionError");
    if (sp->sigma_label == kEps) throw FstError("SigmaMatcher: 0 cannot be used as sigma_label");
  }
  auto require_sorted = [](uint64_t p, uint64_t yes, uint64_t no, const char* which, const char* known) {
    if (!(p & (yes | no))) throw FstError(std::string("Properties are not known : ") + known);
    if (!(p & yes)) throw FstError(std::string("ComposeFst: ") + which + " argument cannot perform required matching (sort?)");
  };
  if (opt.sigma1.enabled && opt.sigma1.sigma_label != kNoLabel)
    require_sorted(fa.props, props::kOLabelSorted, props::kNotOLabelSorted, "1st", "O_LABEL_SORTED | NOT_O_LABEL_SORTED");
  if (opt.sigma2.enabled && opt.sigma2.sigma_label != kNoLabel)
    require_sorted(fb.props, props::kILabelSorted, props::kNotILabelSorted, "2nd", "I_LABEL_SORTED | NOT_I_LABEL_SORTED");
  int side = resolve_match_side(fa.props, fb.props);
  if (fa.num_states >= 0x7FFFFFFFu || fb.num_states >= 0x7FFFFFFFu)
    throw FstError("compose: operands with >= 2^31 states are not supported");
  if (!batch && (!fa.has_start || !fb.has_start)) {
    // empty result (compose_fst_op.rs:389-404, lazy_fst.rs:229-232): the multi-kernel back end builds it
    ComposeOptions plain = opt; plain.sigma1 = SigmaSpec(); plain.sigma2 = SigmaSpec();
    *result = compose_device_waves(fa, fb, plain, stats, s);
    return 0;
  }
  if (batch && (!fb.has_start || batch->n == 0)) throw FstError("batched compose needs a start state on both sides");
  const uint32_t n_starts = batch ? batch->n : 1u;

  ComposeStats local;
  ComposeStats& st = stats ? *stats : local;
  st = ComposeStats();
  Events3 ev;
  B200_CUDA(cudaEventRecord(ev.e[0], s));

  DevBuf<uint32_t> neps1(s), neps2(s);
  bool need_eps = (kind == kSequenceFilter || kind == kAltSequenceFilter || kind == kMatchFilter);
  WsParams P{};
  P.a = FstView{fa.offsets.p, fa.arcs.p, fa.finals.p, nullptr, fa.num_states};
  P.b = FstView{fb.offsets.p, fb.arcs.p, fb.finals.p, nullptr, fb.num_states};
  if (need_eps && !(fa.props & props::kNoOEpsilons)) {
    neps1.reserve_discard(fa.num_states);
    launch_count_eps(P.a.off, P.a.arcs, P.a.n, 1, neps1.p, s);
    P.a.neps = neps1.p; st.kernel_launches++;
  }
  if (need_eps && !(fb.props & props::kNoIEpsilons)) {
    neps2.reserve_discard(fb.num_states);
    launch_count_eps(P.b.off, P.b.arcs, P.b.n, 0, neps2.p, s);
    P.b.neps = neps2.p; st.kernel_launches++;
  }
  P.kind = kind; P.side = side;
  bool built1 = false, built2 = false;
  P.lab1 = label_column(fa, true, kLabelPad, s, &built1);   // cached with the machine: resident operands pay once
  P.lab2 = label_column(fb, false, kLabelPad, s, &built2);
  st.kernel_launches += (built1 ? 1 : 0) + (built2 ? 1 : 0);
  DevBuf<uint32_t> allowed1(s), allowed2(s);
  bool host_temporaries = false;
  auto mk_sigma = [&](const SigmaSpec& sp, uint64_t fprops, DevBuf<uint32_t>& buf) {
    SigmaDev d{};
    if (!sp.enabled) return d;
    d.enabled = 1; d.label = sp.sigma_label;
    d.rewrite_both = sp.rewrite_mode == 1 || (sp.rewrite_mode == 0 && (fprops & props::kAcceptor));
    d.n_allowed = (uint32_t)sp.allowed.size();
    if (d.n_allowed) {
      buf.reserve_discard(d.n_allowed);
      B200_CUDA(cudaMemcpyAsync(buf.p, sp.allowed.data(), (size_t)d.n_allowed * 4, cudaMemcpyHostToDevice, s));
      d.allowed = buf.p;
      host_temporaries = true;
    }
    return d;
  };
  P.sig1 = mk_sigma(opt.sigma1, fa.props, allowed1);
  P.sig2 = mk_sigma(opt.sigma2, fb.props, allowed2);
  if (host_temporaries) B200_CUDA(cudaStreamSynchronize(s));

  // ---- buffers sized from the operands; an overflow reports which one to grow (compose_device retries)
  const size_t sum_states = (size_t)fa.num_states + fb.num_states, sum_arcs = (size_t)fa.num_arcs + fb.num_arcs;
  WsCaps cp = caps ? *caps : WsCaps();
  if (!cp.states) cp.states = std::max<size_t>(1 << 16, 4 * sum_states + 2 * (size_t)n_starts);
  if (!cp.arcs) cp.arcs = std::max<size_t>(1 << 18, 4 * sum_arcs);
  if (!cp.items) cp.items = std::max<size_t>(1 << 18, cp.arcs / 2);
  if (!cp.runs) cp.runs = std::max<size_t>(1 << 20, cp.states / 4);
  if (!cp.waves) cp.waves = 1u << 20;
  cp.states = std::min<size_t>(cp.states, 0x7FFFFFF0ull);
  cp.arcs = std::min<size_t>(cp.arcs, 0xFFFFFFF0ull);
  cp.items = std::min<size_t>(cp.items, 0xFFFFFF00ull);
  cp.runs = std::min<size_t>(cp.runs, 0xFFFFFFF0ull);
  size_t table_cap = 1 << 17;
  while (table_cap < 2 * cp.states && table_cap < (1ull << 31)) table_cap <<= 1;
  if (caps) *caps = cp;

  DevFst out(s);
  out.finals.reserve_discard(cp.states);
  DevBuf<Tr> prov_arcs(s, cp.arcs);
  DevBuf<uint32_t> prov_next(s, cp.arcs);
  DevBuf<unsigned long long> tuples(s, cp.states), dstats(s, 16), cursors(s, 2);
  DevBuf<uint2> st_first(s, cp.states);
  DevBuf<Slot> slots(s, table_cap);
  DevBuf<StRec> st_rec(s, cp.states);
  DevBuf<StCold> st_cold(s, cp.states);
  DevBuf<uint4> recs(s, cp.items);
  DevBuf<uint32_t> arc_loc(s, cp.items + 64), ctl(s, 12), run_src(s, cp.runs), run_cnt(s, cp.runs + 1), wave_lo(s, cp.waves);
  DevBuf<unsigned long long> parts(s, 2 * 2048);
  DevBuf<uint32_t> witems(s, 8192);
  B200_CUDA(cudaMemsetAsync(slots.p, 0xFF, table_cap * sizeof(Slot), s));
  B200_CUDA(cudaMemsetAsync(dstats.p, 0, 16 * sizeof(unsigned long long), s));
  B200_CUDA(cudaMemsetAsync(cursors.p, 0, 2 * sizeof(unsigned long long), s));
  B200_CUDA(cudaMemsetAsync(parts.p, 0, 2 * 2048 * sizeof(unsigned long long), s));  // tag 0 = nothing published yet
  P.tuples = tuples.p; P.states_cap = (uint32_t)cp.states;
  P.out_finals = out.finals.p; P.st_first = st_first.p;
  P.prov_arcs = prov_arcs.p; P.prov_next = prov_next.p; P.arcs_cap = (uint32_t)cp.arcs; P.arc_cursor = cursors.p;
  P.run_src = run_src.p; P.run_cnt = run_cnt.p; P.runs_cap = (uint32_t)cp.runs;
  P.slots = slots.p; P.mask = (uint32_t)table_cap - 1; P.table_cap = (uint32_t)std::min<size_t>(table_cap, 0xFFFFFFFFull);
  P.st_rec = st_rec.p; P.st_cold = st_cold.p; P.st_cursor = reinterpret_cast<uint32_t*>(cursors.p + 1);
  P.recs = recs.p; P.arc_loc = arc_loc.p; P.items_cap = (uint32_t)cp.items;
  P.part_new = parts.p; P.part_sbase = parts.p + 2048; P.witems = witems.p;
  P.ctl = ctl.p; P.stats = dstats.p;
  P.barrier = ctl.p + 6;
  P.wave_lo = wave_lo.p; P.wave_cap = (uint32_t)cp.waves;
  uint32_t start_fs = (kind == kNullFilter || kind == kTrivialFilter || kind == kNoMatchFilter) ? 1u : 0u;
  P.n_starts = n_starts;
  k_ws_init<<<blocks_for(n_starts), kThreads, 0, s>>>(slots.p, P.mask, tuples.p, start_fs,
                                                      batch ? batch->d_starts1 : nullptr, fa.start, fb.start, n_starts, ctl.p);
  st.kernel_launches++;
  float ms_kernel = run_ws(P, sm_count(), s);
  st.kernel_launches++; st.emit_launches = 1;

  uint32_t hctl[8];
  unsigned long long hstats[16];
  B200_CUDA(cudaMemcpyAsync(hctl, ctl.p, 32, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaMemcpyAsync(hstats, dstats.p, 16 * 8, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  hctl[1] |= hctl[7];
  if (hctl[1] != 0) {
    if (hctl[1] & kErrBothRequire) throw FstError("Both sides can't require match");           // compose_fst_op.rs:207-209
    if (hctl[1] & kErrBadSigmaLabel) throw FstError("SigmaMatcher::Find: bad label (sigma)");  // sigma_matcher.rs:205-207
    if (hctl[1] & kErrWatchdog) throw FstError("compose kernel: a grid-wide wait exceeded 10 s (watchdog); no result");
    return (int)hctl[1];  // which pre-sized buffer was too small
  }
  const uint32_t n_states = hctl[2], n_runs = hctl[3];
  st.states_expanded = hstats[0]; st.arcs_iterated = hstats[1]; st.arcs_emitted = hstats[2]; st.waves = hstats[3];
  st.ms_emit_kernel = ms_kernel;
  st.ms_phase[0] = hstats[4] * 1e-6f; st.ms_phase[1] = hstats[5] * 1e-6f;
  st.ms_phase[2] = hstats[6] * 1e-6f; st.ms_phase[3] = hstats[7] * 1e-6f;
  if (std::getenv("B200_COOP_TRACE"))
    std::fprintf(stderr, "[ws] %d CTAs, kernel %.3f ms; CTA 0 / warp 0: match %.3f, reserve %.3f, emit %.3f, barrier wait %.3f, "
                 "rank+setup %.3f, exchange wait %.3f\n", ws_grid_used, ms_kernel, st.ms_phase[0], hstats[9] * 1e-6,
                 st.ms_phase[1], hstats[8] * 1e-6, st.ms_phase[2], st.ms_phase[3]);

  // ---- canonical arc order: prefix over the run lengths, CSR offsets, one move of every run
  const uint32_t n_arcs = (uint32_t)st.arcs_emitted;
  DevBuf<uint32_t> run_dst(s, (size_t)n_runs + 1);
  DevBuf<uint8_t> scan_tmp(s);
  B200_CUDA(cudaMemsetAsync(run_cnt.p + n_runs, 0, 4, s));
  exclusive_sum_u32(run_cnt.p, run_dst.p, (size_t)n_runs + 1, scan_tmp, s);
  out.offsets.reserve_discard((size_t)n_states + 1);
  k_ws_offsets<<<blocks_for((size_t)n_states + 1), kThreads, 0, s>>>(st_first.p, run_dst.p, n_states, n_runs, out.offsets.p);
  const unsigned move_grid = std::min<unsigned>((n_runs + kMoveWarps - 1) / kMoveWarps, (unsigned)sm_count() * 8u);
  DevBuf<uint32_t> next(s);
  if (opt.connect) {  // the trim only needs the resolved next states; it gathers the arcs from their runs itself
    next.reserve_discard(n_arcs ? n_arcs : 1);
    if (n_runs) k_ws_move<false><<<move_grid, kMoveWarps * 32, 0, s>>>(prov_arcs.p, prov_next.p, run_src.p, run_cnt.p, run_dst.p, n_runs, slots.p, nullptr, next.p);
  } else {
    out.arcs.reserve_discard(n_arcs ? n_arcs : 1);
    if (n_runs) k_ws_move<true><<<move_grid, kMoveWarps * 32, 0, s>>>(prov_arcs.p, prov_next.p, run_src.p, run_cnt.p, run_dst.p, n_runs, slots.p, out.arcs.p, nullptr);
  }
  st.kernel_launches += 3;
  out.num_states = n_states; out.num_arcs = n_arcs;
  out.has_start = true; out.start = 0;
  out.props = props::of_compose(fa.props, fb.props);
  B200_CUDA(cudaEventRecord(ev.e[1], s));
  if (opt.connect) {
    uint64_t launches = 0;
    TrimExtras extras;
    extras.tuples = batch ? tuples.p : nullptr;
    extras.n_starts = n_starts;
    extras.out_tag = batch ? batch->out_s1 : nullptr;
    extras.out_start_map = batch ? batch->out_start_map : nullptr;
    ProvArcs pa;
    pa.prov = prov_arcs.p; pa.st_first = st_first.p; pa.run_src = run_src.p; pa.run_cnt = run_cnt.p; pa.next = next.p;
    pa.run_dst = run_dst.p; pa.n_runs = n_runs;
    DevFst trimmed = connect_waves_device(out, wave_lo.p, (uint32_t)st.waves, &launches, s, &extras, &pa);
    st.kernel_launches += launches;
    out = std::move(trimmed);
  } else if (batch) {
    batch->out_s1->reserve_discard(out.num_states ? out.num_states : 1);
    batch->out_start_map->reserve_discard(n_starts);
    launch_unpack_s1(tuples.p, out.num_states, batch->out_s1->p, n_starts, batch->out_start_map->p, s);
    st.kernel_launches++;
  }
  st.states_out = out.num_states; st.arcs_out = out.num_arcs;
  B200_CUDA(cudaEventRecord(ev.e[2], s));
  B200_CUDA(cudaStreamSynchronize(s));
  B200_CUDA(cudaEventElapsedTime(&st.ms_expand, ev.e[0], ev.e[1]));
  B200_CUDA(cudaEventElapsedTime(&st.ms_connect, ev.e[1], ev.e[2]));
  *result = std::move(out);
  return 0;
}

}  // namespace b200
