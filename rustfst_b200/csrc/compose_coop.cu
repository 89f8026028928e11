// compose_coop.cu — the whole composition BFS as ONE persistent cooperative kernel (B200: 148 SMs x resident CTAs).
//
// Same algorithm and the same canonical numbering as compose.cu (see its header for the reference citations); what
// changes is the execution model: instead of ~9 launches and 3 host read-backs per BFS wave, a grid that exactly
// fills the machine stays resident and separates the phases of a wave with grid-wide barriers.  A wave
// (frontier = product ids [lo, hi)) is four phases:
//
//   A  match   a group of G lanes per frontier state: lane j takes item j (item 0 = implicit epsilon loop, item j =
//              j-th arc of the iterated side), binary-searches the sorted side, applies the filter and records
//              (pos, count, filter states); per-state / per-group / per-CTA emission counts
//   B  emit    CTA prefix over the per-CTA counts gives every state its canonical first emission index; the same
//              group re-walks its states and writes the 16-byte output arcs (128-bit stores), looks the destination
//              tuple up in the open-addressed table (CAS insert) and atomicMin's its first emission index
//   C  rank    every arc of the wave, in emission order: "am I the first emission of a new tuple?"; CTA-local
//              exclusive ranks (ballot + popc), per-CTA counts
//   D  resolve CTA prefix over those counts turns (CTA, local rank) into the state id of the next wave; pending
//              arcs get their nextstate, first emitters publish id + tuple
//
// Frontier bounds, arc totals and overflow decisions are recomputed identically by every CTA from the per-CTA
// partial arrays, so control flow is uniform without any broadcast.  All buffers are pre-sized; if one would
// overflow the kernel stops and compose_device() falls back to the growing multi-kernel back end.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

#include "compose_match.cuh"
#include "coop_utils.cuh"

namespace cg = cooperative_groups;

namespace b200 {
namespace {
using namespace composeimpl;
using namespace coop;

constexpr uint32_t kTempFlag = 0x80000000u;   // slot.id holds (cta << 20 | local rank) between phases C and D
constexpr uint32_t kLocalRankBits = 20;

struct CoopParams {
  FstView a, b;
  const uint32_t* lab1; const uint32_t* lab2;  // fst1 olabels / fst2 ilabels as dense arrays (padded)
  int kind, side;
  unsigned long long* tuples; uint32_t states_cap;
  uint32_t* out_offsets; float* out_finals;
  Tr* out_arcs; uint32_t arcs_cap;
  Slot* slots; uint32_t mask; uint32_t table_cap;
  // per-wave scratch, indexed by frontier-local state
  uint32_t* item_loc;     // slice-local exclusive item offset | side bit
  uint32_t* st_arc_loc;   // slice-local emission index of the state's first arc
  uint32_t* st_meta;      // alleps1 | noeps1<<1 | alleps2<<2 | noeps2<<3 | fs<<4 | searched side has sigma<<6 | owner CTA<<8
  uint4* st_off;          // arc ranges of the two component states: alo, ahi, blo, bhi
  // per-wave scratch, indexed by item (active items compacted in place inside each CTA's slice)
  uint4* recs;            // x = first match, y = count|flags, z = frontier-local state, w = iterated arc (abs) or ~0
  uint32_t* arc_loc;      // warp-local emission index of the record's first arc
  uint32_t* wpref_arcs;   // per warp (CTA * 8 + warp): arcs emitted by the lower warps of the same CTA in this wave
  uint32_t items_cap;
  // epoch-tagged count exchanges (coop_utils.cuh), gridDim entries each
  unsigned long long* part_items;   // items of each CTA's state slice
  unsigned long long* part_arcs;    // arcs emitted by each CTA's item slice
  unsigned long long* part_new;     // first emissions found in each CTA's arc slice
  uint32_t* ctl;          // [1] overflow / error flags, [2] #states, [3] #arcs, [4] errors raised while matching
  uint32_t* wave_lo; uint32_t wave_cap;  // first product id of every BFS wave (+ one-past-the-end sentinel)
  unsigned int* barrier;  // arrival counter of grid_barrier (zero-initialised)
  SigmaDev sig1, sig2;    // sigma matcher on fst1 (olabel side) / fst2 (ilabel side)
  uint32_t poll_ns;       // back-off between polls of a count word that is not there yet (0 = spin)
  uint32_t n_starts;      // initial frontier = product ids [0, n_starts) (1 for a plain compose, batch size otherwise)
  unsigned long long* cta_trace;  // optional (B200_COOP_TRACE): per CTA {smid, busy A1, B, C, D} ; may be null
  unsigned long long* stats;  // states_expanded, arcs_iterated, arcs_emitted, waves, ns phase A, B, C, D
};


constexpr uint32_t kTile = kCoopThreads;
constexpr uint32_t kWarps = kCoopThreads / 32;
constexpr int kArcsPerThread = 4;  // rank / resolve phases: consecutive arcs per thread and round
#ifndef B200_EMIT_ARCS_PER_LANE
#define B200_EMIT_ARCS_PER_LANE 1
#endif
constexpr int kEmitArcs = B200_EMIT_ARCS_PER_LANE;  // emit phase: arcs per lane and round (32 * kEmitArcs per warp)

// Optional fine-grained timeline of thread 0 of every CTA (build with -DB200_COOP_PROFILE): SM cycles spent in the
// sub-steps of the A1 and B tiles, summed over the run and averaged over the CTAs by the host.
#ifdef B200_COOP_PROFILE
#define PROF_DECL long long prof_t = 0; unsigned long long prof[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define PROF_START() do { prof_t = clock64(); } while (0)
#define PROF_MARK(k) do { long long now_ = clock64(); prof[k] += (unsigned long long)(now_ - prof_t); prof_t = now_; } while (0)
#define PROF_USE(x) asm volatile("" ::"r"(x) : "memory")
#else
#define PROF_DECL
#define PROF_START() do { } while (0)
#define PROF_MARK(k) do { } while (0)
#define PROF_USE(x) do { } while (0)
#endif

__global__ void k_extract_labels(const Tr* __restrict__ arcs, uint32_t n, uint32_t n_padded, int olabel,
                                 uint32_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = olabel ? __ldg(&arcs[i].olabel) : __ldg(&arcs[i].ilabel);
  else if (i < n_padded) out[i] = kNoLabel;
}
// State table lookup-or-insert: linear probing, kW consecutive slots fetched per round.  Loaded non-empty keys are permanent (no deletions); loaded empties are confirmed by the CAS.  Returns the
// id stored in the slot (kUnassigned for a tuple discovered in this wave) and the slot index in h.
template <int kW>
__device__ __forceinline__ uint32_t table_probe(Slot* slots, uint32_t mask, unsigned long long key, uint32_t& h) {
  while (true) {
    uint4 sv[kW];
#pragma unroll
    for (int q = 0; q < kW; q++) sv[q] = ld_volatile_u4(&slots[(h + q) & mask]);
#pragma unroll
    for (int q = 0; q < kW; q++) {
      const unsigned long long curk = (unsigned long long)sv[q].x | ((unsigned long long)sv[q].y << 32);
      const uint32_t hq = (h + q) & mask;
      if (curk == key) { h = hq; return sv[q].z; }
      if (curk == kEmptyKey) {
        const unsigned long long prev = atomicCAS(&slots[hq].key, kEmptyKey, key);
        if (prev == kEmptyKey) { h = hq; return kUnassigned; }  // fresh insert: discovered in this wave
        if (prev == key) { h = hq; return *reinterpret_cast<volatile uint32_t*>(&slots[hq].id); }
      }
    }
    h = (h + kW) & mask;
  }
}

// Same lookup-or-insert with the first slot already fetched (the emit phase issues the first probes of all the arcs of
// a lane before it waits for any of them; with kEmitArcs = 1 this is the plain probe).
__device__ __forceinline__ uint32_t table_probe_from(Slot* slots, uint32_t mask, unsigned long long key, uint32_t& h,
                                                     uint4 sv) {
  while (true) {
    const unsigned long long curk = (unsigned long long)sv.x | ((unsigned long long)sv.y << 32);
    if (curk == key) return sv.z;
    if (curk == kEmptyKey) {
      const unsigned long long prev = atomicCAS(&slots[h].key, kEmptyKey, key);
      if (prev == kEmptyKey) return kUnassigned;  // fresh insert: discovered in this wave
      if (prev == key) return *reinterpret_cast<volatile uint32_t*>(&slots[h].id);
    }
    h = (h + 1) & mask;
    sv = ld_volatile_u4(&slots[h]);
  }
}

// Per-state setup of a product state that joins the next frontier (compose_fst_op.rs:199-219 match side,
// :420-449 final weight; filter flags as in compose_common.cuh): writes the state's scratch records and returns
// (#items | side bit).  i = frontier-local index, id = product state id, c = CTA that owns the state's slice.
__device__ __forceinline__ uint32_t setup_state(const CoopParams& P, unsigned long long key, uint32_t i, uint32_t id,
                                                uint32_t c) {
  uint32_t fs, s1, s2;
  unpack_key(key, fs, s1, s2);
  const uint32_t alo = __ldg(&P.a.off[s1]), ahi = __ldg(&P.a.off[s1 + 1]);
  const uint32_t blo = __ldg(&P.b.off[s2]), bhi = __ldg(&P.b.off[s2 + 1]);
  const float f1 = __ldg(&P.a.fin[s1]), f2 = __ldg(&P.b.fin[s2]);
  const uint32_t ne1 = P.a.neps ? __ldg(&P.a.neps[s1]) : 0u, ne2 = P.b.neps ? __ldg(&P.b.neps[s2]) : 0u;
  const uint32_t d1 = ahi - alo, d2 = bhi - blo;
  P.st_off[i] = make_uint4(alo, ahi, blo, bhi);
  bool mi = P.side == kMatchInput || (P.side == kMatchBoth && d1 <= d2);
  bool hs1 = false, hs2 = false;
  if (P.sig1.enabled) hs1 = dev_has_sigma<true>(P.sig1, P.a.arcs, alo, ahi);
  if (P.sig2.enabled) hs2 = dev_has_sigma<false>(P.sig2, P.b.arcs, blo, bhi);
  if (P.side == kMatchBoth && (hs1 || hs2)) {  // SigmaMatcher::priority = REQUIRE_PRIORITY (compose_fst_op.rs:199-219)
    if (hs1 && hs2) atomicOr(&P.ctl[1], (uint32_t)kErrBothRequire);
    mi = hs2;
  }
  const bool hs_searched = mi ? hs2 : hs1;
  const uint32_t fl = ((d1 == ne1 && f1 == w_zero()) ? 1u : 0u) | ((ne1 == 0) ? 2u : 0u) |
                      ((d2 == ne2 && f2 == w_zero()) ? 4u : 0u) | ((ne2 == 0) ? 8u : 0u) | (fs << 4) |
                      (hs_searched ? 64u : 0u);
  P.st_meta[i] = fl | (c << 8);
  const float fw = w_times(f1, f2);
  P.out_finals[id] = w_is_zero(fw) ? w_zero() : fw;
  return (1u + (mi ? d1 : d2)) | (mi ? kSideBit : 0u);
}

// One persistent kernel = the whole BFS.  A frontier state belongs to the slice of the CTA that discovered it
// (slice c = frontier-local ids [pref_new[c], pref_new[c+1])).  Per wave:
//   A1 items    : CTA c owns a contiguous slice of the ITEMS (item 0 of a state = the implicit epsilon loop, item j =
//                 j-th arc of the iterated side); load-balanced: the states of a 256-item tile are staged in a
//                 257-entry shared-memory window (offsets, flags, arc ranges); per item: match_range on the sorted
//                 side + filter; active items are compacted in place with CTA-local arc offsets
//   B  arcs     : CTA c emits the arcs of its own items, one thread per ARC (records staged in shared memory);
//                 128-bit gathers, one 128-bit probe of the state table (CAS insert), atomicMin(first emission),
//                 128-bit store
//   C  rank     : contiguous arc slices; ballot/popc ranks of first emissions published through slot.id
//   D  resolve  : CTA prefix -> ids; nextstate patch; tuple publication; and the per-state setup of the NEXT
//                 frontier (the first emitter holds the tuple already), slice-local item offsets
// The three count exchanges (items, arcs, new states) are epoch-tagged per-CTA words: every CTA spins until all
// gridDim words of the current epoch are there and builds the same exclusive prefix in shared memory, which is a
// full grid barrier and the broadcast in one round trip.  Only B -> C (all atomicMin's must have landed) uses the
// arrival-counter barrier.  Control flow stays uniform across the grid by construction.
template <int kMinBlocks>
__global__ void __launch_bounds__(kCoopThreads, kMinBlocks)
k_compose_coop(CoopParams P) {
  __shared__ uint32_t s_warp[2 * (kCoopThreads / 32)];
  __shared__ uint32_t s_wseg[kWarps][32 * kEmitArcs + 4];   // per-warp tile windows (A1: 32 states, B: records)
  __shared__ uint32_t s_wmeta8[kWarps][36];
  __shared__ uint4 s_wrec8[kWarps][32 * kEmitArcs + 1];
  extern __shared__ uint32_t s_dyn[];          // three prefix arrays of gridDim + 1 entries
  const uint32_t G = gridDim.x, c = blockIdx.x, tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
  uint32_t* s_pref_items = s_dyn;
  uint32_t* s_pref_arcs = s_dyn + (G + 1);
  uint32_t* s_pref_new = s_dyn + 2 * (G + 1);

  unsigned int bar_epoch = 0;
  uint32_t tag = 1;                            // epoch of the tagged count exchanges
  uint32_t lo = 0, hi = P.n_starts, base = 0;  // uniform across the grid by construction
  unsigned long long n_states_exp = 0, n_items = 0, n_arcs = 0, n_waves = 0;
  unsigned long long t_a = 0, t_b = 0, t_c = 0, t_d = 0;
  unsigned long long busy[5] = {0, 0, 0, 0, 0};  // this CTA's own work time per phase (setup, A1, B, C, D), waits excluded
  uint32_t overflow = 0;
  PROF_DECL

  // ---------------------------------------------------------------------- setup of the initial frontier [0, n_starts)
  {
    unsigned long long t0 = globaltimer_ns();
    const uint32_t sc0 = (hi + G - 1) / G;
    const uint32_t s_begin = min(hi, c * sc0), s_end = min(hi, s_begin + sc0);
    uint32_t run = 0;
    for (uint32_t i0 = s_begin; i0 < s_end; i0 += kTile) {
      const uint32_t i = i0 + tid;
      uint32_t r = 0;
      if (i < s_end) r = setup_state(P, __ldcg(&P.tuples[i]), i, i, c);
      uint32_t tile_total;
      const uint32_t ex = cta_exclusive_scan(r & ~kSideBit, s_warp, tile_total);
      if (i < s_end) P.item_loc[i] = (run + ex) | (r & kSideBit);
      run += tile_total;
    }
    publish_count(P.part_new, c, tag, s_end - s_begin);
    publish_count(P.part_items, c, tag, run);
    cta_prefix_wait_to_smem(P.part_new, G, tag, s_pref_new, s_warp, P.poll_ns);
    busy[0] += globaltimer_ns() - t0;
  }

  while (lo < hi) {
    const uint32_t F = hi - lo;
    unsigned long long tp0 = globaltimer_ns();
    if (n_waves + 1 >= P.wave_cap) { overflow |= kOvWaves; break; }  // uniform
    if (c == 0 && tid == 0) P.wave_lo[n_waves] = lo;

    // ------------------------------------------------------------------ A1: per-item matching
    cta_prefix_wait_to_smem(P.part_items, G, tag, s_pref_items, s_warp, P.poll_ns);
    unsigned long long ts1 = globaltimer_ns();
    const uint32_t T = s_pref_items[G];
    overflow = __ldcg(&P.ctl[1]);
    if (T > P.items_cap) overflow |= kOvScratch;
    if (overflow) break;  // uniform
    // Every WARP owns a contiguous run of wc items (CTA c = warps 8c .. 8c+7) and walks it in 32-item tiles without
    // any CTA-wide barrier: the kernel is bound by latency and synchronisation, not by bandwidth, and a CTA barrier per
    // 256-item tile made all eight warps wait for the slowest one several times per tile.
    const uint32_t wc = (T + G * kWarps - 1) / (G * kWarps);  // items per warp
    const uint32_t ic = wc * kWarps;                          // items per CTA
    const uint32_t wb = min(T, c * ic + wid * wc), we = min(T, wb + wc);
    uint32_t w_active = 0, w_arcs = 0;  // active records / emitted arcs of this warp (lane-uniform)
    if (wb < we) {
      uint32_t* const win_seg = s_wseg[wid];
      uint32_t* const win_meta = s_wmeta8[wid];
      uint4* const win_rec = s_wrec8[wid];
      // state containing my first item: producing CTA slice from the shared prefix, then a 32-ary search (one probe per
      // lane and round) over that slice's local item offsets
      uint32_t i_cur;
      {
        const uint32_t p = smem_segment(s_pref_items, G, wb);
        const uint32_t want = wb - s_pref_items[p];
        uint32_t l = s_pref_new[p], h = s_pref_new[p + 1];
        while (h - l > 1) {  // invariant: item_loc[l] <= want < item_loc[h] (or h = end of the slice)
          const uint32_t step = (h - l + 31u) >> 5, idx = l + lane * step;
          const bool ok = idx < h && (__ldcg(&P.item_loc[idx]) & ~kSideBit) <= want;
          const uint32_t n_ok = __popc(__ballot_sync(0xFFFFFFFFu, ok));  // ok lanes form a prefix, lane 0 is always ok
          l += (n_ok - 1u) * step;
          h = min(h, l + step);
        }
        i_cur = l;
      }
      // The window of a tile = states i_cur .. i_cur + 32 with their global item offsets, flags and arc ranges.  It is
      // loaded into registers one tile ahead (the loads fly while the current tile is matched) and committed to the
      // warp's shared-memory window at the top of the tile.  Entry 32 only bounds the search.
      uint32_t w_il = 0, w_meta = 0, w_il2 = 0, w_meta2 = 0;
      uint4 w_off = make_uint4(0, 0, 0, 0);
      auto window_load = [&](uint32_t i_base) {
        const uint32_t i = i_base + lane;
        if (i < F) { w_il = __ldcg(&P.item_loc[i]); w_meta = __ldcg(&P.st_meta[i]); w_off = __ldcg(&P.st_off[i]); }
        if (lane == 0 && i_base + 32 < F) { w_il2 = __ldcg(&P.item_loc[i_base + 32]); w_meta2 = __ldcg(&P.st_meta[i_base + 32]); }
      };
      window_load(i_cur);
      for (uint32_t t0 = wb; t0 < we; t0 += 32) {
        PROF_START();
        if (i_cur + lane < F) {
          win_rec[lane] = w_off;
          win_seg[lane] = s_pref_items[w_meta >> 8] + (w_il & ~kSideBit);
          win_meta[lane] = (w_meta & 0xFFu) | (w_il & kSideBit);
        } else {
          win_seg[lane] = T;
        }
        if (lane == 0) win_seg[32] = (i_cur + 32 < F) ? s_pref_items[w_meta2 >> 8] + (w_il2 & ~kSideBit) : T;
        __syncwarp();
        PROF_MARK(0);
        const uint32_t t = t0 + lane;
        uint32_t cnt_out = 0, next_note = 0;
        uint4 rec = make_uint4(0, 0, 0, 0);
        const bool valid = t < we;
        uint32_t k = 0;
        if (valid) {
          k = smem_segment(win_seg, 33, t);
          // the lane on the tile's last item knows which state holds the first item of the next tile
          next_note = i_cur + k + (win_seg[k + 1] <= t + 1 ? 1u : 0u);
        }
        const uint32_t i_next = __shfl_sync(0xFFFFFFFFu, next_note, 31);
        if (t0 + 32 < we) window_load(i_next);
        if (valid) {
          const uint32_t i = i_cur + k, j = t - win_seg[k];
          const uint32_t fl = win_meta[k];
          const bool match_input = (fl & kSideBit) != 0;
          const uint32_t fs = (fl >> 4) & 3u;
          const bool hs_searched = (fl & 64) != 0;
          FsFlags ff;
          ff.alleps1 = fl & 1; ff.noeps1 = fl & 2; ff.alleps2 = fl & 4; ff.noeps2 = fl & 8;
          const uint4 so = win_rec[k];
          // iterated side / searched side, selected without branching (lanes of a warp sit on both sides)
          const uint32_t* __restrict__ it_lab = match_input ? P.lab1 : P.lab2;
          const uint32_t* __restrict__ se_lab = match_input ? P.lab2 : P.lab1;
          const uint32_t it_lo = match_input ? so.x : so.z;
          const uint32_t se_lo = match_input ? so.z : so.x, se_hi = match_input ? so.w : so.y;
          Label label = kNoLabel;
          uint32_t it_idx = 0xFFFFFFFFu;  // absolute index of the iterated arc; all ones = implicit epsilon loop
          if (j != 0) { it_idx = it_lo + j - 1; label = __ldg(&it_lab[it_idx]); }
          PROF_USE(label);
          PROF_MARK(1);
          const bool has_loop = (label == kEps);
          const Label key = (label == kNoLabel) ? kEps : label;
          uint32_t pos, end;
          match_range(se_lab, se_lo, se_hi, key, has_loop, pos, end);
          uint32_t cnt = end - pos;
          PROF_USE(cnt);
          PROF_MARK(2);
          // filter_tr sees (arc1.olabel, arc2.ilabel): the iterated arc's label on its own side, the match on the other
          const uint32_t fs_loop = !has_loop ? kNoFs
                                   : filter_eval(P.kind, fs, ff, match_input ? label : kNoLabel, match_input ? kNoLabel : label);
          const uint32_t fs_real = filter_eval(P.kind, fs, ff, match_input ? label : key, match_input ? key : label);
          bool sigma_mode = false;
          if (P.sig1.enabled | P.sig2.enabled) {
            const SigmaDev& sg = match_input ? P.sig2 : P.sig1;  // matcher of the searched side
            if (sg.enabled) {  // IteratorSigmaMatcher::new (sigma_matcher.rs:196-246)
              if (label == sg.label && sg.label != kNoLabel) atomicOr(&P.ctl[4], (uint32_t)kErrBadSigmaLabel);
              if (!has_loop && cnt == 0 && hs_searched && label != kEps && label != kNoLabel &&
                  dev_sigma_allowed(sg, label)) {
                pos = lower_bound_lab(se_lab, se_lo, se_hi, sg.label);
                end = run_end_lab(se_lab, pos, se_hi, sg.label);
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
                                    ((fs_real & 3u) << 29) | (match_input ? 1u << 31 : 0u), i, it_idx);
        }
        PROF_MARK(3);
        // warp scans: active records by ballot, arcs by shuffles
        const bool act = cnt_out != 0;
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, act);
        const uint32_t ex_act = __popc(bal & ((1u << lane) - 1u));
        uint32_t inc = cnt_out;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, inc, o); if ((int)lane >= o) inc += u; }
        const uint32_t ex_arcs = inc - cnt_out;
        const uint32_t tile_arcs = __shfl_sync(0xFFFFFFFFu, inc, 31);
        PROF_MARK(4);
        if (valid) {
          if (rec.w == 0xFFFFFFFFu) P.st_arc_loc[rec.z] = w_arcs + ex_arcs;  // first arc of state rec.z (warp-local)
          if (act) {
            const uint32_t r = wb + w_active + ex_act;  // in-place compaction inside the warp's own item run
            P.recs[r] = rec;
            P.arc_loc[r] = w_arcs + ex_arcs;
          }
        }
        w_active += __popc(bal); w_arcs += tile_arcs;
        i_cur = i_next;
        __syncwarp();  // the window is overwritten at the top of the next tile
        PROF_MARK(5);
      }
    }
    // warp totals -> CTA total and the warps' exclusive arc offsets inside the CTA
    __syncthreads();
    if (lane == 0) s_warp[wid] = w_arcs;
    __syncthreads();
    uint32_t my_arcs = 0, w_pref = 0;
#pragma unroll
    for (uint32_t w = 0; w < kWarps; w++) { const uint32_t v = s_warp[w]; if (w < wid) w_pref += v; my_arcs += v; }
    if (lane == 0) P.wpref_arcs[c * kWarps + wid] = w_pref;  // read by the CTAs that own the states (CSR offsets)
    publish_count(P.part_arcs, c, tag, my_arcs);
    busy[1] += globaltimer_ns() - ts1;

    // ------------------------------------------------------------------ B: emit
    cta_prefix_wait_to_smem(P.part_arcs, G, tag, s_pref_arcs, s_warp, P.poll_ns);
    unsigned long long tp1 = globaltimer_ns();
    const uint32_t E = s_pref_arcs[G];
    overflow |= __ldcg(&P.ctl[4]);  // raised during A1 only, so every CTA reads the same value here
    if ((unsigned long long)base + E > P.arcs_cap) overflow |= kOvArcs;
    if (((unsigned long long)hi + E) * 2ull > P.table_cap) overflow |= kOvTable;
    const uint32_t arc_chunk = ((E + G - 1) / G + 31u) & ~31u;
    if (arc_chunk > (1u << kLocalRankBits)) overflow |= kOvChunk;
    if (overflow) break;  // uniform
    {
      // every warp emits the arcs of its own records, 32 arcs per tile, again without CTA barriers
      const uint32_t e_off = s_pref_arcs[c] + w_pref;  // canonical wave-local emission index of the warp's first arc
      Tr* __restrict__ wave_arcs = P.out_arcs + base;
      uint32_t* const win_seg = s_wseg[wid];
      uint4* const win_rec = s_wrec8[wid];
      // kEmitArcs arcs per lane and round: the gathers, and then the first table probes, of all the lane's arcs are
      // issued before any of them is waited for.  Measured on C3 (emit phase, 768-thread CTAs): 1 arc per lane 1.39 ms,
      // 2 arcs 1.47 ms; 512-thread CTAs: 2 arcs 1.63 ms, 4 arcs 1.74 ms — more requests in flight per warp make the
      // phase slower, not faster (it is bound by memory-request throughput, not by the dependent chain), so the
      // default stays 1.
      constexpr uint32_t kBT = 32u * kEmitArcs;
      uint32_t cursor = 0;  // first record (warp-local) that can contain the next round's first arc
      uint32_t w_loc[kEmitArcs], w_loc2 = 0;
      uint4 w_rec[kEmitArcs];
#pragma unroll
      for (int q = 0; q < kEmitArcs; q++) { w_loc[q] = 0; w_rec[q] = make_uint4(0, 0, 0, 0); }
      auto window_load = [&](uint32_t cur) {
#pragma unroll
        for (int q = 0; q < kEmitArcs; q++) {
          const uint32_t idx = cur + 32u * q + lane;
          if (idx < w_active) { w_loc[q] = P.arc_loc[wb + idx]; w_rec[q] = P.recs[wb + idx]; }
        }
        if (lane == 0 && cur + kBT < w_active) w_loc2 = P.arc_loc[wb + cur + kBT];
      };
      if (w_arcs) window_load(0);
      for (uint32_t e0 = 0; e0 < w_arcs; e0 += kBT) {
        PROF_START();
#pragma unroll
        for (int q = 0; q < kEmitArcs; q++) {
          const uint32_t idx = 32u * q + lane;
          if (cursor + idx < w_active) { win_seg[idx] = w_loc[q]; win_rec[idx] = w_rec[q]; }
          else win_seg[idx] = w_arcs;
        }
        if (lane == 0) win_seg[kBT] = (cursor + kBT < w_active) ? w_loc2 : w_arcs;
        __syncwarp();
        PROF_MARK(6);
        uint32_t el[kEmitArcs], k[kEmitArcs];
        bool valid[kEmitArcs];
#pragma unroll
        for (int q = 0; q < kEmitArcs; q++) {
          el[q] = e0 + 32u * q + lane;
          valid[q] = el[q] < w_arcs;
          k[q] = valid[q] ? smem_segment(win_seg, kBT + 1, el[q]) : 0u;
        }
        // the lane on the round's last arc knows which record holds the first arc of the next round
        uint32_t next_note = 0;
        if (valid[kEmitArcs - 1])
          next_note = cursor + k[kEmitArcs - 1] + (win_seg[k[kEmitArcs - 1] + 1] <= el[kEmitArcs - 1] + 1 ? 1u : 0u);
        const uint32_t cursor_next = __shfl_sync(0xFFFFFFFFu, next_note, 31);
        if (e0 + kBT < w_arcs) window_load(cursor_next);
        Tr out[kEmitArcs];
        unsigned long long key[kEmitArcs];
        uint32_t h[kEmitArcs];
        uint4 sv[kEmitArcs];
#pragma unroll
        for (int q = 0; q < kEmitArcs; q++) {
          if (valid[q]) {
            const uint4 rec = win_rec[k[q]];
            const uint32_t kk = el[q] - win_seg[k[q]];
            const bool loop_ok = (rec.y >> 26) & 1u;
            const bool match_input = rec.y >> 31;
            const bool it_is_loop = rec.w == 0xFFFFFFFFu, cand_is_loop = loop_ok && kk == 0;
            const Tr* __restrict__ it_arcs = match_input ? P.a.arcs : P.b.arcs;
            const Tr* __restrict__ cd_arcs = match_input ? P.b.arcs : P.a.arcs;
            uint32_t s1 = 0, s2 = 0;
            if (it_is_loop || cand_is_loop) { uint32_t fs; unpack_key(__ldcg(&P.tuples[lo + rec.z]), fs, s1, s2); }
            // implicit epsilon loops (matcher.rs: eps_loop): (0, NO_LABEL) / (NO_LABEL, 0) staying in the same state
            Tr it = match_input ? Tr{kEps, kNoLabel, 0.0f, s1} : Tr{kNoLabel, kEps, 0.0f, s2};
            Tr cand = match_input ? Tr{kNoLabel, kEps, 0.0f, s2} : Tr{kEps, kNoLabel, 0.0f, s1};
            if (!it_is_loop) it = load_tr(&it_arcs[rec.w]);
            if (!cand_is_loop) cand = load_tr(&cd_arcs[rec.x + kk - (loop_ok ? 1u : 0u)]);
            uint32_t fsn = (rec.y >> 27) & 3u;
            if (!cand_is_loop) {
              fsn = (rec.y >> 29) & 3u;
              if ((rec.y >> 25) & 1u) {  // sigma match: relabel (value_openfst, sigma_matcher.rs:249-276)
                const SigmaDev& sg = match_input ? P.sig2 : P.sig1;
                const Label l = match_input ? it.olabel : it.ilabel;
                if (sg.rewrite_both) { if (cand.ilabel == sg.label) cand.ilabel = l; if (cand.olabel == sg.label) cand.olabel = l; }
                else if (match_input) cand.ilabel = l;
                else cand.olabel = l;
              }
            }
            out[q].ilabel = match_input ? it.ilabel : cand.ilabel;   // arc1 = the fst1 arc, arc2 = the fst2 arc
            out[q].olabel = match_input ? cand.olabel : it.olabel;
            out[q].weight = w_times(it.weight, cand.weight);
            key[q] = pack_key(fsn, match_input ? it.nextstate : cand.nextstate,
                              match_input ? cand.nextstate : it.nextstate);
            h[q] = hash_key(key[q]) & P.mask;
            // one slot per probe round: wider rounds (2 / 4 slots fetched together) were measured slower on C3
            sv[q] = ld_volatile_u4(&P.slots[h[q]]);
          }
        }
        PROF_MARK(7);
#pragma unroll
        for (int q = 0; q < kEmitArcs; q++) {
          if (valid[q]) {
            const uint32_t e = e_off + el[q];  // canonical wave-local emission index
            const uint32_t id = table_probe_from(P.slots, P.mask, key[q], h[q], sv[q]);
            if (id != kUnassigned) out[q].nextstate = id;
            else { atomicMin(&P.slots[h[q]].emin, e); out[q].nextstate = kPendingBit | h[q]; }
            store_tr(&wave_arcs[e], out[q]);
          }
        }
        PROF_MARK(8);
        cursor = cursor_next;
        __syncwarp();
        PROF_MARK(9);
      }
      // state -> first arc (CSR offsets of the result) for the states of my slice: the warp that processed item 0 of
      // the state recorded the warp-local emission index of its first arc
      for (uint32_t i = s_pref_new[c] + tid; i < s_pref_new[c + 1]; i += kCoopThreads) {
        const uint32_t t_first = s_pref_items[c] + (__ldcg(&P.item_loc[i]) & ~kSideBit);
        const uint32_t cw = t_first / wc;  // global warp index = CTA * 8 + warp
        P.out_offsets[lo + i] = base + s_pref_arcs[cw / kWarps] + __ldcg(&P.wpref_arcs[cw]) + __ldcg(&P.st_arc_loc[i]);
      }
    }
    busy[2] += globaltimer_ns() - tp1;
    grid_barrier(P.barrier, bar_epoch);
    unsigned long long tp2 = globaltimer_ns();

    // ------------------------------------------------------------------ C: rank first emissions
    // Four consecutive arcs per thread and round (1024 arcs per CTA round): the loads of a round fly together and one
    // CTA scan serves four arcs; a chunk is usually a single round.
    const uint32_t e_begin = min(E, c * arc_chunk), e_end = min(E, e_begin + arc_chunk);
    uint32_t cta_new = 0;
    {
      const Tr* __restrict__ wave_arcs = P.out_arcs + base;
      for (uint32_t r0 = e_begin; r0 < e_end; r0 += kArcsPerThread * kCoopThreads) {
        const uint32_t eb = r0 + kArcsPerThread * tid;
        uint32_t ns[kArcsPerThread], em[kArcsPerThread];
#pragma unroll
        for (int q = 0; q < kArcsPerThread; q++) ns[q] = (eb + q < e_end) ? __ldcg(&wave_arcs[eb + q].nextstate) : 0u;
#pragma unroll
        for (int q = 0; q < kArcsPerThread; q++)
          em[q] = (ns[q] & kPendingBit) ? __ldcg(&P.slots[ns[q] & ~kPendingBit].emin) : 0xFFFFFFFFu;
        uint32_t cnt = 0;
#pragma unroll
        for (int q = 0; q < kArcsPerThread; q++) cnt += (em[q] == eb + q) ? 1u : 0u;  // first emitter of a new tuple
        uint32_t round_total;
        uint32_t r = cta_new + cta_exclusive_scan(cnt, s_warp, round_total);
#pragma unroll
        for (int q = 0; q < kArcsPerThread; q++)
          if (em[q] == eb + q) { P.slots[ns[q] & ~kPendingBit].id = kTempFlag | (c << kLocalRankBits) | r; r++; }
        cta_new += round_total;
      }
    }
    tag++;
    publish_count(P.part_new, c, tag, cta_new);
    busy[3] += globaltimer_ns() - tp2;

    // ------------------------------------------------------------------ D: resolve + setup of the next frontier
    cta_prefix_wait_to_smem(P.part_new, G, tag, s_pref_new, s_warp, P.poll_ns);
    unsigned long long tp3 = globaltimer_ns();
    const uint32_t n_new = s_pref_new[G];
    if ((unsigned long long)hi + n_new > P.states_cap || (unsigned long long)hi + n_new >= 0x7FFFFFFFull) {
      overflow |= kOvStates;
      break;  // uniform
    }
    {
      Tr* __restrict__ wave_arcs = P.out_arcs + base;
      uint32_t run = 0;
      for (uint32_t r0 = e_begin; r0 < e_end; r0 += kArcsPerThread * kCoopThreads) {
        const uint32_t eb = r0 + kArcsPerThread * tid;
        uint32_t ns[kArcsPerThread], nit[kArcsPerThread], inew[kArcsPerThread];
        uint4 sv[kArcsPerThread];
#pragma unroll
        for (int q = 0; q < kArcsPerThread; q++) ns[q] = (eb + q < e_end) ? __ldcg(&wave_arcs[eb + q].nextstate) : 0u;
#pragma unroll
        for (int q = 0; q < kArcsPerThread; q++)
          sv[q] = (ns[q] & kPendingBit) ? ld_volatile_u4(&P.slots[ns[q] & ~kPendingBit]) : make_uint4(0, 0, 0, 0xFFFFFFFFu);
        uint32_t sum = 0;
#pragma unroll
        for (int q = 0; q < kArcsPerThread; q++) {
          nit[q] = 0; inew[q] = 0;
          if (ns[q] & kPendingBit) {
            const uint32_t v = sv[q].z;
            uint32_t id = v;
            if (v & kTempFlag) id = hi + s_pref_new[(v & ~kTempFlag) >> kLocalRankBits] + (v & ((1u << kLocalRankBits) - 1u));
            wave_arcs[eb + q].nextstate = id;
            if (sv[q].w == eb + q) {  // first emitter: publish, and set the state up for the next wave
              P.slots[ns[q] & ~kPendingBit].id = id;
              const unsigned long long key = (unsigned long long)sv[q].x | ((unsigned long long)sv[q].y << 32);
              P.tuples[id] = key;
              inew[q] = id - hi;
              nit[q] = setup_state(P, key, inew[q], id, c);
              sum += nit[q] & ~kSideBit;
            }
          }
        }
        uint32_t round_total;
        uint32_t off = run + cta_exclusive_scan(sum, s_warp, round_total);
#pragma unroll
        for (int q = 0; q < kArcsPerThread; q++)
          if (nit[q]) { P.item_loc[inew[q]] = off | (nit[q] & kSideBit); off += nit[q] & ~kSideBit; }
        run += round_total;
      }
      publish_count(P.part_items, c, tag, run);
    }
    n_states_exp += F; n_items += T; n_arcs += E; n_waves++;
    base += E;
    lo = hi;
    hi += n_new;
    unsigned long long tp4 = globaltimer_ns();
    busy[4] += tp4 - tp3;
    t_a += tp1 - tp0; t_b += tp2 - tp1; t_c += tp3 - tp2; t_d += tp4 - tp3;
  }

  if (tid == 0) {
    for (int k = 0; k < 5; k++) { atomicAdd(&P.stats[8 + k], busy[k]); atomicMax(&P.stats[13 + k], busy[k]); }
    if (P.cta_trace) {
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      P.cta_trace[5 * c] = smid;
      for (int k = 1; k < 5; k++) P.cta_trace[5 * c + k] = busy[k];
    }
#ifdef B200_COOP_PROFILE
    for (int k = 0; k < 12; k++) atomicAdd(&P.stats[18 + k], prof[k]);
#endif
  }
  if (c == 0 && tid == 0) {
    P.ctl[1] = overflow;
    P.ctl[2] = hi;    // number of product states
    P.ctl[3] = base;  // number of arcs
    P.out_offsets[hi] = base;
    P.wave_lo[n_waves] = hi;
    P.stats[0] = n_states_exp; P.stats[1] = n_items - n_states_exp; P.stats[2] = n_arcs; P.stats[3] = n_waves;
    P.stats[4] = t_a; P.stats[5] = t_b; P.stats[6] = t_c; P.stats[7] = t_d;
  }
}

// Seeds the state table with the start tuples: id i <- (start_fs, starts1[i], start2).  The s1 components are
// pairwise distinct (different acceptors of a batch live in disjoint state ranges), so are the keys.
__global__ void k_coop_init(Slot* slots, uint32_t mask, unsigned long long* tuples, uint32_t start_fs,
                            const uint32_t* __restrict__ starts1, uint32_t single_start1, uint32_t start2,
                            uint32_t n_starts, uint32_t* ctl) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { for (int k = 0; k < 8; k++) ctl[k] = 0; }
  if (i >= n_starts) return;
  const unsigned long long key = pack_key(start_fs, starts1 ? starts1[i] : single_start1, start2);
  uint32_t h = hash_key(key) & mask;
  while (atomicCAS(&slots[h].key, kEmptyKey, key) != kEmptyKey) h = (h + 1) & mask;
  slots[h].id = i; slots[h].emin = 0;
  tuples[i] = key;
}

__global__ void k_unpack_s1(const unsigned long long* __restrict__ tuples, uint32_t n, uint32_t* __restrict__ s1_out,
                            uint32_t n_starts, uint32_t* __restrict__ start_map) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { uint32_t fs, s1, s2; unpack_key(tuples[i], fs, s1, s2); s1_out[i] = s1; }
  if (i < n_starts) start_map[i] = i;
}
static int grid_used = 1;
float run_coop(const CoopParams& P0, int sms, cudaStream_t s) {
  CoopParams P = P0;
  // resident CTAs per SM (only the 256-thread build has a choice; see coop_utils.cuh for the measured shapes)
  int minb = 3;
  if (const char* e = std::getenv("B200_COOP_MINBLOCKS")) minb = std::atoi(e);
#if B200_COOP_THREADS > 512
  minb = 1;
  void* kern = (void*)k_compose_coop<1>;
#elif B200_COOP_THREADS > 256
  if (!std::getenv("B200_COOP_MINBLOCKS")) minb = 1;
  if (minb > 3) minb = 3;
  void* kern = (void*)k_compose_coop<2>;
  if (minb == 3) kern = (void*)k_compose_coop<3>;
  else if (minb == 1) kern = (void*)k_compose_coop<1>;
#else
  void* kern = (void*)k_compose_coop<3>;
  if (minb == 5) kern = (void*)k_compose_coop<5>;
  else if (minb == 6) kern = (void*)k_compose_coop<6>;
  else if (minb == 4) kern = (void*)k_compose_coop<4>;
  else if (minb == 2) kern = (void*)k_compose_coop<2>;
  else if (minb == 1) kern = (void*)k_compose_coop<1>;
  else if (minb == 8) kern = (void*)k_compose_coop<8>;
#endif
  int per_sm = 0;
  size_t dyn = 3 * ((size_t)std::min(2047, sms * minb) + 1) * sizeof(uint32_t);
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kCoopThreads, dyn));
  if (per_sm < 1) throw FstError("cooperative compose kernel does not fit on the device");
  if (per_sm > minb) per_sm = minb;
  int grid = sms * per_sm;
  if (grid > 2047) grid = 2047;  // CTA index must fit 11 bits next to the 20-bit local rank; prefix arrays hold 2048
  grid_used = grid;
  dyn = 3 * ((size_t)grid + 1) * sizeof(uint32_t);
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

}  // namespace

const uint32_t* label_column(const DevFst& f, bool olabel, uint32_t pad, cudaStream_t s, bool* built) {
  DevFst& m = const_cast<DevFst&>(f);  // the columns are a cache of derived data
  if (built) *built = false;
  {
    static std::mutex g_create;
    std::lock_guard<std::mutex> g(g_create);
    if (!m.columns) m.columns = std::make_shared<DevLabelColumns>();
  }
  DevLabelColumns& c = *m.columns;
  std::lock_guard<std::mutex> g(c.mu);
  DevBuf<uint32_t>& buf = olabel ? c.olab : c.ilab;
  bool& has = olabel ? c.has_olab : c.has_ilab;
  const cudaStream_t own = f.arcs.s;  // the machine's own stream: the column lives and dies with the machine
  if (!has) {
    if (!c.ready) B200_CUDA(cudaEventCreateWithFlags(&c.ready, cudaEventDisableTiming));
    if (own != s) {  // the arcs may still be in flight on the caller's stream (upload on s, column on own)
      cudaEvent_t e;
      B200_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      B200_CUDA(cudaEventRecord(e, s));
      B200_CUDA(cudaStreamWaitEvent(own, e, 0));
      cudaEventDestroy(e);
    }
    buf = DevBuf<uint32_t>(own, (size_t)f.num_arcs + pad);
    composeimpl::launch_extract_labels(f.arcs.p, f.num_arcs, f.num_arcs + pad, olabel ? 1 : 0, buf.p, own);
    B200_CUDA(cudaEventRecord(c.ready, own));
    has = true;
    if (built) *built = true;
  }
  if (own != s) B200_CUDA(cudaStreamWaitEvent(s, c.ready, 0));
  return buf.p;
}

namespace composeimpl {
void launch_unpack_s1(const unsigned long long* tuples, uint32_t n, uint32_t* s1_out, uint32_t n_starts,
                      uint32_t* start_map, cudaStream_t s) {
  uint32_t m = n > n_starts ? n : n_starts;
  if (m) k_unpack_s1<<<blocks_for(m), kThreads, 0, s>>>(tuples, n, s1_out, n_starts, start_map);
}

void launch_extract_labels(const Tr* arcs, uint32_t n, uint32_t n_padded, int olabel, uint32_t* out, cudaStream_t s) {
  k_extract_labels<<<blocks_for(n_padded), kThreads, 0, s>>>(arcs, n, n_padded, olabel, out);
}
}  // namespace composeimpl

bool compose_device_coop(const DevFst& fa, const DevFst& fb, const ComposeOptions& opt, ComposeStats* stats,
                         cudaStream_t s, DevFst* result, const BatchStarts* batch) {
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
  if ((opt.sigma1.enabled || opt.sigma2.enabled) && !batch && (!fa.has_start || !fb.has_start)) {
    // empty result; the multi-kernel back end (which handles the no-start case) does not take sigma configs
    ComposeOptions plain = opt; plain.sigma1 = SigmaSpec(); plain.sigma2 = SigmaSpec();
    *result = compose_device_waves(fa, fb, plain, stats, s);
    return true;
  }
  if (!batch && (!fa.has_start || !fb.has_start)) return false;  // trivial case: multi-kernel back end
  if (batch && (!fb.has_start || batch->n == 0)) throw FstError("batched compose needs a start state on both sides");
  const uint32_t n_starts = batch ? batch->n : 1u;

  ComposeStats local;
  ComposeStats& st = stats ? *stats : local;
  st = ComposeStats();
  cudaEvent_t ev0, ev1, ev2;
  B200_CUDA(cudaEventCreate(&ev0)); B200_CUDA(cudaEventCreate(&ev1)); B200_CUDA(cudaEventCreate(&ev2));
  B200_CUDA(cudaEventRecord(ev0, s));

  DevBuf<uint32_t> neps1(s), neps2(s);
  bool need_eps = (kind == kSequenceFilter || kind == kAltSequenceFilter || kind == kMatchFilter);
  CoopParams P{};
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
  DevBuf<uint32_t> lab1(s, (size_t)fa.num_arcs + kLabelPad), lab2(s, (size_t)fb.num_arcs + kLabelPad);
  launch_extract_labels(fa.arcs.p, fa.num_arcs, fa.num_arcs + kLabelPad, 1, lab1.p, s);
  launch_extract_labels(fb.arcs.p, fb.num_arcs, fb.num_arcs + kLabelPad, 0, lab2.p, s);
  P.lab1 = lab1.p; P.lab2 = lab2.p; st.kernel_launches += 2;
  DevBuf<uint32_t> allowed1(s), allowed2(s);
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
    }
    return d;
  };
  P.sig1 = mk_sigma(opt.sigma1, fa.props, allowed1);
  P.sig2 = mk_sigma(opt.sigma2, fb.props, allowed2);
  B200_CUDA(cudaStreamSynchronize(s));  // the allowed lists are host temporaries

  // ---- pre-sized buffers (HBM is plentiful: 180 GB); an overflow falls back to the growing back end
  const size_t sum_states = (size_t)fa.num_states + fb.num_states, sum_arcs = (size_t)fa.num_arcs + fb.num_arcs;
  size_t states_cap = std::max<size_t>(1 << 16, 8 * sum_states + 2 * (size_t)n_starts);
  size_t arcs_cap = std::max<size_t>(1 << 18, 4 * sum_arcs);
  states_cap = std::min<size_t>(states_cap, 0x7FFFFFF0ull);
  arcs_cap = std::min<size_t>(arcs_cap, 0xFFFFFFF0ull);
  size_t table_cap = 1 << 17;
  while (table_cap < 2 * states_cap && table_cap < (1ull << 30)) table_cap <<= 1;
  size_t items_cap = std::max<size_t>(1 << 18, arcs_cap / 2);

  DevFst out(s);
  out.offsets.reserve_discard(states_cap + 1);
  out.finals.reserve_discard(states_cap);
  out.arcs.reserve_discard(arcs_cap);
  DevBuf<unsigned long long> tuples(s, states_cap), dstats(s, 30);
  DevBuf<Slot> slots(s, table_cap);
  DevBuf<uint4> recs(s, items_cap);
  DevBuf<uint32_t> arc_loc(s, items_cap), item_loc(s, states_cap), st_arc_loc(s, states_cap), st_meta(s, states_cap), ctl(s, 8);
  DevBuf<unsigned long long> parts(s, 3 * 2048);
  DevBuf<uint32_t> wpref_arcs(s, 2048 * 8);
  DevBuf<uint4> st_off(s, states_cap);
  const uint32_t wave_cap = 1u << 20;
  DevBuf<uint32_t> wave_lo(s, wave_cap);
  B200_CUDA(cudaMemsetAsync(slots.p, 0xFF, table_cap * sizeof(Slot), s));
  B200_CUDA(cudaMemsetAsync(dstats.p, 0, 30 * sizeof(unsigned long long), s));
  B200_CUDA(cudaMemsetAsync(parts.p, 0, 3 * 2048 * sizeof(unsigned long long), s));  // tag 0 = nothing published yet
  P.tuples = tuples.p; P.states_cap = (uint32_t)states_cap;
  P.out_offsets = out.offsets.p; P.out_finals = out.finals.p; P.out_arcs = out.arcs.p; P.arcs_cap = (uint32_t)arcs_cap;
  P.slots = slots.p; P.mask = (uint32_t)table_cap - 1; P.table_cap = (uint32_t)table_cap;
  P.item_loc = item_loc.p; P.st_arc_loc = st_arc_loc.p; P.st_meta = st_meta.p; P.st_off = st_off.p;
  P.recs = recs.p; P.arc_loc = arc_loc.p; P.wpref_arcs = wpref_arcs.p; P.items_cap = (uint32_t)std::min<size_t>(items_cap, 0xFFFFFFF0ull);
  P.part_arcs = parts.p; P.part_items = parts.p + 2048; P.part_new = parts.p + 4096;
  P.ctl = ctl.p; P.stats = dstats.p;
  P.barrier = ctl.p + 6;
  P.wave_lo = wave_lo.p; P.wave_cap = wave_cap;
  uint32_t start_fs = (kind == kNullFilter || kind == kTrivialFilter || kind == kNoMatchFilter) ? 1u : 0u;
  P.n_starts = n_starts;
  if (const char* e = std::getenv("B200_POLL_NS")) P.poll_ns = (uint32_t)std::atoi(e);
  k_coop_init<<<blocks_for(n_starts), kThreads, 0, s>>>(slots.p, P.mask, tuples.p, start_fs,
                                                        batch ? batch->d_starts1 : nullptr, fa.start, fb.start, n_starts, ctl.p);
  st.kernel_launches++;

  DevBuf<unsigned long long> cta_trace(s);
  const bool trace = std::getenv("B200_COOP_TRACE") != nullptr;
  if (trace) { cta_trace.reserve_discard(5 * 2048); P.cta_trace = cta_trace.p; }
  float ms_kernel = run_coop(P, sm_count(), s);
  st.kernel_launches++; st.emit_launches = 1;

  uint32_t hctl[4];
  unsigned long long hstats[30];
  B200_CUDA(cudaMemcpyAsync(hctl, ctl.p, 16, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaMemcpyAsync(hstats, dstats.p, 30 * 8, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  if (hctl[1] != 0) {
    cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev2);
    if (hctl[1] & kErrBothRequire) throw FstError("Both sides can't require match");           // compose_fst_op.rs:207-209
    if (hctl[1] & kErrBadSigmaLabel) throw FstError("SigmaMatcher::Find: bad label (sigma)");  // sigma_matcher.rs:205-207
    return false;  // a pre-sized buffer was too small
  }
  st.states_expanded = hstats[0]; st.arcs_iterated = hstats[1]; st.arcs_emitted = hstats[2]; st.waves = hstats[3];
  st.ms_emit_kernel = ms_kernel;
  st.ms_phase[0] = hstats[4] * 1e-6f; st.ms_phase[1] = hstats[5] * 1e-6f;
  st.ms_phase[2] = hstats[6] * 1e-6f; st.ms_phase[3] = hstats[7] * 1e-6f;
  if (std::getenv("B200_COOP_TRACE")) {
    const char* names[5] = {"init", "A1", "B", "C", "D+setup"};
    for (int k = 0; k < 5; k++)
      std::fprintf(stderr, "[coop] phase %s: CTA busy avg %.3f ms, max %.3f ms\n", names[k],
                   hstats[8 + k] * 1e-6 / grid_used, hstats[13 + k] * 1e-6);
    {
      // The warp schedulers favour the oldest warps, so on every SM the CTA that was launched first finishes a phase
      // first and the youngest last: per-CTA busy time grows with the CTA index although the work is equal.  The time
      // an SM needs for a phase is the busy time of its slowest CTA.
      std::vector<unsigned long long> tr(5 * (size_t)grid_used);
      B200_CUDA(cudaMemcpy(tr.data(), cta_trace.p, tr.size() * 8, cudaMemcpyDeviceToHost));
      for (int ph = 1; ph <= 4; ph++) {
        std::vector<double> sm_max(256, 0.0);
        for (int c = 0; c < grid_used; c++) sm_max[tr[5 * c] & 255] = std::max(sm_max[tr[5 * c] & 255], tr[5 * c + ph] * 1e-6);
        double sum = 0, mx = 0; int n = 0;
        for (int m = 0; m < 256; m++) if (sm_max[m] > 0) { sum += sm_max[m]; mx = std::max(mx, sm_max[m]); n++; }
        std::fprintf(stderr, "[coop] phase %s: slowest CTA per SM avg %.3f ms, max %.3f ms over %d SMs\n", names[ph], sum / n, mx, n);
      }
    }
#ifdef B200_COOP_PROFILE
    const char* pn[12] = {"A1 window fill", "A1 segment+label", "A1 match_range", "A1 filter", "A1 scan", "A1 store+advance",
                          "B window fill", "B segment+gathers", "B table probe", "B store+advance", "-", "-"};
    for (int k = 0; k < 10; k++)
      std::fprintf(stderr, "[coop-prof] %-18s %.1f kcycles per CTA\n", pn[k], hstats[18 + k] * 1e-3 / grid_used);
#endif
  }
  out.num_states = hctl[2]; out.num_arcs = hctl[3];
  out.has_start = true; out.start = 0;
  out.props = props::of_compose(fa.props, fb.props);
  B200_CUDA(cudaEventRecord(ev1, s));
  if (opt.connect) {
    uint64_t launches = 0;
    TrimExtras extras;
    extras.tuples = batch ? tuples.p : nullptr;
    extras.n_starts = n_starts;
    extras.out_tag = batch ? batch->out_s1 : nullptr;
    extras.out_start_map = batch ? batch->out_start_map : nullptr;
    DevFst trimmed = connect_waves_device(out, wave_lo.p, (uint32_t)st.waves, &launches, s, &extras);
    st.kernel_launches += launches;
    out = std::move(trimmed);
  }
  else if (batch) {
    // untrimmed batch result: s1 of every product state + identity start map
    batch->out_s1->reserve_discard(out.num_states ? out.num_states : 1);
    batch->out_start_map->reserve_discard(n_starts);
    launch_unpack_s1(tuples.p, out.num_states, batch->out_s1->p, n_starts, batch->out_start_map->p, s);
    st.kernel_launches++;
  }
  st.states_out = out.num_states; st.arcs_out = out.num_arcs;
  B200_CUDA(cudaEventRecord(ev2, s));
  B200_CUDA(cudaStreamSynchronize(s));
  B200_CUDA(cudaEventElapsedTime(&st.ms_expand, ev0, ev1));
  B200_CUDA(cudaEventElapsedTime(&st.ms_connect, ev1, ev2));
  cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev2);
  *result = std::move(out);
  return true;
}

// Persistent back ends with growth: the warp-stream kernel (compose_ws.cu) is run with capacities derived from the
// operands; whatever overflowed is doubled and the call repeated, so results larger than the first guess (including
// sigma-matcher compositions, which only the persistent kernels implement) are still computed.  Returns false only
// when the result cannot be represented by the kernel at all (the caller then uses the multi-kernel back end).
bool compose_device_persistent(const DevFst& a, const DevFst& b, const ComposeOptions& opt, ComposeStats* stats,
                               cudaStream_t s, DevFst* out, const BatchStarts* batch) {
  DeviceExclusive excl(device_exclusive());  // the persistent kernels want every SM (device_common.cu)
  const char* impl = std::getenv("B200_COMPOSE_IMPL");
  if (impl && std::string(impl) == "coop") return compose_device_coop(a, b, opt, stats, s, out, batch);
  WsCaps caps;
  for (int attempt = 0; attempt < 40; attempt++) {
    const int rc = compose_device_ws(a, b, opt, stats, s, out, batch, &caps);
    if (rc == 0) return true;
    if (rc & kOvChunk) return false;  // one warp would emit >= 2^20 arcs in one wave
    bool grown = false;
    auto grow = [&](size_t& v, size_t limit) { if (v < limit) { v = std::min(limit, v * 2); grown = true; } };
    if (rc & kOvArcs) grow(caps.arcs, 0xFFFFFFF0ull);
    if (rc & (kOvStates | kOvTable)) grow(caps.states, 0x7FFFFFF0ull);
    if (rc & kOvScratch) grow(caps.items, 0xFFFFFF00ull);
    if (rc & kOvRuns) grow(caps.runs, 0xFFFFFFF0ull);
    if (rc & kOvWaves) grow(caps.waves, 0x7FFFFFF0ull);
    if (!grown) return false;
    // A product that keeps outgrowing its buffers would end in an allocation failure somewhere inside the next attempt;
    // say what happened instead.  Working set of an attempt: 20 B per provisional arc, 20 B per match item, ~72 B per
    // state (records, tuples, first-arc index, finals) + a 16-B table slot for at least two slots per state.
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b) {
      const double need = 20.0 * (double)caps.arcs + 20.0 * (double)caps.items + (72.0 + 64.0) * (double)caps.states;
      if (need > 0.9 * (double)total_b)
        throw FstError("compose: the composition outgrew " + std::to_string((unsigned long long)(need / 1e9)) +
                       " GB of working memory (" + std::to_string((unsigned long long)caps.states) + " states, " +
                       std::to_string((unsigned long long)caps.arcs) + " transitions reserved) on a device with " +
                       std::to_string((unsigned long long)(total_b / 1e9)) + " GB; the product is too large for one GPU");
    }
  }
  return false;
}

DevFst compose_device(const DevFst& a, const DevFst& b, const ComposeOptions& opt, ComposeStats* stats,
                      cudaStream_t s) {
  const char* impl = std::getenv("B200_COMPOSE_IMPL");
  bool want_waves = impl && std::string(impl) == "waves";
  if (!want_waves) {
    DevFst out(s);
    if (compose_device_persistent(a, b, opt, stats, s, &out, nullptr)) return out;
  }
  return compose_device_waves(a, b, opt, stats, s);
}

}  // namespace b200
