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

#include "compose_common.cuh"
#include "coop_utils.cuh"

namespace cg = cooperative_groups;

namespace b200 {
namespace {
using namespace composeimpl;
using namespace coop;

constexpr uint32_t kTempFlag = 0x80000000u;   // slot.id holds (cta << 20 | local rank) between phases C and D
constexpr uint32_t kLocalRankBits = 20;

enum Overflow : uint32_t { kOvArcs = 1, kOvStates = 2, kOvTable = 4, kOvScratch = 8, kOvChunk = 16, kOvWaves = 32,
                           kErrBothRequire = 0x100, kErrBadSigmaLabel = 0x200 };

// Device view of one SigmaMatcher (sigma_matcher.rs): arcs labelled sigma_label on the matched side match any
// (allowed) label that has no ordinary match at the state; the matched arc is relabelled.
struct SigmaDev {
  uint32_t enabled; uint32_t label; uint32_t rewrite_both; const uint32_t* allowed; uint32_t n_allowed;
};

struct CoopParams {
  FstView a, b;
  int kind, side;
  unsigned long long* tuples; uint32_t states_cap;
  uint32_t* out_offsets; float* out_finals;
  Tr* out_arcs; uint32_t arcs_cap;
  Slot* slots; uint32_t mask; uint32_t table_cap;
  // per-wave scratch, indexed by frontier-local state
  uint32_t* item_loc;     // slice-local exclusive item offset | side bit
  uint32_t* st_arc_loc;   // slice-local emission index of the state's first arc
  uint8_t* st_flags;      // alleps1 | noeps1<<1 | alleps2<<2 | noeps2<<3 | fs<<4
  uint4* st_off;          // arc ranges of the two component states: alo, ahi, blo, bhi
  // per-wave scratch, indexed by item (active items compacted in place inside each CTA's slice)
  uint4* recs;            // x = first match, y = count|flags, z = frontier-local state, w = iterated arc (abs) or ~0
  uint32_t* arc_loc;      // slice-local emission index of the record's first arc
  uint32_t items_cap;
  uint32_t* part_items;   // gridDim entries: items of each CTA's state slice
  uint32_t* part_arcs;    // gridDim entries: arcs emitted by each CTA's item slice
  uint32_t* part_new;     // gridDim entries: first emissions found in each CTA's arc slice
  uint32_t* ctl;          // [1] overflow flags, [2] #states, [3] #arcs
  uint32_t* wave_lo; uint32_t wave_cap;  // first product id of every BFS wave (+ one-past-the-end sentinel)
  unsigned int* barrier;  // arrival counter of grid_barrier (zero-initialised)
  SigmaDev sig1, sig2;    // sigma matcher on fst1 (olabel side) / fst2 (ilabel side)
  uint32_t n_starts;      // initial frontier = product ids [0, n_starts) (1 for a plain compose, batch size otherwise)
  unsigned long long* stats;  // states_expanded, arcs_iterated, arcs_emitted, waves, ns phase A, B, C, D
};


// does state [lo, hi) of the matched side carry an arc labelled sigma? (has_sigma, sigma_matcher.rs:33-45)
template <bool kByOlabel>
__device__ __forceinline__ bool dev_has_sigma(const SigmaDev& sg, const Tr* arcs, uint32_t lo, uint32_t hi) {
  if (!sg.enabled || sg.label == kNoLabel) return false;
  const uint32_t p = lower_bound_label<kByOlabel>(arcs, lo, hi, sg.label);
  if (p >= hi) return false;
  return (kByOlabel ? __ldg(&arcs[p].olabel) : __ldg(&arcs[p].ilabel)) == sg.label;
}
__device__ __forceinline__ bool dev_sigma_allowed(const SigmaDev& sg, Label l) {
  if (!sg.n_allowed) return true;
  uint32_t lo = 0, hi = sg.n_allowed;
  while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (__ldg(&sg.allowed[mid]) < l) lo = mid + 1; else hi = mid; }
  return lo < sg.n_allowed && __ldg(&sg.allowed[lo]) == l;
}

constexpr uint32_t kSideBit = 0x80000000u;
constexpr uint32_t kTile = kCoopThreads;

// One persistent kernel = the whole BFS.  Per wave:
//   A0 states   : CTA c owns a contiguous slice of the frontier; per state: #items (1 + degree of the iterated side),
//                 which side is iterated, filter flags, final weight; CTA-local exclusive offsets + CTA total
//   A1 items    : CTA c owns a contiguous slice of the ITEMS (load-balanced: states located through a 257-entry
//                 shared-memory window per 256-item tile); per item: binary search of the sorted side + filter;
//                 active items are compacted in place with CTA-local arc offsets; CTA total
//   B  arcs     : CTA c emits the arcs of its own items, one thread per ARC (records located through a shared-
//                 memory window again); 128-bit gathers, CAS insert, atomicMin(first emission), 128-bit store
//   C  rank     : contiguous arc slices; ballot/popc ranks of first emissions published through slot.id
//   D  resolve  : CTA prefix -> ids; nextstate patch; tuple publication
// Five grid barriers per wave; every size is recomputed identically by every CTA from per-CTA partial arrays.
template <int kMinBlocks>
__global__ void __launch_bounds__(kCoopThreads, kMinBlocks)
k_compose_coop(CoopParams P) {
  cg::grid_group grid = cg::this_grid();
  __shared__ uint32_t s_warp[2 * (kCoopThreads / 32)];
  __shared__ uint32_t s_seg[kTile + 2];
  __shared__ uint32_t s_side[kTile + 2];
  __shared__ uint32_t s_misc[4];
  extern __shared__ uint32_t s_dyn[];          // two prefix arrays of gridDim + 1 entries
  uint32_t* s_pref_a = s_dyn;                  // items (A1/B) then new-state counts (D)
  uint32_t* s_pref_b = s_dyn + gridDim.x + 1;  // arcs (B)

  const uint32_t G = gridDim.x, c = blockIdx.x, tid = threadIdx.x;
  unsigned int bar_epoch = 0;
  uint32_t lo = 0, hi = P.n_starts, base = 0;  // uniform across the grid by construction
  unsigned long long n_states_exp = 0, n_items = 0, n_arcs = 0, n_waves = 0;
  unsigned long long t_a = 0, t_b = 0, t_c = 0, t_d = 0;
  unsigned long long busy[5] = {0, 0, 0, 0, 0};  // this CTA's own work time per phase (A0, A1, B, C, D), barriers excluded
  uint32_t overflow = 0;

  while (lo < hi) {
    const uint32_t F = hi - lo;
    unsigned long long tp0 = globaltimer_ns();
    if (n_waves + 1 >= P.wave_cap) { overflow |= kOvWaves; break; }  // uniform
    if (c == 0 && tid == 0) P.wave_lo[n_waves] = lo;
    // ------------------------------------------------------------------ A0: per-state setup
    const uint32_t sc = (F + G - 1) / G;  // states per CTA slice
    {
      const uint32_t s_begin = min(F, c * sc), s_end = min(F, s_begin + sc);
      uint32_t run = 0;
      for (uint32_t i0 = s_begin; i0 < s_end; i0 += kTile) {
        const uint32_t i = i0 + tid;
        uint32_t nitems = 0, side = 0;
        if (i < s_end) {
          uint32_t fs, s1, s2;
          unpack_key(__ldcg(&P.tuples[lo + i]), fs, s1, s2);
          const uint32_t alo = P.a.off[s1], ahi = P.a.off[s1 + 1], blo = P.b.off[s2], bhi = P.b.off[s2 + 1];
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
          nitems = 1u + (mi ? d1 : d2);
          side = mi ? kSideBit : 0u;
          const float f1 = P.a.fin[s1], f2 = P.b.fin[s2];
          const uint32_t ne1 = P.a.neps ? P.a.neps[s1] : 0u, ne2 = P.b.neps ? P.b.neps[s2] : 0u;
          const uint8_t fl = (uint8_t)(((d1 == ne1 && f1 == w_zero()) ? 1 : 0) | ((ne1 == 0) ? 2 : 0) |
                                       ((d2 == ne2 && f2 == w_zero()) ? 4 : 0) | ((ne2 == 0) ? 8 : 0) | (fs << 4) |
                                       (hs_searched ? 64 : 0));
          P.st_flags[i] = fl;
          const float fw = w_times(f1, f2);  // compose_fst_op.rs:420-449
          P.out_finals[lo + i] = w_is_zero(fw) ? w_zero() : fw;
        }
        uint32_t tile_total;
        const uint32_t ex = cta_exclusive_scan(nitems, s_warp, tile_total);
        if (i < s_end) P.item_loc[i] = (run + ex) | side;
        run += tile_total;
      }
      if (tid == 0) P.part_items[c] = run;
    }
    { unsigned long long tb = globaltimer_ns(); busy[0] += tb - tp0; }
    grid_barrier(P.barrier, bar_epoch);
    unsigned long long ts1 = globaltimer_ns();

    // ------------------------------------------------------------------ A1: per-item matching
    cta_prefix_to_smem(P.part_items, G, s_pref_a, s_warp);
    const uint32_t T = s_pref_a[G];
    overflow = __ldcg(&P.ctl[1]);
    if (T > P.items_cap) overflow |= kOvScratch;
    if (overflow) break;  // uniform
    const uint32_t ic = (((T + G - 1) / G) + kTile - 1) / kTile * kTile;  // items per CTA slice (multiple of a tile)
    const uint32_t it_begin = min(T, c * ic), it_end = min(T, it_begin + ic);
    uint32_t my_active = 0, my_arcs = 0;
    if (it_begin < it_end) {
      // state containing my first item: producing CTA slice from the shared prefix, then bisect its local offsets
      if (tid == 0) {
        const uint32_t p = smem_segment(s_pref_a, G, it_begin);
        const uint32_t want = it_begin - s_pref_a[p];
        uint32_t l = min(F, p * sc), h = min(F, l + sc);
        while (h - l > 1) {
          const uint32_t mid = (l + h) >> 1;
          if ((__ldcg(&P.item_loc[mid]) & ~kSideBit) <= want) l = mid; else h = mid;
        }
        s_misc[0] = l;
      }
      __syncthreads();
      uint32_t i_cur = s_misc[0];
      for (uint32_t t0 = it_begin; t0 < it_end; t0 += kTile) {
        // window of item start offsets for states i_cur .. i_cur + 256
        for (uint32_t k = tid; k < kTile + 1; k += kCoopThreads) {
          const uint32_t i = i_cur + k;
          const uint32_t il = i < F ? __ldcg(&P.item_loc[i]) : 0u;
          s_seg[k] = i < F ? s_pref_a[i / sc] + (il & ~kSideBit) : T;
          s_side[k] = il & kSideBit;
        }
        __syncthreads();
        const uint32_t t = t0 + tid;
        uint32_t cnt_out = 0;
        uint4 rec = make_uint4(0, 0, 0, 0);
        if (t < it_end) {
          const uint32_t k = smem_segment(s_seg, kTile + 1, t);
          const uint32_t i = i_cur + k, j = t - s_seg[k];
          const bool match_input = s_side[k] != 0;
          const uint8_t fl = __ldcg(&P.st_flags[i]);
          const uint32_t fs = (fl >> 4) & 3u;
          const bool hs_searched = (fl & 64) != 0;
          FsFlags ff;
          ff.alleps1 = fl & 1; ff.noeps1 = fl & 2; ff.alleps2 = fl & 4; ff.noeps2 = fl & 8;
          const uint4 so = __ldcg(&P.st_off[i]);
          const uint32_t alo = so.x, ahi = so.y, blo = so.z, bhi = so.w;
          Label label;
          uint32_t it_idx = 0xFFFFFFFFu;  // absolute index of the iterated arc; all ones = implicit epsilon loop
          if (j == 0) label = kNoLabel;
          else if (match_input) { it_idx = alo + j - 1; label = __ldg(&P.a.arcs[it_idx].olabel); }
          else { it_idx = blo + j - 1; label = __ldg(&P.b.arcs[it_idx].ilabel); }
          const bool has_loop = (label == kEps);
          const Label key = (label == kNoLabel) ? kEps : label;
          uint32_t pos, end, fs_loop, fs_real;
          if (match_input) {
            pos = has_loop ? blo : lower_bound_label<false>(P.b.arcs, blo, bhi, key);
            end = run_end<false>(P.b.arcs, pos, bhi, key);
            fs_loop = has_loop ? filter_eval(P.kind, fs, ff, label, kNoLabel) : kNoFs;
            fs_real = filter_eval(P.kind, fs, ff, label, key);
          } else {
            pos = has_loop ? alo : lower_bound_label<true>(P.a.arcs, alo, ahi, key);
            end = run_end<true>(P.a.arcs, pos, ahi, key);
            fs_loop = has_loop ? filter_eval(P.kind, fs, ff, kNoLabel, label) : kNoFs;
            fs_real = filter_eval(P.kind, fs, ff, key, label);
          }
          uint32_t cnt = end - pos;
          bool sigma_mode = false;
          const SigmaDev& sg = match_input ? P.sig2 : P.sig1;  // matcher of the searched side
          if (sg.enabled) {  // IteratorSigmaMatcher::new (sigma_matcher.rs:196-246)
            if (label == sg.label && sg.label != kNoLabel) atomicOr(&P.ctl[1], (uint32_t)kErrBadSigmaLabel);
            if (!has_loop && cnt == 0 && hs_searched && label != kEps && label != kNoLabel &&
                dev_sigma_allowed(sg, label)) {
              if (match_input) { pos = lower_bound_label<false>(P.b.arcs, blo, bhi, sg.label); end = run_end<false>(P.b.arcs, pos, bhi, sg.label); }
              else { pos = lower_bound_label<true>(P.a.arcs, alo, ahi, sg.label); end = run_end<true>(P.a.arcs, pos, ahi, sg.label); }
              cnt = end - pos;
              sigma_mode = true;  // the filter sees the relabelled arc: (label, label), i.e. fs_real as computed
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
        uint32_t tile_active, tile_arcs, ex_act, ex_arcs;
        const uint32_t act = cnt_out ? 1u : 0u;
        cta_exclusive_scan2(act, cnt_out, s_warp, ex_act, ex_arcs, tile_active, tile_arcs);
        if (t < it_end) {
          if (rec.w == 0xFFFFFFFFu) P.st_arc_loc[rec.z] = my_arcs + ex_arcs;  // first arc of state rec.z (slice-local)
          if (act) {
            const uint32_t r = it_begin + my_active + ex_act;  // in-place compaction inside my own item slice
            P.recs[r] = rec;
            P.arc_loc[r] = my_arcs + ex_arcs;
          }
        }
        my_active += tile_active; my_arcs += tile_arcs;
        // advance the state window to the state that contains the first item of the next tile
        const uint32_t nxt = t0 + kTile;
        __syncthreads();
        if (tid == 0) s_misc[0] = i_cur + smem_segment(s_seg, kTile + 1, min(nxt, T - 1));
        __syncthreads();
        i_cur = s_misc[0];
      }
    }
    if (tid == 0) P.part_arcs[c] = my_arcs;
    busy[1] += globaltimer_ns() - ts1;
    grid_barrier(P.barrier, bar_epoch);
    unsigned long long tp1 = globaltimer_ns();

    // ------------------------------------------------------------------ B: emit
    cta_prefix_to_smem(P.part_arcs, G, s_pref_b, s_warp);
    const uint32_t E = s_pref_b[G];
    overflow |= __ldcg(&P.ctl[1]);
    if ((unsigned long long)base + E > P.arcs_cap) overflow |= kOvArcs;
    if (((unsigned long long)hi + E) * 2ull > P.table_cap) overflow |= kOvTable;
    const uint32_t arc_chunk = ((E + G - 1) / G + 31u) & ~31u;
    if (arc_chunk > (1u << kLocalRankBits)) overflow |= kOvChunk;
    if (overflow) break;  // uniform
    {
      const uint32_t cta_off = s_pref_b[c];
      Tr* __restrict__ wave_arcs = P.out_arcs + base;
      uint32_t cursor = 0;  // first record (slice-local) that can contain the next tile's first arc
      for (uint32_t e0 = 0; e0 < my_arcs; e0 += kTile) {
        for (uint32_t k = tid; k < kTile + 1; k += kCoopThreads)
          s_seg[k] = (cursor + k) < my_active ? P.arc_loc[it_begin + cursor + k] : my_arcs;
        __syncthreads();
        const uint32_t el = e0 + tid;
        if (el < my_arcs) {
          const uint32_t k = smem_segment(s_seg, kTile + 1, el);
          const uint4 rec = P.recs[it_begin + cursor + k];
          const uint32_t kk = el - s_seg[k];
          const bool loop_ok = (rec.y >> 26) & 1u;
          const bool match_input = rec.y >> 31;
          uint32_t s1 = 0, s2 = 0;
          const bool need_tuple = (rec.w == 0xFFFFFFFFu) || (loop_ok && kk == 0);
          if (need_tuple) { uint32_t fs; unpack_key(__ldcg(&P.tuples[lo + rec.z]), fs, s1, s2); }
          Tr it;
          if (rec.w == 0xFFFFFFFFu) it = match_input ? Tr{kEps, kNoLabel, 0.0f, s1} : Tr{kNoLabel, kEps, 0.0f, s2};
          else it = match_input ? load_tr(&P.a.arcs[rec.w]) : load_tr(&P.b.arcs[rec.w]);
          Tr cand;
          uint32_t fsn;
          if (loop_ok && kk == 0) {
            cand = match_input ? Tr{kNoLabel, kEps, 0.0f, s2} : Tr{kEps, kNoLabel, 0.0f, s1};
            fsn = (rec.y >> 27) & 3u;
          } else {
            const uint32_t idx = rec.x + kk - (loop_ok ? 1u : 0u);
            cand = match_input ? load_tr(&P.b.arcs[idx]) : load_tr(&P.a.arcs[idx]);
            fsn = (rec.y >> 29) & 3u;
            if ((rec.y >> 25) & 1u) {  // sigma match: relabel (value_openfst, sigma_matcher.rs:249-276)
              const SigmaDev& sg = match_input ? P.sig2 : P.sig1;
              const Label l = match_input ? it.olabel : it.ilabel;
              if (sg.rewrite_both) { if (cand.ilabel == sg.label) cand.ilabel = l; if (cand.olabel == sg.label) cand.olabel = l; }
              else if (match_input) cand.ilabel = l;
              else cand.olabel = l;
            }
          }
          const Tr& arc1 = match_input ? it : cand;
          const Tr& arc2 = match_input ? cand : it;
          Tr out;
          out.ilabel = arc1.ilabel;
          out.olabel = arc2.olabel;
          out.weight = w_times(arc1.weight, arc2.weight);
          const unsigned long long key = pack_key(fsn, arc1.nextstate, arc2.nextstate);
          const uint32_t e = cta_off + el;  // canonical wave-local emission index
          uint32_t h = hash_key(key) & P.mask;
          while (true) {
            unsigned long long curk = *reinterpret_cast<volatile unsigned long long*>(&P.slots[h].key);
            if (curk == key) break;
            if (curk == kEmptyKey) {
              unsigned long long prev = atomicCAS(&P.slots[h].key, kEmptyKey, key);
              if (prev == kEmptyKey || prev == key) break;
            }
            h = (h + 1) & P.mask;
          }
          const uint32_t id = *reinterpret_cast<volatile uint32_t*>(&P.slots[h].id);
          if (id != kUnassigned) out.nextstate = id;
          else { atomicMin(&P.slots[h].emin, e); out.nextstate = kPendingBit | h; }
          store_tr(&wave_arcs[e], out);
        }
        __syncthreads();
        if (tid == 0) s_misc[1] = cursor + smem_segment(s_seg, kTile + 1, min(e0 + kTile, my_arcs - 1));
        __syncthreads();
        cursor = s_misc[1];
      }
      // state -> first arc (CSR offsets of the result): slice of the arc owner = slice that processed item 0 of the state
      for (uint32_t i = c * sc + tid; i < min(F, (c + 1) * sc); i += kCoopThreads) {
        const uint32_t t_first = s_pref_a[i / sc] + (__ldcg(&P.item_loc[i]) & ~kSideBit);
        P.out_offsets[lo + i] = base + s_pref_b[t_first / ic] + __ldcg(&P.st_arc_loc[i]);
      }
    }
    busy[2] += globaltimer_ns() - tp1;
    grid_barrier(P.barrier, bar_epoch);
    unsigned long long tp2 = globaltimer_ns();

    // ------------------------------------------------------------------ C: rank first emissions
    const uint32_t e_begin = min(E, c * arc_chunk), e_end = min(E, e_begin + arc_chunk);
    {
      const Tr* __restrict__ wave_arcs = P.out_arcs + base;
      uint32_t cta_new = 0;
      for (uint32_t t0 = e_begin; t0 < e_end; t0 += kCoopThreads) {
        const uint32_t e = t0 + tid;
        bool owner = false;
        uint32_t h = 0;
        if (e < e_end) {
          const uint32_t ns = __ldcg(&wave_arcs[e].nextstate);
          if (ns & kPendingBit) { h = ns & ~kPendingBit; owner = (__ldcg(&P.slots[h].emin) == e); }
        }
        uint32_t tile_total;
        const uint32_t ex = cta_exclusive_scan(owner ? 1u : 0u, s_warp, tile_total);
        if (owner) P.slots[h].id = kTempFlag | (c << kLocalRankBits) | (cta_new + ex);
        cta_new += tile_total;
      }
      if (tid == 0) P.part_new[c] = cta_new;
    }
    busy[3] += globaltimer_ns() - tp2;
    grid_barrier(P.barrier, bar_epoch);
    unsigned long long tp3 = globaltimer_ns();

    // ------------------------------------------------------------------ D: resolve
    cta_prefix_to_smem(P.part_new, G, s_pref_a, s_warp);
    const uint32_t n_new = s_pref_a[G];
    if ((unsigned long long)hi + n_new > P.states_cap || (unsigned long long)hi + n_new >= 0x7FFFFFFFull) {
      overflow |= kOvStates;
      break;  // uniform
    }
    {
      Tr* __restrict__ wave_arcs = P.out_arcs + base;
      for (uint32_t e = e_begin + tid; e < e_end; e += kCoopThreads) {
        const uint32_t ns = __ldcg(&wave_arcs[e].nextstate);
        if (!(ns & kPendingBit)) continue;
        const uint32_t h = ns & ~kPendingBit;
        const uint32_t v = *reinterpret_cast<volatile uint32_t*>(&P.slots[h].id);
        uint32_t id = v;
        if (v & kTempFlag) id = hi + s_pref_a[(v & ~kTempFlag) >> kLocalRankBits] + (v & ((1u << kLocalRankBits) - 1u));
        wave_arcs[e].nextstate = id;
        if (__ldcg(&P.slots[h].emin) == e) {  // first emitter: publish
          P.slots[h].id = id;
          P.tuples[id] = __ldcg(&P.slots[h].key);
        }
      }
    }
    n_states_exp += F; n_items += T; n_arcs += E; n_waves++;
    base += E;
    lo = hi;
    hi += n_new;
    busy[4] += globaltimer_ns() - tp3;
    grid_barrier(P.barrier, bar_epoch);
    unsigned long long tp4 = globaltimer_ns();
    t_a += tp1 - tp0; t_b += tp2 - tp1; t_c += tp3 - tp2; t_d += tp4 - tp3;
  }

  if (tid == 0) {
    for (int k = 0; k < 5; k++) { atomicAdd(&P.stats[8 + k], busy[k]); atomicMax(&P.stats[13 + k], busy[k]); }
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
void launch_unpack_s1(const unsigned long long* tuples, uint32_t n, uint32_t* s1_out, uint32_t n_starts,
                      uint32_t* start_map, cudaStream_t s) {
  uint32_t m = n > n_starts ? n : n_starts;
  if (m) k_unpack_s1<<<blocks_for(m), kThreads, 0, s>>>(tuples, n, s1_out, n_starts, start_map);
}

static int grid_used = 1;
float run_coop(const CoopParams& P0, int sms, cudaStream_t s) {
  CoopParams P = P0;
  // resident CTAs per SM the kernel is compiled for (register budget): 4 -> 64 regs, 5 -> 48, 6 -> 40
  int minb = 6;  // measured best on C3 (40 registers, no spills, 6 x 256 threads per SM)
  if (const char* e = std::getenv("B200_COOP_MINBLOCKS")) minb = std::atoi(e);
  void* kern = (void*)k_compose_coop<6>;
  if (minb == 5) kern = (void*)k_compose_coop<5>;
  else if (minb == 4) kern = (void*)k_compose_coop<4>;
  else if (minb == 3) kern = (void*)k_compose_coop<3>;
  else if (minb == 8) kern = (void*)k_compose_coop<8>;
  int per_sm = 0;
  size_t dyn = 2 * 2049 * sizeof(uint32_t);
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kCoopThreads, dyn));
  if (per_sm < 1) throw FstError("cooperative compose kernel does not fit on the device");
  int grid = sms * per_sm;
  if (grid > 2047) grid = 2047;  // CTA index must fit 11 bits next to the 20-bit local rank; prefix arrays hold 2048
  grid_used = grid;
  dyn = 2 * ((size_t)grid + 1) * sizeof(uint32_t);
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
  DevBuf<unsigned long long> tuples(s, states_cap), dstats(s, 18);
  DevBuf<Slot> slots(s, table_cap);
  DevBuf<uint4> recs(s, items_cap);
  DevBuf<uint32_t> arc_loc(s, items_cap), item_loc(s, states_cap), st_arc_loc(s, states_cap), parts(s, 3 * 2048), ctl(s, 8);
  DevBuf<uint8_t> st_flags(s, states_cap);
  DevBuf<uint4> st_off(s, states_cap);
  const uint32_t wave_cap = 1u << 20;
  DevBuf<uint32_t> wave_lo(s, wave_cap);
  B200_CUDA(cudaMemsetAsync(slots.p, 0xFF, table_cap * sizeof(Slot), s));
  B200_CUDA(cudaMemsetAsync(dstats.p, 0, 18 * sizeof(unsigned long long), s));
  P.tuples = tuples.p; P.states_cap = (uint32_t)states_cap;
  P.out_offsets = out.offsets.p; P.out_finals = out.finals.p; P.out_arcs = out.arcs.p; P.arcs_cap = (uint32_t)arcs_cap;
  P.slots = slots.p; P.mask = (uint32_t)table_cap - 1; P.table_cap = (uint32_t)table_cap;
  P.item_loc = item_loc.p; P.st_arc_loc = st_arc_loc.p; P.st_flags = st_flags.p; P.st_off = st_off.p;
  P.recs = recs.p; P.arc_loc = arc_loc.p; P.items_cap = (uint32_t)std::min<size_t>(items_cap, 0xFFFFFFF0ull);
  P.part_arcs = parts.p; P.part_items = parts.p + 2048; P.part_new = parts.p + 4096;
  P.ctl = ctl.p; P.stats = dstats.p;
  P.barrier = ctl.p + 6;
  P.wave_lo = wave_lo.p; P.wave_cap = wave_cap;
  uint32_t start_fs = (kind == kNullFilter || kind == kTrivialFilter || kind == kNoMatchFilter) ? 1u : 0u;
  P.n_starts = n_starts;
  k_coop_init<<<blocks_for(n_starts), kThreads, 0, s>>>(slots.p, P.mask, tuples.p, start_fs,
                                                        batch ? batch->d_starts1 : nullptr, fa.start, fb.start, n_starts, ctl.p);
  st.kernel_launches++;

  float ms_kernel = run_coop(P, sm_count(), s);
  st.kernel_launches++; st.emit_launches = 1;

  uint32_t hctl[4];
  unsigned long long hstats[18];
  B200_CUDA(cudaMemcpyAsync(hctl, ctl.p, 16, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaMemcpyAsync(hstats, dstats.p, 18 * 8, cudaMemcpyDeviceToHost, s));
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
    const char* names[5] = {"A0", "A1", "B", "C", "D"};
    for (int k = 0; k < 5; k++)
      std::fprintf(stderr, "[coop] phase %s: CTA busy avg %.3f ms, max %.3f ms\n", names[k],
                   hstats[8 + k] * 1e-6 / grid_used, hstats[13 + k] * 1e-6);
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

DevFst compose_device(const DevFst& a, const DevFst& b, const ComposeOptions& opt, ComposeStats* stats,
                      cudaStream_t s) {
  const char* impl = std::getenv("B200_COMPOSE_IMPL");
  bool want_waves = impl && std::string(impl) == "waves";
  if (!want_waves) {
    DevFst out(s);
    if (compose_device_coop(a, b, opt, stats, s, &out, nullptr)) return out;
  }
  return compose_device_waves(a, b, opt, stats, s);
}

}  // namespace b200
