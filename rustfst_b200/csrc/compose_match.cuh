// compose_match.cuh — device pieces shared by the persistent composition kernels (compose_coop.cu, compose_ws.cu):
// overflow / error flags, the sigma-matcher view, and the sorted-matcher search on dense label arrays.
#pragma once
#include "compose_common.cuh"

namespace b200 {
namespace composeimpl {

enum Overflow : uint32_t { kOvArcs = 1, kOvStates = 2, kOvTable = 4, kOvScratch = 8, kOvChunk = 16, kOvWaves = 32, kOvRuns = 64,
                           kErrBothRequire = 0x100, kErrBadSigmaLabel = 0x200, kErrWatchdog = 0x400 };

// Device view of one SigmaMatcher (sigma_matcher.rs): arcs labelled sigma_label on the matched side match any
// (allowed) label that has no ordinary match at the state; the matched arc is relabelled.
struct SigmaDev {
  uint32_t enabled; uint32_t label; uint32_t rewrite_both; const uint32_t* allowed; uint32_t n_allowed;
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
// Lower bound + equal run of `key` in the label-sorted slice [lo, hi) with as few DEPENDENT loads as possible: the
// kernel is latency-bound, so a 4-ary narrowing (3 independent probes per round) is followed by one round that loads
// a 16-arc window at once and counts "< key" and "== key" (sorted_matcher.rs:141-142,166-184: lower_bound_by, then
// iterate while the label matches).  from_lo = the epsilon-loop search, which starts at lo instead of bisecting.
// The matcher compares fst1 output labels with fst2 input labels only, so both are kept once more as dense 4-byte
// arrays (lab1[i] = fst1 arc i .olabel, lab2[i] = fst2 arc i .ilabel; padded by kLabelPad entries): the label of an
// iterated arc is a coalesced 4-byte load, and the 16-label window of a search is five aligned 128-bit loads instead
// of sixteen strided ones.  Lanes matching on different sides run the same instruction stream (pointer select).
constexpr uint32_t kLabelPad = 32;
__device__ __forceinline__ uint32_t lower_bound_lab(const uint32_t* __restrict__ lab, uint32_t lo, uint32_t hi, Label key) {
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(&lab[mid]) < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ uint32_t run_end_lab(const uint32_t* __restrict__ lab, uint32_t pos, uint32_t hi, Label key) {
  uint32_t p = pos;
  const uint32_t lim = pos + 8 < hi ? pos + 8 : hi;
  while (p < lim) { if (__ldg(&lab[p]) != key) return p; p++; }
  if (p == hi) return p;
  return lower_bound_lab(lab, p, hi, key + 1);  // key + 1 cannot overflow: kNoLabel is never searched for
}
// Lower bound + equal run of `key` in the sorted label slice [lo, hi) with as few DEPENDENT loads as possible
// (sorted_matcher.rs:141-142,166-184: lower_bound_by, then iterate while the label matches): 4-ary narrowing (three
// independent probes per round) down to 16 labels, then one round that fetches the window and counts "< key" and
// "== key".  from_lo = the epsilon-loop search, which starts at lo instead of bisecting.
__device__ __forceinline__ void match_range(const uint32_t* __restrict__ lab, uint32_t lo, uint32_t hi, Label key,
                                            bool from_lo, uint32_t& pos, uint32_t& end) {
  uint32_t l = lo, h = hi;
  if (!from_lo) {
    while (h - l > 16) {  // invariant: labels below l are < key, labels from h on are >= key
      const uint32_t q = (h - l) >> 2, m1 = l + q, m2 = m1 + q, m3 = m2 + q;
      const Label x1 = __ldg(&lab[m1]), x2 = __ldg(&lab[m2]), x3 = __ldg(&lab[m3]);
      if (x1 >= key) h = m1;
      else if (x2 >= key) { l = m1 + 1; h = m2; }
      else if (x3 >= key) { l = m2 + 1; h = m3; }
      else l = m3 + 1;
    }
  }
  // window [l, l + 16) lies inside the five aligned quads starting at l & ~3 (reads past hi hit the padding)
  const uint32_t l4 = l & ~3u;
  const uint4* __restrict__ q4 = reinterpret_cast<const uint4*>(lab + l4);
  const uint32_t w_end = min(hi, l + 16);
  uint4 v[5];
#pragma unroll
  for (int k = 0; k < 5; k++) v[k] = __ldg(q4 + k);
  // "< key" / "== key" collected as 20-bit masks (one compare + one predicated OR per label), cut to the valid window
  // [l, w_end) once
  uint32_t lt_m = 0, eq_m = 0;
#pragma unroll
  for (int k = 0; k < 5; k++) {
    const uint32_t xs[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (xs[u] < key) lt_m |= 1u << (4 * k + u);
      if (xs[u] == key) eq_m |= 1u << (4 * k + u);
    }
  }
  const uint32_t vm = ((1u << (w_end - l4)) - 1u) & ~((1u << (l - l4)) - 1u);
  const uint32_t n_lt = __popc(lt_m & vm), n_eq = __popc(eq_m & vm);
  pos = l + n_lt;
  end = pos + n_eq;
  if (end == l + 16 && end < hi) end = run_end_lab(lab, end, hi, key);  // run leaves the window (rare)
}


// match_range for callers that only need the position of a NON-EMPTY run (pos is unspecified when end == pos): on a
// sorted slice the first label equal to the key IS the lower bound, so the 20 "< key" compares of match_range are
// replaced by one find-first-set on the "== key" mask (the window compare loop was 9 % of all executed instructions).
__device__ __forceinline__ void match_range_eq(const uint32_t* __restrict__ lab, uint32_t lo, uint32_t hi, Label key,
                                               bool from_lo, uint32_t& pos, uint32_t& end) {
  uint32_t l = lo, h = hi;
  if (!from_lo) {
    while (h - l > 16) {  // invariant: labels below l are < key, labels from h on are >= key
      const uint32_t q = (h - l) >> 2, m1 = l + q, m2 = m1 + q, m3 = m2 + q;
      const Label x1 = __ldg(&lab[m1]), x2 = __ldg(&lab[m2]), x3 = __ldg(&lab[m3]);
      if (x1 >= key) h = m1;
      else if (x2 >= key) { l = m1 + 1; h = m2; }
      else if (x3 >= key) { l = m2 + 1; h = m3; }
      else l = m3 + 1;
    }
  }
  const uint32_t l4 = l & ~3u;
  const uint4* __restrict__ q4 = reinterpret_cast<const uint4*>(lab + l4);
  const uint32_t w_end = min(hi, l + 16);
  uint4 v[5];
#pragma unroll
  for (int k = 0; k < 5; k++) v[k] = __ldg(q4 + k);
  uint32_t eq_m = 0;
#pragma unroll
  for (int k = 0; k < 5; k++) {
    const uint32_t xs[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (xs[u] == key) eq_m |= 1u << (4 * k + u);
  }
  eq_m &= ((1u << (w_end - l4)) - 1u) & ~((1u << (l - l4)) - 1u);
  if (eq_m) {
    pos = l4 + (uint32_t)__ffs((int)eq_m) - 1u;
    end = pos + __popc(eq_m);
  } else {
    pos = end = l;
    // more labels beyond the window: the run may begin right behind it if everything in the window is smaller
    if (w_end < hi && __ldg(&lab[w_end - 1]) < key) pos = end = w_end;
  }
  if (end == l + 16 && end < hi) end = run_end_lab(lab, end, hi, key);  // run leaves the window (rare)
}

// The same search in three steps, so that a lane that handles several items can issue the window loads of all of them
// before it evaluates any (the kernels are latency-bound: loads in flight together cost one round trip).
__device__ __forceinline__ uint32_t match_narrow(const uint32_t* __restrict__ lab, uint32_t lo, uint32_t hi, Label key,
                                                 bool from_lo) {
  uint32_t l = lo, h = hi;
  if (!from_lo) {
    while (h - l > 16) {  // invariant: labels below l are < key, labels from h on are >= key
      const uint32_t q = (h - l) >> 2, m1 = l + q, m2 = m1 + q, m3 = m2 + q;
      const Label x1 = __ldg(&lab[m1]), x2 = __ldg(&lab[m2]), x3 = __ldg(&lab[m3]);
      if (x1 >= key) h = m1;
      else if (x2 >= key) { l = m1 + 1; h = m2; }
      else if (x3 >= key) { l = m2 + 1; h = m3; }
      else l = m3 + 1;
    }
  }
  return l;
}
struct LabelWindow { uint4 v[5]; };
__device__ __forceinline__ void match_window_load(const uint32_t* __restrict__ lab, uint32_t l, LabelWindow& w) {
  const uint4* __restrict__ q4 = reinterpret_cast<const uint4*>(lab + (l & ~3u));
#pragma unroll
  for (int k = 0; k < 5; k++) w.v[k] = __ldg(q4 + k);
}
__device__ __forceinline__ void match_window_eval(const uint32_t* __restrict__ lab, const LabelWindow& w, uint32_t l,
                                                  uint32_t hi, Label key, uint32_t& pos, uint32_t& end) {
  const uint32_t l4 = l & ~3u, w_end = min(hi, l + 16);
  uint32_t lt_m = 0, eq_m = 0;
#pragma unroll
  for (int k = 0; k < 5; k++) {
    const uint32_t xs[4] = {w.v[k].x, w.v[k].y, w.v[k].z, w.v[k].w};
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (xs[u] < key) lt_m |= 1u << (4 * k + u);
      if (xs[u] == key) eq_m |= 1u << (4 * k + u);
    }
  }
  const uint32_t vm = ((1u << (w_end - l4)) - 1u) & ~((1u << (l - l4)) - 1u);
  pos = l + __popc(lt_m & vm);
  end = pos + __popc(eq_m & vm);
  if (end == l + 16 && end < hi) end = run_end_lab(lab, end, hi, key);  // run leaves the window (rare)
}

// s1_out[i] = fst1 component of tuples[i]; start_map[i] = i (untrimmed batch results).  Defined in compose_coop.cu.
void launch_unpack_s1(const unsigned long long* tuples, uint32_t n, uint32_t* s1_out, uint32_t n_starts,
                      uint32_t* start_map, cudaStream_t s);
// lab[i] = arcs[i].olabel (olabel != 0) or .ilabel, padded with kNoLabel up to n_padded entries.  Defined in compose_coop.cu.
void launch_extract_labels(const Tr* arcs, uint32_t n, uint32_t n_padded, int olabel, uint32_t* out, cudaStream_t s);

}  // namespace composeimpl
}  // namespace b200
