// compose_common.cuh — device helpers shared by the two composition back ends (compose.cu: one kernel per phase with
// host-sized launches; compose_coop.cu: one persistent cooperative kernel per compose): packed state tuples, the
// open-addressed state table slot, 128-bit arc loads/stores, the six epsilon filters and the sorted-matcher searches.
#pragma once
#include "algos.h"

namespace b200 {
namespace composeimpl {

constexpr unsigned long long kEmptyKey = ~0ull;
constexpr uint32_t kUnassigned = 0xFFFFFFFFu;
constexpr uint32_t kPendingBit = 0x80000000u;
constexpr uint32_t kNoFs = 0xFFu;

enum MatchSide : int { kMatchInput = 0, kMatchOutput = 1, kMatchBoth = 2 };

struct __align__(16) Slot {
  unsigned long long key;  // packed (s1, s2, filter state); all ones = empty
  uint32_t id;             // product state id, kUnassigned until the wave that discovers it resolves
  uint32_t emin;           // smallest wave-local emission index that reached the tuple in its discovery wave
};

struct FstView {
  const uint32_t* off;
  const Tr* arcs;
  const float* fin;
  const uint32_t* neps;  // #arcs with epsilon on the matched side per state (fst1: olabel, fst2: ilabel); may be null
  uint32_t n;
};

__host__ __device__ __forceinline__ unsigned long long pack_key(uint32_t fs, uint32_t s1, uint32_t s2) {
  return (unsigned long long)s1 | ((unsigned long long)s2 << 31) | ((unsigned long long)fs << 62);
}
__device__ __forceinline__ void unpack_key(unsigned long long k, uint32_t& fs, uint32_t& s1, uint32_t& s2) {
  s1 = (uint32_t)(k & 0x7FFFFFFFull);
  s2 = (uint32_t)((k >> 31) & 0x7FFFFFFFull);
  fs = (uint32_t)(k >> 62);
}
__host__ __device__ __forceinline__ uint32_t hash_key(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return (uint32_t)k;
}

__device__ __forceinline__ Tr load_tr(const Tr* p) {  // one 128-bit read-only load per arc
  int4 v = __ldg(reinterpret_cast<const int4*>(p));
  Tr t;
  t.ilabel = (uint32_t)v.x; t.olabel = (uint32_t)v.y; t.weight = __int_as_float(v.z); t.nextstate = (uint32_t)v.w;
  return t;
}
__device__ __forceinline__ void store_tr(Tr* p, const Tr& t) {
  *reinterpret_cast<int4*>(p) = make_int4((int)t.ilabel, (int)t.olabel, __float_as_int(t.weight), (int)t.nextstate);
}

// Filter state flags of a product state (sequence_compose_filter.rs:134-148, alt_sequence:139-153, match:132-161)
struct FsFlags { bool alleps1, noeps1, alleps2, noeps2; };

__device__ __forceinline__ FsFlags state_flags(const FstView& a, const FstView& b, uint32_t s1, uint32_t s2) {
  uint32_t na1 = a.off[s1 + 1] - a.off[s1], na2 = b.off[s2 + 1] - b.off[s2];
  uint32_t ne1 = a.neps ? a.neps[s1] : 0u, ne2 = b.neps ? b.neps[s2] : 0u;
  bool fin1 = a.fin[s1] != w_zero(), fin2 = b.fin[s2] != w_zero();
  FsFlags f;
  f.alleps1 = (na1 == ne1) && !fin1; f.noeps1 = (ne1 == 0);
  f.alleps2 = (na2 == ne2) && !fin2; f.noeps2 = (ne2 == 0);
  return f;
}

// filter_tr of the six filters; only arc1.olabel and arc2.ilabel are ever inspected.
// Returns the next filter state or kNoFs.  Bool-valued filters use 1 = true, "false" = no state.
__device__ __forceinline__ uint32_t filter_eval(int kind, uint32_t fs, const FsFlags& f, Label ol1, Label il2) {
  switch (kind) {
    case kSequenceFilter:  // sequence_compose_filter.rs:150-171
      if (ol1 == kNoLabel) return f.alleps1 ? kNoFs : (f.noeps1 ? 0u : 1u);
      if (il2 == kNoLabel) return fs != 0 ? kNoFs : 0u;
      if (ol1 == kEps) return kNoFs;
      return 0u;
    case kAltSequenceFilter:  // alt_sequence_compose_filter.rs:156-177
      if (il2 == kNoLabel) return f.alleps2 ? kNoFs : (f.noeps2 ? 0u : 1u);
      if (ol1 == kNoLabel) return fs == 1 ? kNoFs : 0u;
      if (ol1 == kEps) return kNoFs;
      return 0u;
    case kMatchFilter:  // match_compose_filter.rs:163-206
      if (il2 == kNoLabel) {
        if (fs == 0) return f.noeps2 ? 0u : (f.alleps2 ? kNoFs : 1u);
        return fs == 1 ? 1u : kNoFs;
      }
      if (ol1 == kNoLabel) {
        if (fs == 0) return f.noeps1 ? 0u : (f.alleps1 ? kNoFs : 2u);
        return fs == 2 ? 2u : kNoFs;
      }
      if (ol1 == kEps) return fs == 0 ? 0u : kNoFs;
      return 0u;
    case kNullFilter:  // null_compose_filter.rs:124-131
      return (ol1 == kNoLabel || il2 == kNoLabel) ? kNoFs : 1u;
    case kTrivialFilter:  // trivial_compose_filter.rs:122-124
      return 1u;
    default:  // kNoMatchFilter, no_match_compose_filter.rs:124-128
      return (ol1 != kEps || il2 != kEps) ? 1u : kNoFs;
  }
}

// superslice lower_bound_by on one label field of a state's arc slice (sorted_matcher.rs:141-142)
template <bool kByOlabel>
__device__ __forceinline__ uint32_t lower_bound_label(const Tr* arcs, uint32_t lo, uint32_t hi, Label key) {
  while (lo < hi) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    Label l = kByOlabel ? __ldg(&arcs[mid].olabel) : __ldg(&arcs[mid].ilabel);
    if (l < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
template <bool kByOlabel>
__device__ __forceinline__ uint32_t run_end(const Tr* arcs, uint32_t pos, uint32_t hi, Label key) {
  // the iterator yields arcs while label == key (sorted_matcher.rs:166-184); for long runs fall back to bisection
  uint32_t p = pos;
  uint32_t lim = pos + 8 < hi ? pos + 8 : hi;
  while (p < lim) {
    Label l = kByOlabel ? __ldg(&arcs[p].olabel) : __ldg(&arcs[p].ilabel);
    if (l != key) return p;
    p++;
  }
  if (p == hi) return p;
  // key + 1 cannot overflow here: kNoLabel is never searched for (it is mapped to epsilon)
  return lower_bound_label<kByOlabel>(arcs, p, hi, key + 1);
}

// Largest index t in [0, n) with a[t] <= x   (a is an exclusive prefix sum, a[0] = 0 <= x)
__device__ __forceinline__ uint32_t find_segment(const uint32_t* a, uint32_t n, uint32_t x) {
  uint32_t lo = 0, hi = n;
  while (hi - lo > 1) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(&a[mid]) <= x) lo = mid; else hi = mid;
  }
  return lo;
}


// Match side from the stored property bits (compose_fst_op.rs:169-197, sorted_matcher.rs:56-85); throws the
// reference's errors when neither side can be matched.
inline int resolve_match_side(uint64_t props_a, uint64_t props_b) {
  auto mtype = [](uint64_t p, uint64_t yes, uint64_t no) { return (p & yes) ? 1 : ((p & no) ? 0 : -1); };
  int t1 = mtype(props_a, props::kOLabelSorted, props::kNotOLabelSorted);
  int t2 = mtype(props_b, props::kILabelSorted, props::kNotILabelSorted);
  if (t1 == 1 && t2 == 1) return kMatchBoth;
  if (t1 == 1) return kMatchOutput;
  if (t2 == 1) return kMatchInput;
  if (t1 == -1)  // matcher1.match_type(true) -> properties_check fails first (fst_traits/fst.rs:166-176)
    throw FstError("Properties are not known : O_LABEL_SORTED | NOT_O_LABEL_SORTED. Properties of the Fst : " +
                   std::to_string(props_a));
  if (t2 == -1)
    throw FstError("Properties are not known : I_LABEL_SORTED | NOT_I_LABEL_SORTED. Properties of the Fst : " +
                   std::to_string(props_b));
  throw FstError(
      "ComposeFst: 1st argument cannot match on output labels and 2nd argument cannot match on input labels "
      "(sort?).");
}

// neps[s] = number of arcs of state s whose olabel (by_olabel) / ilabel is epsilon.  Defined in compose.cu.
void launch_count_eps(const uint32_t* off, const Tr* arcs, uint32_t n, int by_olabel, uint32_t* neps, cudaStream_t s);

}  // namespace composeimpl
}  // namespace b200
