// fst_types.h — arc record, label/state constants, tropical weight arithmetic and the 64-bit FstProperties word,
// shared by the host container, the C-ABI and the CUDA kernels of librustfst_b200.
//
// Reference interfaces mirrored (paths relative to /root/reference):
//   Tr {ilabel, olabel, weight, nextstate}          rustfst/src/tr.rs:6-15, rustfst-ffi/src/tr.rs:9-21
//   EPS_LABEL / NO_LABEL / NO_STATE_ID / KDELTA     rustfst/src/lib.rs:236,269,292,298
//   TropicalWeight plus/times/approx-eq             rustfst/src/semirings/tropical_weight.rs:53-70,
//                                                   rustfst/src/semirings/semiring.rs:159-168
//   FstProperties bit layout and mutation masks     rustfst/src/fst_properties/properties.rs:21-103,107-330
#pragma once
#include <cstdint>
#include <limits>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif

namespace b200 {

using Label = uint32_t;
using StateId = uint32_t;

constexpr Label kEps = 0;
constexpr Label kNoLabel = 0xFFFFFFFFu;
constexpr StateId kNoState = 0xFFFFFFFFu;
constexpr float kDelta = 1.0f / 1024.0f;

struct alignas(16) Tr {
  Label ilabel;
  Label olabel;
  float weight;
  StateId nextstate;
};
static_assert(sizeof(Tr) == 16, "Tr must be the 16-byte wire/ABI record");

B200_HD float w_zero() {
#if defined(__CUDA_ARCH__)
  return __int_as_float(0x7f800000);
#else
  return std::numeric_limits<float>::infinity();
#endif
}
B200_HD bool w_approx_eq(float a, float b) { return a <= (b + kDelta) && b <= (a + kDelta); }
B200_HD bool w_is_zero(float w) { return w_approx_eq(w, w_zero()); }
B200_HD bool w_is_one(float w) { return w_approx_eq(w, 0.0f); }
B200_HD float w_plus(float a, float b) { return (b < a) ? b : a; }
B200_HD float w_times(float a, float b) {
  if (a == w_zero()) return a;
  if (b == w_zero()) return b;
  return a + b;
}

namespace props {
constexpr uint64_t kExpanded = 1, kMutable = 2;
constexpr uint64_t kAcceptor = 0x0000000000010000ULL, kNotAcceptor = 0x0000000000020000ULL;
constexpr uint64_t kIDeterministic = 0x0000000000040000ULL, kNotIDeterministic = 0x0000000000080000ULL;
constexpr uint64_t kODeterministic = 0x0000000000100000ULL, kNotODeterministic = 0x0000000000200000ULL;
constexpr uint64_t kEpsilons = 0x0000000000400000ULL, kNoEpsilons = 0x0000000000800000ULL;
constexpr uint64_t kIEpsilons = 0x0000000001000000ULL, kNoIEpsilons = 0x0000000002000000ULL;
constexpr uint64_t kOEpsilons = 0x0000000004000000ULL, kNoOEpsilons = 0x0000000008000000ULL;
constexpr uint64_t kILabelSorted = 0x0000000010000000ULL, kNotILabelSorted = 0x0000000020000000ULL;
constexpr uint64_t kOLabelSorted = 0x0000000040000000ULL, kNotOLabelSorted = 0x0000000080000000ULL;
constexpr uint64_t kWeighted = 0x0000000100000000ULL, kUnweighted = 0x0000000200000000ULL;
constexpr uint64_t kCyclic = 0x0000000400000000ULL, kAcyclic = 0x0000000800000000ULL;
constexpr uint64_t kInitialCyclic = 0x0000001000000000ULL, kInitialAcyclic = 0x0000002000000000ULL;
constexpr uint64_t kTopSorted = 0x0000004000000000ULL, kNotTopSorted = 0x0000008000000000ULL;
constexpr uint64_t kAccessible = 0x0000010000000000ULL, kNotAccessible = 0x0000020000000000ULL;
constexpr uint64_t kCoAccessible = 0x0000040000000000ULL, kNotCoAccessible = 0x0000080000000000ULL;
constexpr uint64_t kString = 0x0000100000000000ULL, kNotString = 0x0000200000000000ULL;
constexpr uint64_t kWeightedCycles = 0x0000400000000000ULL, kUnweightedCycles = 0x0000800000000000ULL;

constexpr uint64_t kBinary = 0x7ULL;
constexpr uint64_t kTrinary = 0x0000ffffffff0000ULL;
constexpr uint64_t kPosTrinary = kTrinary & 0x5555555555555555ULL;
constexpr uint64_t kNegTrinary = kTrinary & 0xaaaaaaaaaaaaaaaaULL;
constexpr uint64_t kAll = kBinary | kTrinary;

// Pairs written as (positive | negative) keep the masks below readable.
constexpr uint64_t kAcceptorPair = kAcceptor | kNotAcceptor;
constexpr uint64_t kDetPairs = kIDeterministic | kNotIDeterministic | kODeterministic | kNotODeterministic;
constexpr uint64_t kEpsPairs = kEpsilons | kNoEpsilons | kIEpsilons | kNoIEpsilons | kOEpsilons | kNoOEpsilons;
constexpr uint64_t kSortPairs = kILabelSorted | kNotILabelSorted | kOLabelSorted | kNotOLabelSorted;
constexpr uint64_t kWeightPair = kWeighted | kUnweighted;
constexpr uint64_t kCyclePair = kCyclic | kAcyclic;
constexpr uint64_t kInitCyclePair = kInitialCyclic | kInitialAcyclic;
constexpr uint64_t kTopPair = kTopSorted | kNotTopSorted;
constexpr uint64_t kAccessPair = kAccessible | kNotAccessible;
constexpr uint64_t kCoAccessPair = kCoAccessible | kNotCoAccessible;
constexpr uint64_t kStringPair = kString | kNotString;
constexpr uint64_t kWCyclesPair = kWeightedCycles | kUnweightedCycles;

// properties.rs null_properties(): what VectorFst::new() starts with.
constexpr uint64_t kNull = kAcceptor | kIDeterministic | kODeterministic | kNoEpsilons | kNoIEpsilons | kNoOEpsilons |
                           kILabelSorted | kOLabelSorted | kUnweighted | kAcyclic | kInitialAcyclic | kTopSorted |
                           kAccessible | kCoAccessible | kString | kUnweightedCycles;
// Masks of what survives each mutation (properties.rs set_start/set_final/add_state/add_arc/delete_states/arcsort).
constexpr uint64_t kKeepOnSetStart = kAcceptorPair | kDetPairs | kEpsPairs | kSortPairs | kWeightPair | kCyclePair |
                                     kTopPair | kCoAccessPair | kWCyclesPair;
constexpr uint64_t kKeepOnSetFinal = kAcceptorPair | kDetPairs | kEpsPairs | kSortPairs | kCyclePair | kInitCyclePair |
                                     kTopPair | kAccessPair | kWCyclesPair;
constexpr uint64_t kKeepOnAddState = kAcceptorPair | kDetPairs | kEpsPairs | kSortPairs | kWeightPair | kCyclePair |
                                     kInitCyclePair | kTopPair | kNotAccessible | kNotCoAccessible | kNotString |
                                     kWCyclesPair;
constexpr uint64_t kKeepOnAddArc = kNotAcceptor | kNotIDeterministic | kNotODeterministic | kEpsilons | kIEpsilons |
                                   kOEpsilons | kNotILabelSorted | kNotOLabelSorted | kWeighted | kCyclic |
                                   kInitialCyclic | kNotTopSorted | kAccessible | kCoAccessible | kWeightedCycles;
constexpr uint64_t kKeepOnDeleteStates = kAcceptor | kIDeterministic | kODeterministic | kNoEpsilons | kNoIEpsilons |
                                         kNoOEpsilons | kILabelSorted | kOLabelSorted | kUnweighted | kAcyclic |
                                         kInitialAcyclic | kTopSorted | kUnweightedCycles;
constexpr uint64_t kKeepOnArcSort = kAcceptorPair | kDetPairs | kEpsPairs | kWeightPair | kCyclePair | kInitCyclePair |
                                    kTopPair | kAccessPair | kCoAccessPair | kStringPair | kWCyclesPair;

// fst_properties/utils.rs:4-9 — a property is known iff one bit of its pair is set.
inline uint64_t known(uint64_t p) {
  return kBinary | (p & kTrinary) | ((p & kPosTrinary) << 1) | ((p & kNegTrinary) >> 1);
}

// fst_properties/mutate_properties.rs:7-13
inline uint64_t on_set_start(uint64_t p) {
  uint64_t out = p & kKeepOnSetStart;
  if (p & kAcyclic) out |= kInitialAcyclic;
  return out;
}
// fst_properties/mutate_properties.rs:15-37 (old == nullptr: state was not final)
inline uint64_t on_set_final(uint64_t p, const float* old_w, const float* new_w) {
  uint64_t out = p;
  if (old_w && !w_is_zero(*old_w) && !w_is_one(*old_w)) out &= ~kWeighted;
  if (new_w && !w_is_zero(*new_w) && !w_is_one(*new_w)) { out |= kWeighted; out &= ~kUnweighted; }
  return out & (kKeepOnSetFinal | kWeightPair);
}
inline uint64_t on_add_state(uint64_t p) { return p & kKeepOnAddState; }
// fst_properties/mutate_properties.rs:43-100
inline uint64_t on_add_tr(uint64_t p, StateId state, const Tr& tr, const Tr* prev) {
  uint64_t out = p;
  if (tr.ilabel != tr.olabel) { out |= kNotAcceptor; out &= ~kAcceptor; }
  if (tr.ilabel == kEps) {
    out |= kIEpsilons; out &= ~kNoIEpsilons;
    if (tr.olabel == kEps) { out |= kEpsilons; out &= ~kNoEpsilons; }
  }
  if (tr.olabel == kEps) { out |= kOEpsilons; out &= ~kNoOEpsilons; }
  if (prev) {
    if (prev->ilabel > tr.ilabel) { out |= kNotILabelSorted; out &= ~kILabelSorted; }
    if (prev->olabel > tr.olabel) { out |= kNotOLabelSorted; out &= ~kOLabelSorted; }
  }
  if (!w_is_zero(tr.weight) && !w_is_one(tr.weight)) { out |= kWeighted; out &= ~kUnweighted; }
  if (tr.nextstate <= state) { out |= kNotTopSorted; out &= ~kTopSorted; }
  out &= kKeepOnAddArc | kAcceptor | kNoEpsilons | kNoIEpsilons | kNoOEpsilons | kILabelSorted | kOLabelSorted |
         kUnweighted | kTopSorted;
  if (out & kTopSorted) out |= kAcyclic | kInitialAcyclic;
  return out;
}
// fst_properties/mutate_properties.rs:151-184
inline uint64_t of_compose(uint64_t p1, uint64_t p2) {
  uint64_t out = kAccessible;
  if ((p1 & kAcceptor) && (p2 & kAcceptor)) {
    out |= kAcceptor;
    out |= (kNoEpsilons | kNoIEpsilons | kNoOEpsilons | kAcyclic | kInitialAcyclic) & p1 & p2;
    if ((p1 & kNoIEpsilons) && (p2 & kNoIEpsilons)) out |= (kIDeterministic | kODeterministic) & p1 & p2;
  } else {
    out |= (kAcceptor | kNoIEpsilons | kAcyclic | kInitialAcyclic) & p1 & p2;
    if ((p1 & kNoIEpsilons) && (p2 & kNoIEpsilons)) out |= kIDeterministic & p1 & p2;
  }
  return out;
}
// algorithms/connect.rs:61-64 after VectorFst::del_states (mutable_fst.rs:186)
inline uint64_t after_connect(uint64_t p) { return (p & kKeepOnDeleteStates) | kAccessible | kCoAccessible; }
// fst_properties/mutate_properties.rs:662-672
inline uint64_t of_shortest_path(uint64_t p, bool tree) {
  uint64_t out = p | kAcyclic | kInitialAcyclic | kAccessible | kUnweightedCycles;
  if (!tree) out |= kCoAccessible;
  return out;
}
// fst_properties/mutate_properties.rs:622-638
inline uint64_t of_reverse(uint64_t p, bool has_superinitial) {
  uint64_t out = (kAcceptorPair | kEpsilons | kIEpsilons | kOEpsilons | kUnweighted | kCyclePair | kWCyclesPair) & p;
  if (has_superinitial) out |= kWeighted & p;
  return out;
}
// The effect of on_add_tr over a whole set of arcs, from the OR of their "events".  The sequential rule only ever
// moves a pair from its positive to its negative bit, masks with the same constant after every arc and re-derives
// ACYCLIC / INITIAL_ACYCLIC from TOP_SORTED, so the result depends on the set of events, not on their order
// (checked against the sequential replay by tests/test_host_api.py and the reverse parity tests).
enum ArcEvent : uint32_t {
  kEvNotAcceptor = 1, kEvIEps = 2, kEvEps = 4, kEvOEps = 8, kEvIUnsorted = 16, kEvOUnsorted = 32, kEvWeighted = 64,
  kEvNotTopSorted = 128
};
B200_HD uint32_t arc_events(StateId state, const Tr& tr, const Tr* prev) {
  uint32_t e = 0;
  if (tr.ilabel != tr.olabel) e |= kEvNotAcceptor;
  if (tr.ilabel == kEps) { e |= kEvIEps; if (tr.olabel == kEps) e |= kEvEps; }
  if (tr.olabel == kEps) e |= kEvOEps;
  if (prev) { if (prev->ilabel > tr.ilabel) e |= kEvIUnsorted; if (prev->olabel > tr.olabel) e |= kEvOUnsorted; }
  if (!w_is_zero(tr.weight) && !w_is_one(tr.weight)) e |= kEvWeighted;
  if (tr.nextstate <= state) e |= kEvNotTopSorted;
  return e;
}
inline uint64_t apply_arc_events(uint64_t p, uint32_t ev, bool any_arc) {
  if (!any_arc) return p;
  uint64_t out = p;
  if (ev & kEvNotAcceptor) { out |= kNotAcceptor; out &= ~kAcceptor; }
  if (ev & kEvIEps) { out |= kIEpsilons; out &= ~kNoIEpsilons; }
  if (ev & kEvEps) { out |= kEpsilons; out &= ~kNoEpsilons; }
  if (ev & kEvOEps) { out |= kOEpsilons; out &= ~kNoOEpsilons; }
  if (ev & kEvIUnsorted) { out |= kNotILabelSorted; out &= ~kILabelSorted; }
  if (ev & kEvOUnsorted) { out |= kNotOLabelSorted; out &= ~kOLabelSorted; }
  if (ev & kEvWeighted) { out |= kWeighted; out &= ~kUnweighted; }
  if (ev & kEvNotTopSorted) { out |= kNotTopSorted; out &= ~kTopSorted; }
  out &= kKeepOnAddArc | kAcceptor | kNoEpsilons | kNoIEpsilons | kNoOEpsilons | kILabelSorted | kOLabelSorted |
         kUnweighted | kTopSorted;
  if (out & kTopSorted) out |= kAcyclic | kInitialAcyclic;
  return out;
}
// algorithms/tr_sort.rs:21-28,39-46
inline uint64_t after_tr_sort(uint64_t p, bool ilabel) {
  uint64_t out = (p & kKeepOnArcSort) | (ilabel ? kILabelSorted : kOLabelSorted);
  if (p & kAcceptor) out |= (ilabel ? kOLabelSorted : kILabelSorted);
  return out;
}
}  // namespace props
}  // namespace b200
