// host_fst.h — host-side VectorFst<TropicalWeight> container behind the opaque CFst handle of the C-ABI.
//
// Mirrors the reference container's observable behaviour (paths relative to /root/reference):
//   rustfst/src/fst_impls/vector_fst/{data_structure,fst,mutable_fst}.rs  (state = final weight + arc vector,
//   property word maintained on every mutation), rustfst/src/fst_impls/vector_fst/serializable_fst.rs (OpenFst
//   binary "vector" format), rustfst/src/fst_traits/macros.rs (text display).
//
// Layout is B200-first rather than a Vec<Arc<Vec<Tr>>>: the canonical representation is one CSR block
// (u32 state->arc offsets, 16-byte arcs, f32 final weights) that is copied to/from HBM with three plain
// memcpys.  Incremental mutation through the FFI (vec_fst_add_tr & co) uses a per-state "builder"
// representation that is frozen to CSR on first algorithmic use.
#pragma once
#include <algorithm>
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstring>
#include <fstream>
#include <mutex>
#include <new>
#include <type_traits>
#include <utility>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "fst_types.h"

namespace b200 {

struct FstError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// Runs fn(begin, end) over [0, n) on a few host threads (bulk copies of multi-hundred-megabyte machines: file parsing,
// serialisation, the repacking of batched compositions).
template <class F>
void parallel_ranges(size_t n, F fn, size_t min_parallel = 4096) {
  unsigned hw = std::thread::hardware_concurrency();
  size_t nt = std::min<size_t>(hw ? hw : 1, 8);
  if (n < min_parallel || nt <= 1) { fn((size_t)0, n); return; }
  std::vector<std::thread> th;
  std::exception_ptr err;
  std::mutex mu;
  const size_t chunk = (n + nt - 1) / nt;
  for (size_t t = 0; t < nt; t++) {
    const size_t b = t * chunk, e = std::min(n, b + chunk);
    if (b >= e) break;
    th.emplace_back([&, b, e] {
      try { fn(b, e); } catch (...) { std::lock_guard<std::mutex> g(mu); err = std::current_exception(); }
    });
  }
  for (auto& t : th) t.join();
  if (err) std::rethrow_exception(err);
}

// Large host arrays live in page-locked memory taken from a process-wide pool (device_common.cu) so that the three
// CSR arrays move to and from HBM at PCIe/C2C line rate with plain cudaMemcpyAsync; on a box without a CUDA device the
// pool degrades to malloc.  The allocator default-initialises (no zero fill of buffers that are overwritten anyway).
void* host_pool_alloc(size_t bytes);
void host_pool_free(void* p, size_t bytes) noexcept;

template <class T>
struct PoolAllocator {
  using value_type = T;
  PoolAllocator() noexcept = default;
  template <class U> PoolAllocator(const PoolAllocator<U>&) noexcept {}
  T* allocate(size_t n) { return static_cast<T*>(host_pool_alloc(n * sizeof(T))); }
  void deallocate(T* p, size_t n) noexcept { host_pool_free(p, n * sizeof(T)); }
  template <class U> void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) {
    ::new (static_cast<void*>(p)) U;  // default-init: trivially constructible types stay uninitialised
  }
  template <class U, class... Args> void construct(U* p, Args&&... args) {
    ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...);
  }
  template <class U> bool operator==(const PoolAllocator<U>&) const noexcept { return true; }
  template <class U> bool operator!=(const PoolAllocator<U>&) const noexcept { return false; }
};
template <class T> using PoolVec = std::vector<T, PoolAllocator<T>>;

// Plain CSR block: what travels to and from the device.
struct CsrFst {
  PoolVec<uint32_t> offsets{0};  // num_states + 1
  PoolVec<Tr> arcs;
  PoolVec<float> finals;         // +inf == not final
  std::vector<StateId> inf_finals;   // states that are final with weight +inf (Some(inf)); sorted, almost always empty
  bool has_start = false;
  StateId start = 0;
  uint64_t props = props::kNull;
  size_t num_states() const { return finals.size(); }
};

// Every algorithm (host or device) indexes per-state arrays with `start` and with every arc's `nextstate`; a machine
// that arrives from outside (file image, CSR ingest) is checked once, here, before anything walks it.
inline void validate_state_ids(const CsrFst& c, const std::string& err) {
  const size_t n = c.num_states();
  if (c.has_start && (size_t)c.start >= n) throw FstError(err);
  if (c.offsets.size() != n + 1 || c.offsets[n] != c.arcs.size()) throw FstError(err);
  std::atomic<bool> bad{false};
  const Tr* arcs = c.arcs.data();
  parallel_ranges(c.arcs.size(), [&](size_t lo, size_t hi) {
    uint32_t mx = 0;
    for (size_t e = lo; e < hi; e++) mx = arcs[e].nextstate > mx ? arcs[e].nextstate : mx;
    if (hi > lo && (size_t)mx >= n) bad = true;
  }, 1 << 18);
  if (bad) throw FstError(err);
}

// All trinary properties of a machine, recomputed from its content — the result of compute_fst_properties(fst,
// all_properties(), .., use_stored = false) (rustfst/src/fst_properties/compute_fst_properties.rs:14-208).  The
// reference derives the DFS group (cyclic, initial-cyclic, accessible, coaccessible, SCC membership for the
// weighted-cycles test) from one Tarjan visit with roots start, 0, 1, 2, …; all of these are facts about the graph,
// not about the visiting order, so they are computed here by an iterative Tarjan over the CSR plus one forward
// reachability pass.  The label / weight / shape group is the reference's single pass over the arcs.
inline uint64_t compute_properties_all(const CsrFst& c) {
  using namespace props;
  const size_t n = c.num_states();
  const uint32_t* off = c.offsets.data();
  const Tr* arcs = c.arcs.data();
  uint64_t out = 0;
  std::vector<int32_t> scc(n, -1);
  if (!c.has_start) {
    // dfs_visit returns before visiting anything (dfs_visit.rs:103-109): the visitor's initial word stands and every
    // state keeps the same (unset) component
    out |= kAcyclic | kInitialAcyclic | kAccessible | kCoAccessible;
  } else {
    // ---- Tarjan SCC over all states (iterative), cycle detection
    std::vector<int32_t> index(n, -1), low(n, 0);
    std::vector<uint8_t> onstack(n, 0);
    std::vector<uint32_t> tstack, cursor(n, 0);
    struct Frame { uint32_t s; };
    std::vector<uint32_t> call;
    int32_t next_index = 0, nscc = 0;
    bool cyclic = false;
    for (size_t root = 0; root < n; root++) {
      if (index[root] >= 0) continue;
      call.push_back((uint32_t)root);
      index[root] = low[root] = next_index++;
      tstack.push_back((uint32_t)root); onstack[root] = 1; cursor[root] = off[root];
      while (!call.empty()) {
        const uint32_t s = call.back();
        if (cursor[s] < off[s + 1]) {
          const uint32_t t = arcs[cursor[s]++].nextstate;
          if (index[t] < 0) {
            index[t] = low[t] = next_index++;
            tstack.push_back(t); onstack[t] = 1; cursor[t] = off[t];
            call.push_back(t);
          } else if (onstack[t]) {
            if (index[t] < low[s]) low[s] = index[t];
          }
        } else {
          call.pop_back();
          if (!call.empty() && low[s] < low[call.back()]) low[call.back()] = low[s];
          if (low[s] == index[s]) {
            size_t members = 0;
            uint32_t t;
            do { t = tstack.back(); tstack.pop_back(); onstack[t] = 0; scc[t] = nscc; members++; } while (t != s);
            if (members > 1) cyclic = true;
            nscc++;
          }
        }
      }
    }
    for (size_t s = 0; s < n && !cyclic; s++)
      for (uint32_t e = off[s]; e < off[s + 1]; e++) if (arcs[e].nextstate == s) { cyclic = true; break; }
    out |= cyclic ? kCyclic : kAcyclic;
    // the start state lies on a cycle iff it has a self loop or shares its component with another state
    bool initial_cyclic = false;
    for (uint32_t e = off[c.start]; e < off[c.start + 1]; e++) if (arcs[e].nextstate == c.start) initial_cyclic = true;
    if (!initial_cyclic)
      for (size_t s = 0; s < n; s++) if (s != c.start && scc[s] == scc[c.start]) { initial_cyclic = true; break; }
    out |= initial_cyclic ? kInitialCyclic : kInitialAcyclic;
    // ---- accessible: forward reachability from the start state
    std::vector<uint8_t> acc(n, 0);
    std::vector<uint32_t> work{c.start};
    acc[c.start] = 1;
    size_t n_acc = 1;
    while (!work.empty()) {
      const uint32_t s = work.back(); work.pop_back();
      for (uint32_t e = off[s]; e < off[s + 1]; e++) {
        const uint32_t t = arcs[e].nextstate;
        if (!acc[t]) { acc[t] = 1; n_acc++; work.push_back(t); }
      }
    }
    out |= (n_acc == n) ? kAccessible : kNotAccessible;
    // ---- coaccessible: Tarjan numbers components in reverse topological order (successors first), so one sweep over
    // the components in creation order propagates "reaches a final state" from successors to predecessors
    std::vector<uint8_t> comp_co((size_t)nscc, 0);
    std::vector<std::vector<uint32_t>> by_comp((size_t)nscc);
    for (size_t s = 0; s < n; s++) by_comp[scc[s]].push_back((uint32_t)s);
    bool all_co = true;
    for (int32_t k = 0; k < nscc; k++) {
      bool co = false;
      for (uint32_t s : by_comp[k]) {
        if (c.finals[s] != w_zero() || std::binary_search(c.inf_finals.begin(), c.inf_finals.end(), s)) co = true;
        for (uint32_t e = off[s]; e < off[s + 1] && !co; e++) {
          const int32_t kt = scc[arcs[e].nextstate];
          if (kt != k && comp_co[kt]) co = true;
        }
        if (co) break;
      }
      comp_co[k] = co;
      if (!co) all_co = false;
    }
    out |= all_co ? kCoAccessible : kNotCoAccessible;
  }
  // ---- label / weight / shape group (compute_fst_properties.rs:56-196)
  out |= kAcceptor | kNoEpsilons | kNoIEpsilons | kNoOEpsilons | kILabelSorted | kOLabelSorted | kUnweighted |
         kTopSorted | kString | kIDeterministic | kODeterministic | kUnweightedCycles;
  auto flip = [&](uint64_t neg, uint64_t pos) { out |= neg; out &= ~pos; };
  size_t nfinal = 0;
  std::vector<Label> il, ol;
  for (size_t s = 0; s < n; s++) {
    const uint32_t lo = off[s], hi = off[s + 1];
    il.clear(); ol.clear();
    for (uint32_t e = lo; e < hi; e++) {
      const Tr& tr = arcs[e];
      il.push_back(tr.ilabel); ol.push_back(tr.olabel);
      if (tr.ilabel != tr.olabel) flip(kNotAcceptor, kAcceptor);
      if (tr.ilabel == kEps && tr.olabel == kEps) flip(kEpsilons, kNoEpsilons);
      if (tr.ilabel == kEps) flip(kIEpsilons, kNoIEpsilons);
      if (tr.olabel == kEps) flip(kOEpsilons, kNoOEpsilons);
      if (e > lo) {
        if (tr.ilabel < arcs[e - 1].ilabel) flip(kNotILabelSorted, kILabelSorted);
        if (tr.olabel < arcs[e - 1].olabel) flip(kNotOLabelSorted, kOLabelSorted);
      }
      if (!w_is_one(tr.weight) && !w_is_zero(tr.weight)) {
        flip(kWeighted, kUnweighted);
        if ((out & kUnweightedCycles) && scc[s] == scc[tr.nextstate]) flip(kWeightedCycles, kUnweightedCycles);
      }
      if (tr.nextstate <= s) flip(kNotTopSorted, kTopSorted);
      if (tr.nextstate != s + 1) flip(kNotString, kString);
    }
    std::sort(il.begin(), il.end()); std::sort(ol.begin(), ol.end());
    if (std::adjacent_find(il.begin(), il.end()) != il.end()) flip(kNotIDeterministic, kIDeterministic);
    if (std::adjacent_find(ol.begin(), ol.end()) != ol.end()) flip(kNotODeterministic, kODeterministic);
    if (nfinal > 0) flip(kNotString, kString);
    const bool inf_final = std::binary_search(c.inf_finals.begin(), c.inf_finals.end(), (StateId)s);
    if (c.finals[s] != w_zero() || inf_final) {
      if (!w_is_one(c.finals[s])) flip(kWeighted, kUnweighted);
      nfinal++;
    } else if (hi - lo != 1) {
      flip(kNotString, kString);
    }
  }
  if (c.has_start && c.start != 0) flip(kNotString, kString);
  return out;
}

// isomorphic (rustfst/src/algorithms/isomorphic.rs:49-160, delta = KDELTA): breadth-first pairing of states from the
// two start states; the arcs of a pair are compared after sorting by (ilabel, olabel, weight, nextstate).  Kept verbatim,
// including the one-permutation limitation (an error, not `false`, when a mismatch follows arcs that are equal as an
// unweighted automaton).  Used as the result verifier SURVEY.md §8f names.
inline bool isomorphic(const CsrFst& a, const CsrFst& b, float delta = kDelta) {
  if (!a.has_start && !b.has_start) return true;
  if (!a.has_start || !b.has_start) return false;
  auto approx = [delta](float x, float y) { return std::fabs(x - y) <= delta; };  // TropicalWeight::approx_equal
  auto final_of = [](const CsrFst& f, StateId s, float* w) {
    if (f.finals[s] != w_zero()) { *w = f.finals[s]; return true; }
    if (std::binary_search(f.inf_finals.begin(), f.inf_finals.end(), s)) { *w = w_zero(); return true; }
    return false;
  };
  auto less = [](const Tr& x, const Tr& y) {
    if (x.ilabel != y.ilabel) return x.ilabel < y.ilabel;
    if (x.olabel != y.olabel) return x.olabel < y.olabel;
    if (x.weight < y.weight) return true;
    if (x.weight > y.weight) return false;
    return x.nextstate < y.nextstate;
  };
  std::vector<int64_t> pair(a.num_states(), -1);
  std::vector<std::pair<StateId, StateId>> queue;
  size_t head = 0;
  bool non_det = false;
  auto pair_state = [&](StateId s1, StateId s2) {
    if (pair[s1] == (int64_t)s2) return true;
    if (pair[s1] >= 0) return false;
    pair[s1] = s2;
    queue.emplace_back(s1, s2);
    return true;
  };
  pair_state(a.start, b.start);
  std::vector<Tr> t1, t2;
  while (head < queue.size()) {
    const auto [s1, s2] = queue[head++];
    bool ok = true;
    float w1 = 0, w2 = 0;
    const bool f1 = final_of(a, s1, &w1), f2 = final_of(b, s2, &w2);
    if (f1 != f2 || (f1 && !approx(w1, w2))) ok = false;
    if (ok && a.offsets[s1 + 1] - a.offsets[s1] != b.offsets[s2 + 1] - b.offsets[s2]) ok = false;
    if (ok) {
      t1.assign(a.arcs.begin() + a.offsets[s1], a.arcs.begin() + a.offsets[s1 + 1]);
      t2.assign(b.arcs.begin() + b.offsets[s2], b.arcs.begin() + b.offsets[s2 + 1]);
      std::stable_sort(t1.begin(), t1.end(), less);
      std::stable_sort(t2.begin(), t2.end(), less);
      for (size_t i = 0; i < t1.size() && ok; i++) {
        if (t1[i].ilabel != t2[i].ilabel || t1[i].olabel != t2[i].olabel || !approx(t1[i].weight, t2[i].weight) ||
            !pair_state(t1[i].nextstate, t2[i].nextstate)) {
          ok = false;
          break;
        }
        if (i > 0 && t1[i].ilabel == t1[i - 1].ilabel && t1[i].olabel == t1[i - 1].olabel &&
            approx(t1[i].weight, t1[i - 1].weight))
          non_det = true;
      }
    }
    if (!ok) {
      if (non_det)
        throw FstError("Isomorphic: Non-determinism as an unweighted automaton. state1 = " + std::to_string(s1) +
                       " state2 = " + std::to_string(s2));
      return false;
    }
  }
  return true;
}

class HostFst {
 public:
  HostFst() = default;
  explicit HostFst(CsrFst&& csr) : csr_(std::move(csr)), is_builder_(false) {
    has_start_ = csr_.has_start; start_ = csr_.start; props_ = csr_.props;
  }
  HostFst(const HostFst& o) {
    std::lock_guard<std::mutex> g(o.mu_);
    b_ = o.b_; csr_ = o.csr_; is_builder_ = o.is_builder_;
    has_start_ = o.has_start_; start_ = o.start_; props_ = o.props_; max_next_ = o.max_next_;
  }

  // ---- inspection (fst_impls/vector_fst/fst.rs:40-106)
  size_t num_states() const { std::lock_guard<std::mutex> g(mu_); return n_states(); }
  bool start(StateId* out) const { std::lock_guard<std::mutex> g(mu_); if (has_start_) *out = start_; return has_start_; }
  uint64_t properties() const { std::lock_guard<std::mutex> g(mu_); return props_; }
  void set_properties(uint64_t p) { std::lock_guard<std::mutex> g(mu_); props_ = p & props::kTrinary; }
  bool final_weight(StateId s, float* out) const {
    std::lock_guard<std::mutex> g(mu_);
    check_state(s);
    if (is_builder_) { if (b_[s].has_final) *out = b_[s].final_w; return b_[s].has_final; }
    if (csr_.finals[s] != w_zero()) { *out = csr_.finals[s]; return true; }
    if (is_inf_final(s)) { *out = w_zero(); return true; }
    return false;
  }
  size_t num_trs(StateId s) const {
    std::lock_guard<std::mutex> g(mu_);
    check_state(s);
    return is_builder_ ? b_[s].trs.size() : (size_t)(csr_.offsets[s + 1] - csr_.offsets[s]);
  }
  std::vector<Tr> get_trs(StateId s) const {
    std::lock_guard<std::mutex> g(mu_);
    check_state(s);
    if (is_builder_) return b_[s].trs;
    return std::vector<Tr>(csr_.arcs.begin() + csr_.offsets[s], csr_.arcs.begin() + csr_.offsets[s + 1]);
  }

  // ---- mutation (fst_impls/vector_fst/mutable_fst.rs)
  StateId add_state() {  // :79-84
    std::lock_guard<std::mutex> g(mu_);
    StateId id = (StateId)n_states();
    if (is_builder_) b_.emplace_back();
    else { csr_.offsets.push_back(csr_.offsets.back()); csr_.finals.push_back(w_zero()); }
    props_ = props::on_add_state(props_);
    return id;
  }
  void set_start(StateId s) {  // :35-44
    std::lock_guard<std::mutex> g(mu_);
    if (s >= n_states()) throw FstError("The state " + std::to_string(s) + " doesn't exist");
    has_start_ = true; start_ = s;
    props_ = props::on_set_start(props_);
  }
  void set_final(StateId s, float w) {  // :51-64
    std::lock_guard<std::mutex> g(mu_);
    if (s >= n_states()) throw FstError("Stateid " + std::to_string(s) + " doesn't exist");
    float old_w; bool had = final_unlocked(s, &old_w);
    props_ = props::on_set_final(props_, had ? &old_w : nullptr, &w);
    set_final_unlocked(s, true, w);
  }
  void delete_final_weight(StateId s) {  // :283-291
    std::lock_guard<std::mutex> g(mu_);
    if (s >= n_states()) throw FstError("State " + std::to_string(s) + " doesn't exist");
    float old_w; bool had = final_unlocked(s, &old_w);
    props_ = props::on_set_final(props_, had ? &old_w : nullptr, nullptr);
    set_final_unlocked(s, false, w_zero());
  }
  void add_tr(StateId s, const Tr& tr) {  // :236-245 + data_structure.rs:80-91
    std::lock_guard<std::mutex> g(mu_);
    if (s >= n_states()) throw FstError("State " + std::to_string(s) + " doesn't exist");
    const Tr* prev = nullptr;
    Tr prev_copy;
    if (!is_builder_ && (size_t)s + 1 == n_states()) {  // append to the tail state keeps the CSR valid
      if (csr_.offsets[s + 1] > csr_.offsets[s]) { prev_copy = csr_.arcs.back(); prev = &prev_copy; }
      csr_.arcs.push_back(tr);
      csr_.offsets[s + 1]++;
    } else {
      to_builder_unlocked();
      if (!b_[s].trs.empty()) { prev_copy = b_[s].trs.back(); prev = &prev_copy; }
      b_[s].trs.push_back(tr);
    }
    if (tr.nextstate > max_next_) max_next_ = tr.nextstate;
    props_ = props::on_add_tr(props_, s, tr, prev);
  }
  void set_tr(StateId s, size_t idx, const Tr& tr) {  // trs_iter_mut: properties of a set arc are unknown
    std::lock_guard<std::mutex> g(mu_);
    check_state(s);
    to_builder_unlocked();
    if (idx >= b_[s].trs.size()) throw FstError("transition index out of range");
    b_[s].trs[idx] = tr;
    if (tr.nextstate > max_next_) max_next_ = tr.nextstate;
    props_ &= props::kBinary;  // properties.rs set_arc_properties() == empty
  }
  void del_all_states() {  // :191-199
    std::lock_guard<std::mutex> g(mu_);
    b_.clear(); csr_ = CsrFst(); is_builder_ = true;
    has_start_ = false; props_ = props::kNull;
  }
  void tr_sort(bool ilabel) {  // algorithms/tr_sort.rs:51-62 (stable)
    std::lock_guard<std::mutex> g(mu_);
    auto by_i = [](const Tr& a, const Tr& b) { return a.ilabel < b.ilabel; };
    auto by_o = [](const Tr& a, const Tr& b) { return a.olabel < b.olabel; };
    size_t n = n_states();
    for (size_t s = 0; s < n; s++) {
      Tr *b, *e;
      if (is_builder_) { b = b_[s].trs.data(); e = b + b_[s].trs.size(); }
      else { b = csr_.arcs.data() + csr_.offsets[s]; e = csr_.arcs.data() + csr_.offsets[s + 1]; }
      if (ilabel) std::stable_sort(b, e, by_i); else std::stable_sort(b, e, by_o);
    }
    props_ = props::after_tr_sort(props_, ilabel) & props::kTrinary;
  }

  // state_sort (algorithms/state_sort.rs:16-78): old state s becomes state order[s]; arcs keep their order, next states
  // are mapped; the property word keeps only statesort_properties() (properties.rs:319-349).
  void state_sort(const std::vector<uint32_t>& order) {
    const CsrFst& c = checked();
    std::lock_guard<std::mutex> g(mu_);
    const size_t n = c.num_states();
    if (order.size() != n)
      throw FstError("StateSort : Bad order vector size : " + std::to_string(order.size()) + ". Expected " +
                     std::to_string(n));
    if (!has_start_) return;
    CsrFst o;
    o.offsets.resize(n + 1);
    o.finals.resize(n);
    o.arcs.resize(c.arcs.size());
    std::vector<uint32_t> inv(n);
    for (size_t s = 0; s < n; s++) inv[order[s]] = (uint32_t)s;
    size_t w = 0;
    for (size_t t = 0; t < n; t++) {
      const uint32_t s = inv[t];
      o.offsets[t] = (uint32_t)w;
      o.finals[t] = c.finals[s];
      for (uint32_t e = c.offsets[s]; e < c.offsets[s + 1]; e++) {
        Tr tr = c.arcs[e];
        tr.nextstate = order[tr.nextstate];
        o.arcs[w++] = tr;
      }
    }
    o.offsets[n] = (uint32_t)w;
    for (StateId s : c.inf_finals) o.inf_finals.push_back(order[s]);
    std::sort(o.inf_finals.begin(), o.inf_finals.end());
    start_ = order[start_];
    props_ &= props::kTrinary & ~(props::kTopPair | props::kStringPair);
    o.has_start = true; o.start = start_; o.props = props_;
    csr_ = std::move(o);
  }
  void or_properties(uint64_t bits) { std::lock_guard<std::mutex> g(mu_); props_ |= bits & props::kTrinary; }

  // compute_and_update_properties_all (fst_traits/mutable_fst.rs:435-446): the stored word is returned untouched when
  // every property is already known (use_stored = true), otherwise everything is recomputed.
  uint64_t compute_and_update_properties_all() {
    const CsrFst& c = checked();
    std::lock_guard<std::mutex> g(mu_);
    if ((props::known(props_) & props::kAll) != props::kAll) props_ = compute_properties_all(c) & props::kTrinary;
    return props_;
  }

  // ---- CSR access for the device path. freeze() makes CSR the live representation.
  const CsrFst& freeze() const {
    std::lock_guard<std::mutex> g(mu_);
    to_csr_unlocked();
    csr_.has_start = has_start_; csr_.start = start_; csr_.props = props_;
    return csr_;
  }
  // freeze() for an algorithm: add_tr / set_tr accept any `nextstate` (as the reference does), so before anything
  // indexes by it the machine is checked for arcs into states that do not exist.  max_next_ is an upper bound of every
  // nextstate added since the last full check, so the scan only runs when it can fail.
  const CsrFst& checked() const {
    const CsrFst& c = freeze();
    std::lock_guard<std::mutex> g(mu_);
    const size_t n = c.num_states();
    if (!c.arcs.empty() && (size_t)max_next_ >= n) {
      StateId mx = 0;
      for (const Tr& t : c.arcs) mx = t.nextstate > mx ? t.nextstate : mx;
      if ((size_t)mx >= n)
        throw FstError("transition to state " + std::to_string(mx) + " but the FST has only " + std::to_string(n) + " states");
      max_next_ = mx;
    }
    return c;
  }
  void replace(CsrFst&& csr) {  // in-place algorithms (fst_connect) install their result
    std::lock_guard<std::mutex> g(mu_);
    csr_ = std::move(csr); b_.clear(); is_builder_ = false;
    has_start_ = csr_.has_start; start_ = csr_.start; props_ = csr_.props;
  }

  // ---- equality: data_structure.rs:36-41 (ignores properties/symbol tables, weights approx-equal)
  bool equals(const HostFst& o) const {
    const CsrFst& a = freeze();
    const CsrFst& b = o.freeze();
    if (a.has_start != b.has_start || (a.has_start && a.start != b.start)) return false;
    if (a.num_states() != b.num_states()) return false;
    if (a.inf_finals != b.inf_finals) return false;
    for (size_t s = 0; s < a.num_states(); s++) {
      bool fa = a.finals[s] != w_zero(), fb = b.finals[s] != w_zero();
      if (fa != fb) return false;
      if (fa && !w_approx_eq(a.finals[s], b.finals[s])) return false;
      uint32_t na = a.offsets[s + 1] - a.offsets[s], nb = b.offsets[s + 1] - b.offsets[s];
      if (na != nb) return false;
      const Tr* x = a.arcs.data() + a.offsets[s];
      const Tr* y = b.arcs.data() + b.offsets[s];
      for (uint32_t i = 0; i < na; i++)
        if (x[i].ilabel != y[i].ilabel || x[i].olabel != y[i].olabel || x[i].nextstate != y[i].nextstate ||
            !w_approx_eq(x[i].weight, y[i].weight))
          return false;
    }
    return true;
  }

  // ---- text form: fst_traits/macros.rs:1-70 (write_fst!(self, f, show_weight_one=true, use_symt=true))
  std::string display() const {
    const CsrFst& c = freeze();
    std::string out;
    if (!c.has_start) return out;
    auto fmt_w = [](float w) {
      if (w == w_zero()) return std::string("inf");
      char buf[64];
      auto r = std::to_chars(buf, buf + sizeof(buf), w, std::chars_format::fixed);
      return std::string(buf, r.ptr);
    };
    auto one_state = [&](StateId s) {
      for (uint32_t i = c.offsets[s]; i < c.offsets[s + 1]; i++) {
        const Tr& t = c.arcs[i];
        out += std::to_string(s) + "\t" + std::to_string(t.nextstate) + "\t" + std::to_string(t.ilabel) + "\t" +
               std::to_string(t.olabel) + "\t" + fmt_w(t.weight) + "\n";
      }
    };
    one_state(c.start);
    for (size_t s = 0; s < c.num_states(); s++) if (s != c.start) one_state((StateId)s);
    for (size_t s = 0; s < c.num_states(); s++) {
      bool inf_final = std::binary_search(c.inf_finals.begin(), c.inf_finals.end(), (StateId)s);
      if (c.finals[s] != w_zero() || inf_final) out += std::to_string(s) + "\t" + fmt_w(c.finals[s]) + "\n";
    }
    return out;
  }

 private:
  struct BState {
    bool has_final = false;
    float final_w = 0.0f;
    std::vector<Tr> trs;
  };
  size_t n_states() const { return is_builder_ ? b_.size() : csr_.finals.size(); }
  void check_state(StateId s) const {
    if (s >= n_states()) throw FstError("State " + std::to_string(s) + " doesn't exist");
  }
  bool is_inf_final(StateId s) const {
    return !csr_.inf_finals.empty() && std::binary_search(csr_.inf_finals.begin(), csr_.inf_finals.end(), s);
  }
  bool final_unlocked(StateId s, float* w) const {
    if (is_builder_) { *w = b_[s].final_w; return b_[s].has_final; }
    if (csr_.finals[s] != w_zero()) { *w = csr_.finals[s]; return true; }
    if (is_inf_final(s)) { *w = w_zero(); return true; }
    return false;
  }
  void set_final_unlocked(StateId s, bool has, float w) {
    if (is_builder_) { b_[s].has_final = has; b_[s].final_w = w; return; }
    auto it = std::lower_bound(csr_.inf_finals.begin(), csr_.inf_finals.end(), s);
    bool present = it != csr_.inf_finals.end() && *it == s;
    bool want_inf = has && w == w_zero();
    if (want_inf && !present) csr_.inf_finals.insert(it, s);
    if (!want_inf && present) csr_.inf_finals.erase(it);
    csr_.finals[s] = has ? w : w_zero();
  }
  void to_builder_unlocked() const {
    if (is_builder_) return;
    size_t n = csr_.finals.size();
    b_.assign(n, BState());
    for (size_t s = 0; s < n; s++) {
      b_[s].trs.assign(csr_.arcs.begin() + csr_.offsets[s], csr_.arcs.begin() + csr_.offsets[s + 1]);
      if (csr_.finals[s] != w_zero()) { b_[s].has_final = true; b_[s].final_w = csr_.finals[s]; }
    }
    for (StateId s : csr_.inf_finals) { b_[s].has_final = true; b_[s].final_w = w_zero(); }
    csr_ = CsrFst();
    is_builder_ = true;
  }
  void to_csr_unlocked() const {
    if (!is_builder_) return;
    size_t n = b_.size(), total = 0;
    for (auto& st : b_) total += st.trs.size();
    if (total > 0xFFFFFFF0ull) throw FstError("FST has more than 2^32 transitions");
    CsrFst c;
    c.offsets.resize(n + 1);
    c.arcs.resize(total);
    c.finals.resize(n);
    size_t o = 0;
    for (size_t s = 0; s < n; s++) {
      c.offsets[s] = (uint32_t)o;
      if (!b_[s].trs.empty()) std::memcpy(c.arcs.data() + o, b_[s].trs.data(), b_[s].trs.size() * sizeof(Tr));
      o += b_[s].trs.size();
      c.finals[s] = b_[s].has_final ? b_[s].final_w : w_zero();
      if (b_[s].has_final && b_[s].final_w == w_zero()) c.inf_finals.push_back((StateId)s);
    }
    c.offsets[n] = (uint32_t)o;
    csr_ = std::move(c);
    b_.clear(); b_.shrink_to_fit();
    is_builder_ = false;
  }

  mutable std::mutex mu_;
  mutable std::vector<BState> b_;
  mutable CsrFst csr_;
  mutable bool is_builder_ = true;
  mutable StateId max_next_ = 0;  // upper bound of the nextstate of every arc added through add_tr / set_tr
  bool has_start_ = false;
  StateId start_ = 0;
  uint64_t props_ = props::kNull;
};

// ---------------------------------------------------------------------------------------------------------------
// OpenFst binary "vector" format — rustfst/src/parsers/bin_fst/fst_header.rs:71-137,
// rustfst/src/fst_impls/vector_fst/serializable_fst.rs:46-88 (store), :129-168 (load)
// ---------------------------------------------------------------------------------------------------------------
namespace io {
constexpr int32_t kMagic = 2125659606;
constexpr int32_t kSymtMagic = 2125658996;

struct Reader {
  const uint8_t* p; size_t n; size_t off = 0;
  template <class T> T get() {
    if (off + sizeof(T) > n) throw FstError("Error while parsing binary VectorFst. Error kind Eof");
    T v; std::memcpy(&v, p + off, sizeof(T)); off += sizeof(T); return v;
  }
  std::string str() {
    int32_t len = get<int32_t>();
    if (len < 0 || off + (size_t)len > n) throw FstError("Error while parsing binary VectorFst. Error kind Eof");
    std::string s((const char*)p + off, (size_t)len); off += (size_t)len; return s;
  }
};
// Symbol tables are presentation metadata outside the hot path: parsed to find the body, then dropped.
inline void skip_symbol_table(Reader& r) {  // parsers/bin_symt/nom_parser.rs
  if (r.get<int32_t>() != kSymtMagic) throw FstError("Error while parsing symbolTable from binary VectorFst");
  r.str();
  r.get<int64_t>();
  int64_t n = r.get<int64_t>();
  for (int64_t i = 0; i < n; i++) { r.str(); r.get<int64_t>(); }
}

inline CsrFst parse_vector_fst(const uint8_t* data, size_t len) {
  Reader r{data, len};
  if (r.get<int32_t>() != kMagic) throw FstError("Error while parsing binary VectorFst. Error kind Verify");
  if (r.str() != "vector") throw FstError("Error while parsing binary VectorFst. Error kind Verify");
  if (r.str() != "standard") throw FstError("Error while parsing binary VectorFst. Error kind Verify");
  if (r.get<int32_t>() < 2) throw FstError("Error while parsing binary VectorFst. Error kind Verify");
  uint32_t flags = r.get<uint32_t>();
  if (flags & ~7u) throw FstError("Error while parsing binary VectorFst. Error kind MapRes");
  uint64_t props_word = r.get<uint64_t>();
  int64_t start = r.get<int64_t>();
  int64_t num_states = r.get<int64_t>();
  r.get<int64_t>();  // header num_arcs is unreliable for vector files (OpenFst writes 0); per-state counts rule
  if (flags & 1) skip_symbol_table(r);
  if (flags & 2) skip_symbol_table(r);
  if (num_states < 0) throw FstError("Error while parsing binary VectorFst. Error kind Count");
  // every state record takes at least 12 bytes: a header that claims more states than the file can hold is rejected
  // before anything is allocated from it
  if ((uint64_t)num_states > (uint64_t)(r.n - r.off) / 12) throw FstError("Error while parsing binary VectorFst. Error kind Eof");
  CsrFst c;
  c.props = props_word & props::kTrinary;  // FstProperties::from_bits_truncate
  c.has_start = start != -1;
  c.start = (StateId)start;
  c.offsets.resize((size_t)num_states + 1);
  c.finals.resize((size_t)num_states);
  // One pass to size (sequential: every state record tells where the next one starts), one pass to copy (host
  // threads): the arc records are already in the 16-byte device layout.
  std::vector<uint64_t> rec_at((size_t)num_states);  // byte offset of every state record
  size_t total = 0;
  {
    Reader q = r;
    for (int64_t s = 0; s < num_states; s++) {
      rec_at[s] = q.off;
      q.get<float>();
      int64_t na = q.get<int64_t>();
      if (na < 0 || q.off + (size_t)na * 16 > q.n) throw FstError("Error while parsing binary VectorFst. Error kind Eof");
      c.offsets[s] = (uint32_t)total;
      q.off += (size_t)na * 16;
      total += (size_t)na;
      if (total > 0xFFFFFFF0ull) throw FstError("FST has more than 2^32 transitions");
    }
  }
  c.offsets[num_states] = (uint32_t)total;
  c.arcs.resize(total);
  const uint8_t* base = r.p;
  parallel_ranges((size_t)num_states, [&](size_t lo, size_t hi) {
    for (size_t s = lo; s < hi; s++) {
      float fw;
      std::memcpy(&fw, base + rec_at[s], 4);
      // parsers/bin_fst/utils_parsing.rs:17-26: None iff approx-equal to zero() (i.e. +inf)
      c.finals[s] = w_approx_eq(fw, w_zero()) ? w_zero() : fw;
      const size_t na = c.offsets[s + 1] - c.offsets[s];
      if (na) std::memcpy(c.arcs.data() + c.offsets[s], base + rec_at[s] + 12, na * 16);
    }
  }, 1 << 16);
  // the reference is memory-safe on files whose start / nextstate fields point outside the machine (its algorithms
  // panic); here such a file is refused at the door
  validate_state_ids(c, "Error while parsing binary VectorFst. Error kind Verify");
  return c;
}

constexpr size_t kVectorHeaderBytes = 4 + (4 + 6) + (4 + 8) + 4 + 4 + 8 + 8 + 8 + 8;
inline size_t vector_fst_bytes(const CsrFst& c) { return kVectorHeaderBytes + c.num_states() * 12 + c.arcs.size() * 16; }
// Serialises into caller-provided memory of vector_fst_bytes(c) bytes (no intermediate buffer, no zero fill).
inline void store_vector_fst_into(const CsrFst& c, uint8_t* out) {
  const size_t n = c.num_states();
  // header: magic, "vector", "standard", version, flags, properties, start, #states, #arcs
  uint8_t* w = out;
  auto put = [&](const void* p, size_t k) { std::memcpy(w, p, k); w += k; };
  auto put_i32 = [&](int32_t v) { put(&v, 4); };
  auto put_i64 = [&](int64_t v) { put(&v, 8); };
  auto put_str = [&](const char* s) { int32_t l = (int32_t)std::strlen(s); put_i32(l); put(s, (size_t)l); };
  put_i32(kMagic);
  put_str("vector");
  put_str("standard");
  put_i32(2);
  uint32_t flags = 0; put(&flags, 4);
  uint64_t p = c.props | props::kExpanded | props::kMutable; put(&p, 8);
  put_i64(c.has_start ? (int64_t)c.start : -1);
  put_i64((int64_t)n);
  put_i64((int64_t)c.arcs.size());
  uint8_t* body = w;  // state s starts at body + 12 * s + 16 * offsets[s]
  parallel_ranges(n, [&](size_t lo, size_t hi) {
    for (size_t s = lo; s < hi; s++) {
      uint8_t* q = body + 12 * s + 16 * (size_t)c.offsets[s];
      const int64_t na = c.offsets[s + 1] - c.offsets[s];
      std::memcpy(q, &c.finals[s], 4);
      std::memcpy(q + 4, &na, 8);
      if (na) std::memcpy(q + 12, c.arcs.data() + c.offsets[s], (size_t)na * 16);
    }
  }, 1 << 16);
}
inline std::vector<uint8_t> store_vector_fst(const CsrFst& c) {
  std::vector<uint8_t> buf(vector_fst_bytes(c));
  store_vector_fst_into(c, buf.data());
  return buf;
}

// OpenFst binary "const" format — rustfst/src/fst_impls/const_fst/serializable_fst.rs:30-79 (store), :180-236 (load),
// const_fst/mod.rs:11-14 (versions: 1 = aligned to 16 bytes, 2 = packed).  A const FST *is* a CSR block (one state array
// {final weight, first arc, #arcs, #input eps, #output eps}, one arc array), so both directions are straight copies.
inline CsrFst parse_const_fst(const uint8_t* data, size_t len) {
  const char* kErr = "Error while parsing binary ConstFst";
  Reader r{data, len};
  try {
    if (r.get<int32_t>() != kMagic) throw FstError(kErr);
    if (r.str() != "const") throw FstError(kErr);
    if (r.str() != "standard") throw FstError(kErr);
    const int32_t version = r.get<int32_t>();
    if (version < 1) throw FstError(kErr);
    const uint32_t flags = r.get<uint32_t>();
    if (flags & ~7u) throw FstError(kErr);
    const uint64_t props_word = r.get<uint64_t>();
    const int64_t start = r.get<int64_t>();
    const int64_t num_states = r.get<int64_t>();
    const int64_t num_trs = r.get<int64_t>();
    if (flags & 1) skip_symbol_table(r);
    if (flags & 2) skip_symbol_table(r);
    if (num_states < 0 || num_trs < 0 || (uint64_t)num_trs > 0xFFFFFFF0ull) throw FstError(kErr);
    // 20 bytes per state record, 16 per arc: reject impossible counts before allocating from them
    if ((uint64_t)num_states > (uint64_t)(r.n - r.off) / 20 || (uint64_t)num_trs > (uint64_t)(r.n - r.off) / 16) throw FstError(kErr);
    const bool aligned = version == 1;
    auto align = [&](int64_t count) {
      if (aligned && count > 0 && r.off % 16 != 0) {
        r.off += 16 - r.off % 16;
        if (r.off > r.n) throw FstError(kErr);
      }
    };
    align(num_states);
    struct St { float fw; int32_t pos, ntrs, nie, noe; };
    std::vector<St> sts((size_t)num_states);
    for (int64_t s = 0; s < num_states; s++) {
      sts[s].fw = r.get<float>(); sts[s].pos = r.get<int32_t>(); sts[s].ntrs = r.get<int32_t>();
      sts[s].nie = r.get<int32_t>(); sts[s].noe = r.get<int32_t>();
    }
    align(num_trs);
    if (r.off + (size_t)num_trs * 16 > r.n) throw FstError(kErr);
    const uint8_t* arcs = r.p + r.off;
    CsrFst c;
    c.props = props_word & props::kTrinary;  // FstProperties::from_bits_truncate
    c.has_start = start != -1;
    c.start = (StateId)start;
    c.offsets.resize((size_t)num_states + 1);
    c.finals.resize((size_t)num_states);
    size_t total = 0;
    for (int64_t s = 0; s < num_states; s++) {
      if (sts[s].pos < 0 || sts[s].ntrs < 0 || (int64_t)sts[s].pos + sts[s].ntrs > num_trs) throw FstError(kErr);
      total += (size_t)sts[s].ntrs;
    }
    c.arcs.resize(total);
    size_t o = 0;
    for (int64_t s = 0; s < num_states; s++) {
      c.finals[s] = w_approx_eq(sts[s].fw, w_zero()) ? w_zero() : sts[s].fw;  // utils_parsing.rs:17-26
      c.offsets[s] = (uint32_t)o;
      if (sts[s].ntrs) std::memcpy(c.arcs.data() + o, arcs + (size_t)sts[s].pos * 16, (size_t)sts[s].ntrs * 16);
      o += (size_t)sts[s].ntrs;
    }
    c.offsets[num_states] = (uint32_t)o;
    validate_state_ids(c, kErr);
    return c;
  } catch (const FstError&) {
    throw FstError(kErr);  // load() maps every parse failure to this message (serializable_fst.rs:36-40)
  }
}

inline std::vector<uint8_t> store_const_fst(const CsrFst& c) {
  const size_t n = c.num_states();
  std::vector<uint8_t> buf;
  buf.reserve(64 + n * 20 + c.arcs.size() * 16);
  auto put = [&](const void* p, size_t k) { const uint8_t* b = (const uint8_t*)p; buf.insert(buf.end(), b, b + k); };
  auto put_i32 = [&](int32_t v) { put(&v, 4); };
  auto put_i64 = [&](int64_t v) { put(&v, 8); };
  auto put_str = [&](const char* s) { int32_t l = (int32_t)std::strlen(s); put_i32(l); put(s, (size_t)l); };
  put_i32(kMagic);
  put_str("const");
  put_str("standard");
  put_i32(2);  // CONST_FILE_VERSION: packed
  uint32_t flags = 0; put(&flags, 4);
  uint64_t p = c.props | props::kExpanded; put(&p, 8);  // static_properties() = EXPANDED (data_structure.rs:32-36)
  put_i64(c.has_start ? (int64_t)c.start : -1);
  put_i64((int64_t)n);
  put_i64((int64_t)c.arcs.size());
  for (size_t s = 0; s < n; s++) {
    put(&c.finals[s], 4);
    const uint32_t lo = c.offsets[s], hi = c.offsets[s + 1];
    int32_t nie = 0, noe = 0;
    for (uint32_t e = lo; e < hi; e++) { nie += c.arcs[e].ilabel == kEps; noe += c.arcs[e].olabel == kEps; }
    put_i32((int32_t)lo); put_i32((int32_t)(hi - lo)); put_i32(nie); put_i32(noe);
  }
  if (!c.arcs.empty()) put(c.arcs.data(), c.arcs.size() * 16);
  return buf;
}

inline std::vector<uint8_t> read_file(const std::string& path) {
  std::ifstream in(path, std::ios::binary | std::ios::ate);
  if (!in) throw FstError("Error while opening file \"" + path + "\"");
  const std::streamsize size = in.tellg();
  in.seekg(0, std::ios::beg);
  std::vector<uint8_t> buf((size_t)(size > 0 ? size : 0));
  if (size > 0 && !in.read(reinterpret_cast<char*>(buf.data()), size))
    throw FstError("Error while reading file \"" + path + "\"");
  return buf;
}
inline void write_file(const std::string& path, const uint8_t* data, size_t size) {
  std::ofstream out(path, std::ios::binary);
  if (!out) throw FstError("Error while creating file \"" + path + "\"");
  out.write((const char*)data, (std::streamsize)size);
}
inline void write_file(const std::string& path, const std::vector<uint8_t>& b) {
  std::ofstream out(path, std::ios::binary);
  if (!out) throw FstError("Error while creating file \"" + path + "\"");
  out.write((const char*)b.data(), (std::streamsize)b.size());
}
}  // namespace io
}  // namespace b200
