// oracle/oracle.hpp — CPU ORACLE (TEST INFRASTRUCTURE ONLY).
//
// A single-threaded C++17 restatement of the reference's (garvys-org/rustfst @ 8e1391d, v1.3.1)
// `algorithms::compose` and `algorithms::shortest_path` for VectorFst<TropicalWeight>, with the same
// data-structure class as the reference (per-state arc vectors, hash map + id vector state table,
// FIFO BFS materialisation, sequential DFS connect, queue-ordered label-correcting SSSP).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use
// anything in this directory, and only as the checker / the CPU baseline.  The product
// (rustfst_b200/csrc) never includes, links or calls it.
//
// Parity pinning: the Rust reference cannot be built in this image (no cargo/rustc) and the
// OpenFst-generated goldens are git-ignored upstream, so the oracle is pinned against the in-tree
// known-answer tests (rustfst-python/tests/algorithms/test_compose.py:13-154,
// test_shortest_path.py:5-51, doc-test compose_static.rs:313-321) — see tests/test_oracle_kat.py —
// and is otherwise a line-by-line restatement.  Every function cites the reference file:line.
// The nshortest > 1 route (shortest_distance, reverse, n_shortest_path) has no in-tree known answer; it is pinned
// by the criterion of the reference's own test (tests_openfst/algorithms/shortest_path.rs:62-92: number of paths,
// weights position by position, paths exist in the input) against a brute-force enumeration of all paths of small
// machines — tests/test_oracle_nshortest.py.  unique = true (determinize_with_distance of the reversed machine first)
// is restated with ONE stated difference: the reference rebuilds every weighted subset from HashMap::values() of a
// RandomState map (determinize_fsa_op.rs:154-165), so the element order of a subset — hence subset identity, state
// numbering and even the number of states of its own determinized machine — changes from process to process; here
// subsets are kept sorted by state.  The n paths and their weights do not depend on that order; they are pinned by the
// same brute-force criterion over DISTINCT label sequences.
//
// All paths below are relative to /root/reference/.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <deque>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <functional>
#include <cmath>
#include <algorithm>
#include <fstream>

namespace oracle {

// rustfst/src/lib.rs:236,269,292,298 (feature state-label-u32: rustfst/Cargo.toml:19-21)
using Label = uint32_t;
using StateId = uint32_t;
constexpr Label EPS_LABEL = 0;
constexpr Label NO_LABEL = 0xFFFFFFFFu;
constexpr StateId NO_STATE_ID = 0xFFFFFFFFu;
constexpr float KDELTA = 1.0f / 1024.0f;
constexpr float W_ZERO = std::numeric_limits<float>::infinity();
constexpr float W_ONE = 0.0f;

// ---------------------------------------------------------------------------------------------
// TropicalWeight — rustfst/src/semirings/tropical_weight.rs:53-70, semirings/semiring.rs:159-168
// ---------------------------------------------------------------------------------------------
inline bool w_eq(float w1, float w2) {  // approx ==, semiring.rs:159-168
  return w1 <= (w2 + KDELTA) && w2 <= (w1 + KDELTA);  // f32 arithmetic (SSE; never build with -ffast-math)
}
inline bool w_is_zero(float w) { return w_eq(w, W_ZERO); }  // semiring.rs:67-69
inline bool w_is_one(float w) { return w_eq(w, W_ONE); }    // semiring.rs:70-72
inline float w_plus(float a, float b) { return (b < a) ? b : a; }  // tropical_weight.rs:53-58
inline float w_times(float a, float b) {                           // tropical_weight.rs:60-70
  if (a == W_ZERO) return a;
  if (b == W_ZERO) return b;
  return a + b;
}

// ---------------------------------------------------------------------------------------------
// FstProperties — rustfst/src/fst_properties/properties.rs:21-103 (bit layout) and masks
// ---------------------------------------------------------------------------------------------
namespace P {
constexpr uint64_t EXPANDED = 1, MUTABLE = 2, ERROR = 4;
constexpr uint64_t ACCEPTOR = 0x0000000000010000ULL, NOT_ACCEPTOR = 0x0000000000020000ULL;
constexpr uint64_t I_DETERMINISTIC = 0x0000000000040000ULL, NOT_I_DETERMINISTIC = 0x0000000000080000ULL;
constexpr uint64_t O_DETERMINISTIC = 0x0000000000100000ULL, NOT_O_DETERMINISTIC = 0x0000000000200000ULL;
constexpr uint64_t EPSILONS = 0x0000000000400000ULL, NO_EPSILONS = 0x0000000000800000ULL;
constexpr uint64_t I_EPSILONS = 0x0000000001000000ULL, NO_I_EPSILONS = 0x0000000002000000ULL;
constexpr uint64_t O_EPSILONS = 0x0000000004000000ULL, NO_O_EPSILONS = 0x0000000008000000ULL;
constexpr uint64_t I_LABEL_SORTED = 0x0000000010000000ULL, NOT_I_LABEL_SORTED = 0x0000000020000000ULL;
constexpr uint64_t O_LABEL_SORTED = 0x0000000040000000ULL, NOT_O_LABEL_SORTED = 0x0000000080000000ULL;
constexpr uint64_t WEIGHTED = 0x0000000100000000ULL, UNWEIGHTED = 0x0000000200000000ULL;
constexpr uint64_t CYCLIC = 0x0000000400000000ULL, ACYCLIC = 0x0000000800000000ULL;
constexpr uint64_t INITIAL_CYCLIC = 0x0000001000000000ULL, INITIAL_ACYCLIC = 0x0000002000000000ULL;
constexpr uint64_t TOP_SORTED = 0x0000004000000000ULL, NOT_TOP_SORTED = 0x0000008000000000ULL;
constexpr uint64_t ACCESSIBLE = 0x0000010000000000ULL, NOT_ACCESSIBLE = 0x0000020000000000ULL;
constexpr uint64_t COACCESSIBLE = 0x0000040000000000ULL, NOT_COACCESSIBLE = 0x0000080000000000ULL;
constexpr uint64_t STRING = 0x0000100000000000ULL, NOT_STRING = 0x0000200000000000ULL;
constexpr uint64_t WEIGHTED_CYCLES = 0x0000400000000000ULL, UNWEIGHTED_CYCLES = 0x0000800000000000ULL;

constexpr uint64_t BINARY = 0x7ULL;                     // properties.rs binary_properties
constexpr uint64_t TRINARY = 0x0000ffffffff0000ULL;     // trinary_properties
constexpr uint64_t POS_TRINARY = TRINARY & 0x5555555555555555ULL;
constexpr uint64_t NEG_TRINARY = TRINARY & 0xaaaaaaaaaaaaaaaaULL;
constexpr uint64_t ALL = BINARY | TRINARY;

// properties.rs null_properties (properties of an empty machine)
constexpr uint64_t NULL_PROPS = ACCEPTOR | I_DETERMINISTIC | O_DETERMINISTIC | NO_EPSILONS | NO_I_EPSILONS |
                                NO_O_EPSILONS | I_LABEL_SORTED | O_LABEL_SORTED | UNWEIGHTED | ACYCLIC |
                                INITIAL_ACYCLIC | TOP_SORTED | ACCESSIBLE | COACCESSIBLE | STRING |
                                UNWEIGHTED_CYCLES;
// properties.rs set_start_properties()
constexpr uint64_t SET_START = ACCEPTOR | NOT_ACCEPTOR | I_DETERMINISTIC | NOT_I_DETERMINISTIC | O_DETERMINISTIC |
                               NOT_O_DETERMINISTIC | EPSILONS | NO_EPSILONS | I_EPSILONS | NO_I_EPSILONS |
                               O_EPSILONS | NO_O_EPSILONS | I_LABEL_SORTED | NOT_I_LABEL_SORTED | O_LABEL_SORTED |
                               NOT_O_LABEL_SORTED | WEIGHTED | UNWEIGHTED | CYCLIC | ACYCLIC | TOP_SORTED |
                               NOT_TOP_SORTED | COACCESSIBLE | NOT_COACCESSIBLE | WEIGHTED_CYCLES |
                               UNWEIGHTED_CYCLES;
// properties.rs set_final_properties()
constexpr uint64_t SET_FINAL = ACCEPTOR | NOT_ACCEPTOR | I_DETERMINISTIC | NOT_I_DETERMINISTIC | O_DETERMINISTIC |
                               NOT_O_DETERMINISTIC | EPSILONS | NO_EPSILONS | I_EPSILONS | NO_I_EPSILONS |
                               O_EPSILONS | NO_O_EPSILONS | I_LABEL_SORTED | NOT_I_LABEL_SORTED | O_LABEL_SORTED |
                               NOT_O_LABEL_SORTED | CYCLIC | ACYCLIC | INITIAL_CYCLIC | INITIAL_ACYCLIC |
                               TOP_SORTED | NOT_TOP_SORTED | ACCESSIBLE | NOT_ACCESSIBLE | WEIGHTED_CYCLES |
                               UNWEIGHTED_CYCLES;
// properties.rs add_state_properties()
constexpr uint64_t ADD_STATE = ACCEPTOR | NOT_ACCEPTOR | I_DETERMINISTIC | NOT_I_DETERMINISTIC | O_DETERMINISTIC |
                               NOT_O_DETERMINISTIC | EPSILONS | NO_EPSILONS | I_EPSILONS | NO_I_EPSILONS |
                               O_EPSILONS | NO_O_EPSILONS | I_LABEL_SORTED | NOT_I_LABEL_SORTED | O_LABEL_SORTED |
                               NOT_O_LABEL_SORTED | WEIGHTED | UNWEIGHTED | CYCLIC | ACYCLIC | INITIAL_CYCLIC |
                               INITIAL_ACYCLIC | TOP_SORTED | NOT_TOP_SORTED | NOT_ACCESSIBLE | NOT_COACCESSIBLE |
                               NOT_STRING | WEIGHTED_CYCLES | UNWEIGHTED_CYCLES;
// properties.rs add_arc_properties()
constexpr uint64_t ADD_ARC = NOT_ACCEPTOR | NOT_I_DETERMINISTIC | NOT_O_DETERMINISTIC | EPSILONS | I_EPSILONS |
                             O_EPSILONS | NOT_I_LABEL_SORTED | NOT_O_LABEL_SORTED | WEIGHTED | CYCLIC |
                             INITIAL_CYCLIC | NOT_TOP_SORTED | ACCESSIBLE | COACCESSIBLE | WEIGHTED_CYCLES;
// properties.rs delete_states_properties()
constexpr uint64_t DELETE_STATES = ACCEPTOR | I_DETERMINISTIC | O_DETERMINISTIC | NO_EPSILONS | NO_I_EPSILONS |
                                   NO_O_EPSILONS | I_LABEL_SORTED | O_LABEL_SORTED | UNWEIGHTED | ACYCLIC |
                                   INITIAL_ACYCLIC | TOP_SORTED | UNWEIGHTED_CYCLES;
// properties.rs arcsort_properties()
constexpr uint64_t ARCSORT = ACCEPTOR | NOT_ACCEPTOR | I_DETERMINISTIC | NOT_I_DETERMINISTIC | O_DETERMINISTIC |
                             NOT_O_DETERMINISTIC | EPSILONS | NO_EPSILONS | I_EPSILONS | NO_I_EPSILONS | O_EPSILONS |
                             NO_O_EPSILONS | WEIGHTED | UNWEIGHTED | CYCLIC | ACYCLIC | INITIAL_CYCLIC |
                             INITIAL_ACYCLIC | TOP_SORTED | NOT_TOP_SORTED | ACCESSIBLE | NOT_ACCESSIBLE |
                             COACCESSIBLE | NOT_COACCESSIBLE | STRING | NOT_STRING | WEIGHTED_CYCLES |
                             UNWEIGHTED_CYCLES;

// fst_properties/utils.rs:4-9
inline uint64_t known_properties(uint64_t props) {
  return BINARY | (props & TRINARY) | ((props & POS_TRINARY) << 1) | ((props & NEG_TRINARY) >> 1);
}
}  // namespace P

// rustfst/src/tr.rs:6-15
struct Tr {
  Label ilabel;
  Label olabel;
  float weight;
  StateId nextstate;
};

// rustfst/src/fst_impls/vector_fst/data_structure.rs:29-34
struct State {
  bool has_final = false;
  float final_weight = W_ZERO;
  std::vector<Tr> trs;
  size_t niepsilons = 0;
  size_t noepsilons = 0;
};

// fst_properties/mutate_properties.rs:43-100
inline uint64_t add_tr_properties(uint64_t inprops, StateId state, const Tr& tr, const Tr* prev_tr) {
  uint64_t out = inprops;
  if (tr.ilabel != tr.olabel) { out |= P::NOT_ACCEPTOR; out &= ~P::ACCEPTOR; }
  if (tr.ilabel == EPS_LABEL) {
    out |= P::I_EPSILONS; out &= ~P::NO_I_EPSILONS;
    if (tr.olabel == EPS_LABEL) { out |= P::EPSILONS; out &= ~P::NO_EPSILONS; }
  }
  if (tr.olabel == EPS_LABEL) { out |= P::O_EPSILONS; out &= ~P::NO_O_EPSILONS; }
  if (prev_tr) {
    if (prev_tr->ilabel > tr.ilabel) { out |= P::NOT_I_LABEL_SORTED; out &= ~P::I_LABEL_SORTED; }
    if (prev_tr->olabel > tr.olabel) { out |= P::NOT_O_LABEL_SORTED; out &= ~P::O_LABEL_SORTED; }
  }
  if (!w_is_zero(tr.weight) && !w_is_one(tr.weight)) { out |= P::WEIGHTED; out &= ~P::UNWEIGHTED; }
  if (tr.nextstate <= state) { out |= P::NOT_TOP_SORTED; out &= ~P::TOP_SORTED; }
  out &= P::ADD_ARC | P::ACCEPTOR | P::NO_EPSILONS | P::NO_I_EPSILONS | P::NO_O_EPSILONS | P::I_LABEL_SORTED |
         P::O_LABEL_SORTED | P::UNWEIGHTED | P::TOP_SORTED;
  if (out & P::TOP_SORTED) out |= P::ACYCLIC | P::INITIAL_ACYCLIC;
  return out;
}

// ---------------------------------------------------------------------------------------------
// VectorFst<TropicalWeight> — fst_impls/vector_fst/{data_structure,fst,mutable_fst}.rs
// ---------------------------------------------------------------------------------------------
struct Fst {
  std::vector<State> states;
  bool has_start = false;
  StateId start = 0;
  uint64_t props = P::NULL_PROPS;  // mutable_fst.rs:25-33

  size_t num_states() const { return states.size(); }
  size_t num_trs_total() const { size_t n = 0; for (auto& s : states) n += s.trs.size(); return n; }

  StateId add_state() {  // mutable_fst.rs:79-84
    states.emplace_back();
    props &= P::ADD_STATE;
    return (StateId)(states.size() - 1);
  }
  void add_states(size_t n) {  // mutable_fst.rs:86-90
    states.resize(states.size() + n);
    props &= P::ADD_STATE;
  }
  void set_start(StateId s) {  // mutable_fst.rs:35-44 + mutate_properties.rs:7-13
    if (s >= states.size()) throw std::runtime_error("The state doesn't exist");
    has_start = true; start = s;
    uint64_t out = props & P::SET_START;
    if (props & P::ACYCLIC) out |= P::INITIAL_ACYCLIC;
    props = out;
  }
  void set_final(StateId s, float w) {  // mutable_fst.rs:51-64 + mutate_properties.rs:15-37
    if (s >= states.size()) throw std::runtime_error("Stateid doesn't exist");
    State& st = states[s];
    uint64_t out = props;
    if (st.has_final && !w_is_zero(st.final_weight) && !w_is_one(st.final_weight)) out &= ~P::WEIGHTED;
    if (!w_is_zero(w) && !w_is_one(w)) { out |= P::WEIGHTED; out &= ~P::UNWEIGHTED; }
    out &= P::SET_FINAL | P::WEIGHTED | P::UNWEIGHTED;
    props = out;
    st.has_final = true; st.final_weight = w;
  }
  void add_tr(StateId s, const Tr& tr) {  // mutable_fst.rs:236-245 + data_structure.rs:80-91
    if (s >= states.size()) throw std::runtime_error("State doesn't exist");
    State& st = states[s];
    if (tr.ilabel == EPS_LABEL) st.niepsilons++;
    if (tr.olabel == EPS_LABEL) st.noepsilons++;
    st.trs.push_back(tr);
    const Tr* prev = st.trs.size() > 1 ? &st.trs[st.trs.size() - 2] : nullptr;
    props = add_tr_properties(props, s, st.trs.back(), prev);
  }
  // mutable_fst.rs:255-281 (properties are overwritten by the caller afterwards)
  void set_trs_unchecked(StateId s, std::vector<Tr>&& trs) {
    State& st = states[s];
    st.trs = std::move(trs);
    uint64_t pr = props;
    size_t ni = 0, no = 0;
    for (size_t i = 0; i < st.trs.size(); i++) {
      pr = add_tr_properties(pr, s, st.trs[i], i >= 1 ? &st.trs[i - 1] : nullptr);
      if (st.trs[i].ilabel == EPS_LABEL) ni++;
      if (st.trs[i].olabel == EPS_LABEL) no++;
    }
    st.niepsilons = ni; st.noepsilons = no;
    props = pr;
  }
  // mutable_fst.rs:132-189
  void del_states(const std::vector<StateId>& dstates) {
    std::vector<int64_t> new_id(states.size(), 0);
    for (StateId s : dstates) new_id[s] = -1;
    size_t nstates = 0;
    for (size_t s = 0; s < states.size(); s++) {
      if (new_id[s] != -1) {
        new_id[s] = (int64_t)nstates;
        if (s != nstates) std::swap(states[nstates], states[s]);
        nstates++;
      }
    }
    states.resize(nstates);
    for (size_t s = 0; s < states.size(); s++) {
      State& st = states[s];
      std::vector<Tr> kept;
      kept.reserve(st.trs.size());
      for (Tr& tr : st.trs) {
        int64_t t = new_id[tr.nextstate];
        if (t != -1) { tr.nextstate = (StateId)t; kept.push_back(tr); }
        else {
          if (tr.ilabel == EPS_LABEL) st.niepsilons--;
          if (tr.olabel == EPS_LABEL) st.noepsilons--;
        }
      }
      st.trs.swap(kept);
    }
    if (has_start) {
      int64_t ns = new_id[start];
      if (ns == -1) has_start = false; else start = (StateId)ns;
    }
    props &= P::DELETE_STATES;
  }
  void set_properties_with_mask(uint64_t p, uint64_t mask) {  // mutable_fst.rs:411-414
    props &= ~mask;
    props |= p & mask;
  }
};

// VectorFst::eq — data_structure.rs:36-41 (+ derive(PartialEq) on state; approx weights)
inline bool fst_equal(const Fst& a, const Fst& b) {
  if (a.has_start != b.has_start) return false;
  if (a.has_start && a.start != b.start) return false;
  if (a.states.size() != b.states.size()) return false;
  for (size_t s = 0; s < a.states.size(); s++) {
    const State &x = a.states[s], &y = b.states[s];
    if (x.has_final != y.has_final) return false;
    if (x.has_final && !w_eq(x.final_weight, y.final_weight)) return false;
    if (x.trs.size() != y.trs.size()) return false;
    if (x.niepsilons != y.niepsilons || x.noepsilons != y.noepsilons) return false;
    for (size_t i = 0; i < x.trs.size(); i++) {
      const Tr &p = x.trs[i], &q = y.trs[i];
      if (p.ilabel != q.ilabel || p.olabel != q.olabel || p.nextstate != q.nextstate || !w_eq(p.weight, q.weight))
        return false;
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// DFS — algorithms/dfs_visit.rs:97-187 (iterative, OpenFst-style; roots: start, then 0,1,2,...)
// ---------------------------------------------------------------------------------------------
struct Visitor {
  virtual ~Visitor() = default;
  virtual bool init_state(StateId s, StateId root) = 0;
  virtual bool tree_tr(StateId s, const Tr& tr) = 0;
  virtual bool back_tr(StateId s, const Tr& tr) = 0;
  virtual bool forward_or_cross_tr(StateId s, const Tr& tr) = 0;
  virtual void finish_state(StateId s, bool has_parent, StateId parent) = 0;
  virtual void finish_visit() = 0;
};

inline void dfs_visit(const Fst& fst, Visitor& v, bool access_only) {
  if (!fst.has_start) { v.finish_visit(); return; }
  const StateId start = fst.start;
  const size_t nstates = fst.num_states();
  enum : uint8_t { White = 0, Grey = 1, Black = 2 };
  std::vector<uint8_t> color(nstates, White);
  struct Frame { StateId s; size_t pos; };
  std::vector<Frame> stack;
  bool dfs = true;
  size_t root = start;
  while (true) {
    if (!dfs || root >= nstates) break;
    color[root] = Grey;
    stack.push_back({(StateId)root, 0});
    dfs = v.init_state((StateId)root, (StateId)root);
    while (!stack.empty()) {
      Frame& fr = stack.back();
      StateId s = fr.s;
      const auto& trs = fst.states[s].trs;
      if (!dfs || fr.pos >= trs.size()) {
        color[s] = Black;
        stack.pop_back();
        if (!stack.empty()) {
          v.finish_state(s, true, stack.back().s);
          stack.back().pos++;
        } else {
          v.finish_state(s, false, 0);
        }
        continue;
      }
      const Tr& tr = trs[fr.pos];
      uint8_t nc = color[tr.nextstate];
      if (nc == White) {
        dfs = v.tree_tr(s, tr);
        if (!dfs) break;  // dfs_visit.rs:152-155 (breaks the inner loop, leaving the stack)
        color[tr.nextstate] = Grey;
        StateId ns = tr.nextstate;
        stack.push_back({ns, 0});
        dfs = v.init_state(ns, (StateId)root);
      } else if (nc == Grey) {
        dfs = v.back_tr(s, tr);
        stack.back().pos++;
      } else {
        dfs = v.forward_or_cross_tr(s, tr);
        stack.back().pos++;
      }
    }
    if (access_only) break;
    root = (root == start) ? 0 : root + 1;
    while (root < nstates && color[root] != White) root++;
  }
  v.finish_visit();
}

// algorithms/visitors/scc_visitors.rs:10-180 (also the body of connect.rs ConnectVisitor)
struct SccVisitor : Visitor {
  const Fst& fst;
  std::vector<int32_t> scc;
  std::vector<uint8_t> access, coaccess;
  StateId start;
  size_t nstates = 0;
  std::vector<int32_t> dfnumber, lowlink;
  std::vector<uint8_t> onstack;
  std::vector<StateId> scc_stack;
  int32_t nscc = 0;
  uint64_t props;
  explicit SccVisitor(const Fst& f)
      : fst(f), scc(f.num_states(), -1), access(f.num_states(), 0), coaccess(f.num_states(), 0),
        start(f.has_start ? f.start : NO_STATE_ID), dfnumber(f.num_states(), -1), lowlink(f.num_states(), -1),
        onstack(f.num_states(), 0) {
    props = P::ACYCLIC | P::INITIAL_ACYCLIC | P::ACCESSIBLE | P::COACCESSIBLE;
  }
  bool init_state(StateId s, StateId root) override {
    scc_stack.push_back(s);
    dfnumber[s] = (int32_t)nstates; lowlink[s] = (int32_t)nstates; onstack[s] = 1;
    // connect.rs:112 — access[s] = (root == start)
    access[s] = (root == start);
    if (root != start) { props |= P::NOT_ACCESSIBLE; props &= ~P::ACCESSIBLE; }
    nstates++;
    return true;
  }
  bool tree_tr(StateId, const Tr&) override { return true; }
  bool back_tr(StateId s, const Tr& tr) override {
    StateId t = tr.nextstate;
    if (dfnumber[t] < lowlink[s]) lowlink[s] = dfnumber[t];
    if (coaccess[t]) coaccess[s] = 1;
    props |= P::CYCLIC; props &= ~P::ACYCLIC;
    if (t == start) { props |= P::INITIAL_CYCLIC; props &= ~P::INITIAL_ACYCLIC; }
    return true;
  }
  bool forward_or_cross_tr(StateId s, const Tr& tr) override {
    StateId t = tr.nextstate;
    if (dfnumber[t] < dfnumber[s] && onstack[t] && dfnumber[t] < lowlink[s]) lowlink[s] = dfnumber[t];
    if (coaccess[t]) coaccess[s] = 1;
    return true;
  }
  void finish_state(StateId s, bool has_parent, StateId parent) override {
    if (fst.states[s].has_final) coaccess[s] = 1;
    if (dfnumber[s] == lowlink[s]) {
      bool scc_coaccess = false;
      size_t i = scc_stack.size();
      StateId t;
      do { i--; t = scc_stack[i]; if (coaccess[t]) scc_coaccess = true; } while (s != t);
      do {
        t = scc_stack.back();
        scc[t] = nscc;
        if (scc_coaccess) coaccess[t] = 1;
        onstack[t] = 0;
        scc_stack.pop_back();
      } while (s != t);
      if (!scc_coaccess) { props |= P::NOT_COACCESSIBLE; props &= ~P::COACCESSIBLE; }
      nscc++;
    }
    if (has_parent) {
      if (coaccess[s]) coaccess[parent] = 1;
      if (lowlink[s] < lowlink[parent]) lowlink[parent] = lowlink[s];
    }
  }
  void finish_visit() override {  // scc_visitors.rs:172-179
    for (auto& c : scc) c = nscc - 1 - c;
  }
};

// algorithms/connect.rs:51-66
inline void connect(Fst& fst) {
  SccVisitor v(fst);
  dfs_visit(fst, v, false);
  std::vector<StateId> dstates;
  for (size_t s = 0; s < fst.num_states(); s++)
    if (!v.access[s] || !v.coaccess[s]) dstates.push_back((StateId)s);
  fst.del_states(dstates);
  fst.set_properties_with_mask(P::ACCESSIBLE | P::COACCESSIBLE, P::ACCESSIBLE | P::COACCESSIBLE);
}

// fst_properties/compute_fst_properties.rs:14-208 with mask = all_properties, use_stored = false.
// (Used when materialising fixture inputs: rustfst-tests-data/main.cpp:1048,1193 force all bits.)
inline uint64_t compute_fst_properties_all(const Fst& fst) {
  uint64_t comp = 0;  // fst_props & binary_properties (binary bits are never stored)
  const uint64_t dfs_props = P::ACYCLIC | P::CYCLIC | P::INITIAL_ACYCLIC | P::INITIAL_CYCLIC | P::ACCESSIBLE |
                             P::NOT_ACCESSIBLE | P::COACCESSIBLE | P::NOT_COACCESSIBLE;
  SccVisitor v(fst);
  // SccVisitor::new(fst, true, true): with compute_access the visitor marks access for every root
  // (scc_visitors.rs:66-76); only props are consumed here.
  dfs_visit(fst, v, false);
  comp |= dfs_props & v.props;
  const std::vector<int32_t>& sccs = v.scc;
  comp |= P::ACCEPTOR | P::NO_EPSILONS | P::NO_I_EPSILONS | P::NO_O_EPSILONS | P::I_LABEL_SORTED | P::O_LABEL_SORTED |
          P::UNWEIGHTED | P::TOP_SORTED | P::STRING;
  comp |= P::I_DETERMINISTIC | P::O_DETERMINISTIC | P::UNWEIGHTED_CYCLES;
  size_t nfinal = 0;
  for (size_t state = 0; state < fst.num_states(); state++) {
    std::unordered_set<Label> il, ol;
    const Tr* prev = nullptr;
    for (const Tr& tr : fst.states[state].trs) {
      if (il.count(tr.ilabel)) { comp |= P::NOT_I_DETERMINISTIC; comp &= ~P::I_DETERMINISTIC; }
      if (ol.count(tr.olabel)) { comp |= P::NOT_O_DETERMINISTIC; comp &= ~P::O_DETERMINISTIC; }
      if (tr.ilabel != tr.olabel) { comp |= P::NOT_ACCEPTOR; comp &= ~P::ACCEPTOR; }
      if (tr.ilabel == 0 && tr.olabel == 0) { comp |= P::EPSILONS; comp &= ~P::NO_EPSILONS; }
      if (tr.ilabel == 0) { comp |= P::I_EPSILONS; comp &= ~P::NO_I_EPSILONS; }
      if (tr.olabel == 0) { comp |= P::O_EPSILONS; comp &= ~P::NO_O_EPSILONS; }
      if (prev) {
        if (tr.ilabel < prev->ilabel) { comp |= P::NOT_I_LABEL_SORTED; comp &= ~P::I_LABEL_SORTED; }
        if (tr.olabel < prev->olabel) { comp |= P::NOT_O_LABEL_SORTED; comp &= ~P::O_LABEL_SORTED; }
      }
      if (!w_is_one(tr.weight) && !w_is_zero(tr.weight)) {
        comp |= P::WEIGHTED; comp &= ~P::UNWEIGHTED;
        if ((comp & P::UNWEIGHTED_CYCLES) && sccs[state] == sccs[tr.nextstate]) {
          comp |= P::WEIGHTED_CYCLES; comp &= ~P::UNWEIGHTED_CYCLES;
        }
      }
      if (tr.nextstate <= state) { comp |= P::NOT_TOP_SORTED; comp &= ~P::TOP_SORTED; }
      if (tr.nextstate != state + 1) { comp |= P::NOT_STRING; comp &= ~P::STRING; }
      prev = &tr;
      il.insert(tr.ilabel); ol.insert(tr.olabel);
    }
    if (nfinal > 0) { comp |= P::NOT_STRING; comp &= ~P::STRING; }
    if (fst.states[state].has_final) {
      if (!w_is_one(fst.states[state].final_weight)) { comp |= P::WEIGHTED; comp &= ~P::UNWEIGHTED; }
      nfinal++;
    } else if (fst.states[state].trs.size() != 1) {
      comp |= P::NOT_STRING; comp &= ~P::STRING;
    }
  }
  if (fst.has_start && fst.start != 0) { comp |= P::NOT_STRING; comp &= ~P::STRING; }
  return comp;
}

// algorithms/tr_sort.rs:14-62 (stable per-state sort; Vec::sort_by is stable)
inline void tr_sort(Fst& fst, bool ilabel_comp) {
  uint64_t props = fst.props;
  for (auto& st : fst.states) {
    if (ilabel_comp)
      std::stable_sort(st.trs.begin(), st.trs.end(), [](const Tr& a, const Tr& b) { return a.ilabel < b.ilabel; });
    else
      std::stable_sort(st.trs.begin(), st.trs.end(), [](const Tr& a, const Tr& b) { return a.olabel < b.olabel; });
  }
  uint64_t out = (props & P::ARCSORT) | (ilabel_comp ? P::I_LABEL_SORTED : P::O_LABEL_SORTED);
  if (props & P::ACCEPTOR) out |= (ilabel_comp ? P::O_LABEL_SORTED : P::I_LABEL_SORTED);
  fst.set_properties_with_mask(out, P::ALL);
}

// ---------------------------------------------------------------------------------------------
// compose — algorithms/compose/*
// ---------------------------------------------------------------------------------------------
// rustfst-ffi/src/algorithms/compose.rs:20-33 enum values
enum ComposeFilter : int { AUTO = 0, NULLF = 1, TRIVIAL = 2, SEQUENCE = 3, ALT_SEQUENCE = 4, MATCH = 5, NO_MATCH = 6 };

// fst_properties/mutate_properties.rs:151-184
inline uint64_t compose_properties(uint64_t p1, uint64_t p2) {
  uint64_t out = 0;
  if ((p1 & P::ACCEPTOR) && (p2 & P::ACCEPTOR)) {
    out |= P::ACCEPTOR | P::ACCESSIBLE;
    out |= (P::NO_EPSILONS | P::NO_I_EPSILONS | P::NO_O_EPSILONS | P::ACYCLIC | P::INITIAL_ACYCLIC) & p1 & p2;
    if ((p1 & P::NO_I_EPSILONS) && (p2 & P::NO_I_EPSILONS)) out |= (P::I_DETERMINISTIC | P::O_DETERMINISTIC) & p1 & p2;
  } else {
    out |= P::ACCESSIBLE;
    out |= (P::ACCEPTOR | P::NO_I_EPSILONS | P::ACYCLIC | P::INITIAL_ACYCLIC) & p1 & p2;
    if ((p1 & P::NO_I_EPSILONS) && (p2 & P::NO_I_EPSILONS)) out |= P::I_DETERMINISTIC & p1 & p2;
  }
  return out;
}

enum MatchType { MatchInput, MatchOutput, MatchBoth, MatchNone, MatchUnknown };

// matchers/sorted_matcher.rs:56-85
inline MatchType sorted_match_type(const Fst& fst, MatchType mt, bool test) {
  uint64_t true_prop = (mt == MatchInput) ? P::I_LABEL_SORTED : P::O_LABEL_SORTED;
  uint64_t false_prop = (mt == MatchInput) ? P::NOT_I_LABEL_SORTED : P::NOT_O_LABEL_SORTED;
  uint64_t props = fst.props;
  if (test) {  // fst_traits/fst.rs:166-176 properties_check
    uint64_t known = P::known_properties(props);
    if ((known & (true_prop | false_prop)) != (true_prop | false_prop))
      throw std::runtime_error("Properties are not known");
  }
  if (props & true_prop) return mt;
  if (props & false_prop) return MatchNone;
  return MatchUnknown;
}

// The fs part of ComposeStateTuple: IntegerFilterState(u32) or TrivialFilterState(bool) mapped to
// u32 {0,1,2} resp. {0=false,1=true}; NO_STATE_ID / false mean "no state".
struct FilterCtx {
  int kind;  // ComposeFilter (AUTO already mapped to SEQUENCE)
  uint32_t fs;
  bool alleps1, noeps1, alleps2, noeps2;
};
constexpr uint32_t FS_NO_STATE = 0xFFFFFFFFu;

inline bool filter_is_trivial_state(int kind) { return kind == NULLF || kind == TRIVIAL || kind == NO_MATCH; }
inline uint32_t filter_start(int kind) { return filter_is_trivial_state(kind) ? 1u : 0u; }

// compose_filters/*::filter_tr (sequence:150-171, alt_sequence:156-177, match:163-206, null:124-131,
// trivial:122-124, no_match:124-128). arc1 is always the fst1-side arc, arc2 the fst2-side arc.
inline uint32_t filter_tr(const FilterCtx& c, const Tr& arc1, const Tr& arc2) {
  switch (c.kind) {
    case SEQUENCE:
      if (arc1.olabel == NO_LABEL) return c.alleps1 ? FS_NO_STATE : (c.noeps1 ? 0u : 1u);
      if (arc2.ilabel == NO_LABEL) return c.fs != 0 ? FS_NO_STATE : 0u;
      if (arc1.olabel == EPS_LABEL) return FS_NO_STATE;
      return 0u;
    case ALT_SEQUENCE:
      if (arc2.ilabel == NO_LABEL) return c.alleps2 ? FS_NO_STATE : (c.noeps2 ? 0u : 1u);
      if (arc1.olabel == NO_LABEL) return c.fs == 1 ? FS_NO_STATE : 0u;
      if (arc1.olabel == EPS_LABEL) return FS_NO_STATE;
      return 0u;
    case MATCH:
      if (arc2.ilabel == NO_LABEL) {
        if (c.fs == 0) return c.noeps2 ? 0u : (c.alleps2 ? FS_NO_STATE : 1u);
        if (c.fs == 1) return 1u;
        return FS_NO_STATE;
      }
      if (arc1.olabel == NO_LABEL) {
        if (c.fs == 0) return c.noeps1 ? 0u : (c.alleps1 ? FS_NO_STATE : 2u);
        if (c.fs == 2) return 2u;
        return FS_NO_STATE;
      }
      if (arc1.olabel == EPS_LABEL) return c.fs == 0 ? 0u : FS_NO_STATE;
      return 0u;
    case NULLF:
      return (arc1.olabel == NO_LABEL || arc2.ilabel == NO_LABEL) ? FS_NO_STATE : 1u;
    case TRIVIAL:
      return 1u;
    case NO_MATCH:
      // TrivialFilterState::new(cond); new_no_state() == new(false)  (trivial_filter_state.rs)
      return (arc1.olabel != EPS_LABEL || arc2.ilabel != EPS_LABEL) ? 1u : FS_NO_STATE;
  }
  throw std::runtime_error("bad filter");
}

struct ComposeStateTuple {  // compose_state_tuple.rs:11-15
  uint32_t fs; StateId s1, s2;
  bool operator==(const ComposeStateTuple& o) const { return fs == o.fs && s1 == o.s1 && s2 == o.s2; }
};
struct TupleHash {
  size_t operator()(const ComposeStateTuple& t) const {
    uint64_t h = (uint64_t)t.s1 * 0x9E3779B97F4A7C15ULL;
    h ^= ((uint64_t)t.s2 + 0x7F4A7C15ULL) * 0xC2B2AE3D27D4EB4FULL + (h << 6) + (h >> 2);
    h ^= (uint64_t)t.fs * 0x165667B19E3779F9ULL;
    return (size_t)(h ^ (h >> 29));
  }
};
// lazy/state_table.rs:20-64,102-125 — tuple -> dense id in insertion order, id -> tuple
struct StateTable {
  std::unordered_map<ComposeStateTuple, StateId, TupleHash> map;
  std::vector<ComposeStateTuple> tuples;
  StateId find_id(const ComposeStateTuple& t) {
    auto it = map.find(t);
    if (it != map.end()) return it->second;
    StateId id = (StateId)tuples.size();
    map.emplace(t, id);
    tuples.push_back(t);
    return id;
  }
};

// matchers/sigma_matcher.rs: SigmaMatcherConfig {sigma_label, rewrite_mode, sigma_allowed_matches}
struct SigmaConfig {
  bool enabled = false;
  Label sigma_label = NO_LABEL;
  int rewrite_mode = 0;  // 0 Auto (rewrite both iff ACCEPTOR), 1 Always, 2 Never  (matchers/mod.rs:69-75)
  bool has_allowed = false;
  std::unordered_set<Label> allowed;
};

struct ComposeConfig {
  int filter = AUTO;
  bool connect = true;
  SigmaConfig sigma1, sigma2;  // matcher1_config / matcher2_config (compose_static.rs:80-97)
};

struct ComposeStats {
  uint64_t states_expanded = 0, arcs_iterated = 0, arcs_emitted = 0;
};

// SortedMatcher lookup of `label` among the arcs of one state on the matched side: [pos, end) run.
// by_olabel selects the field (MatchOutput on fst1) — sorted_matcher.rs:124-184.
inline void sorted_run(const std::vector<Tr>& trs, bool by_olabel, Label ml, size_t* pos, size_t* end) {
  size_t p = std::lower_bound(trs.begin(), trs.end(), ml, [by_olabel](const Tr& x, Label l) {
               return (by_olabel ? x.olabel : x.ilabel) < l; }) - trs.begin();
  size_t e = p;
  while (e < trs.size() && (by_olabel ? trs[e].olabel : trs[e].ilabel) == ml) e++;
  *pos = p; *end = e;
}
// has_sigma (sigma_matcher.rs:33-45): the inner sorted matcher finds at least one arc labelled sigma
inline bool state_has_sigma(const std::vector<Tr>& trs, bool by_olabel, const SigmaConfig& sc) {
  if (!sc.enabled || sc.sigma_label == NO_LABEL) return false;
  if (sc.sigma_label == EPS_LABEL) return true;  // (never constructed: SigmaMatcher::new rejects it)
  size_t p, e;
  sorted_run(trs, by_olabel, sc.sigma_label, &p, &e);
  return e > p;
}

// compose_static.rs:198-298 + compose_fst_op.rs + lazy_fst.rs:226-269
inline Fst compose(const Fst& fst1, const Fst& fst2, const ComposeConfig& cfg, ComposeStats* stats = nullptr) {
  int kind = cfg.filter == AUTO ? SEQUENCE : cfg.filter;  // compose_fst.rs:58-92 (new_auto = Sequence filter)
  if (kind < NULLF || kind > NO_MATCH) throw std::runtime_error("EnumConversionError");

  // compose_static.rs:219-223
  if (cfg.filter == AUTO && (cfg.sigma1.enabled || cfg.sigma2.enabled))
    throw std::runtime_error("Custom MatcherConfig not supported with AutoFilter");
  // SigmaMatcher::new (sigma_matcher.rs:55-84)
  for (const SigmaConfig* sc : {&cfg.sigma1, &cfg.sigma2}) {
    if (!sc->enabled) continue;
    if (sc->rewrite_mode < 0 || sc->rewrite_mode > 2) throw std::runtime_error("EnumConversionError");
    if (sc->sigma_label == EPS_LABEL) throw std::runtime_error("SigmaMatcher: 0 cannot be used as sigma_label");
  }
  const bool rewrite_both1 = cfg.sigma1.rewrite_mode == 1 || (cfg.sigma1.rewrite_mode == 0 && (fst1.props & P::ACCEPTOR));
  const bool rewrite_both2 = cfg.sigma2.rewrite_mode == 1 || (cfg.sigma2.rewrite_mode == 0 && (fst2.props & P::ACCEPTOR));
  // compose_fst_op.rs:170-179: a sigma matcher carries REQUIRE_MATCH (sigma_matcher.rs:126-132)
  if (cfg.sigma1.enabled && cfg.sigma1.sigma_label != NO_LABEL &&
      sorted_match_type(fst1, MatchOutput, true) != MatchOutput)
    throw std::runtime_error("ComposeFst: 1st argument cannot perform required matching (sort?)");
  if (cfg.sigma2.enabled && cfg.sigma2.sigma_label != NO_LABEL &&
      sorted_match_type(fst2, MatchInput, true) != MatchInput)
    throw std::runtime_error("ComposeFst: 2nd argument cannot perform required matching (sort?)");

  // compose_fst_op.rs:169-197 match_type (SortedMatcher flags are empty => REQUIRE_MATCH tests are no-ops)
  MatchType type1 = sorted_match_type(fst1, MatchOutput, false);
  MatchType type2 = sorted_match_type(fst2, MatchInput, false);
  MatchType mt;
  if (type1 == MatchOutput && type2 == MatchInput) mt = MatchBoth;
  else if (type1 == MatchOutput) mt = MatchOutput;
  else if (type2 == MatchInput) mt = MatchInput;
  else if (sorted_match_type(fst1, MatchOutput, true) == MatchOutput) mt = MatchOutput;
  else if (sorted_match_type(fst2, MatchInput, true) == MatchInput) mt = MatchInput;
  else
    throw std::runtime_error(
        "ComposeFst: 1st argument cannot match on output labels and 2nd argument cannot match on input labels "
        "(sort?).");

  const uint64_t cprops = compose_properties(fst1.props, fst2.props);

  Fst out;  // F2::new()
  // compose_fst_op.rs:389-404 compute_start; lazy_fst.rs:226-232
  if (!fst1.has_start || !fst2.has_start) {
    if (cfg.connect) connect(out);
    return out;
  }
  StateTable table;
  StateId start_id = table.find_id({filter_start(kind), fst1.start, fst2.start});
  out.add_states((size_t)start_id + 1);
  out.set_start(start_id);

  // lazy_fst.rs:236-259 — FIFO BFS; ids are handed out by find_id at emission time so the queue order
  // equals id order.
  std::deque<StateId> queue;
  std::vector<uint8_t> visited(start_id + 1, 0);
  visited[start_id] = 1;
  queue.push_back(start_id);
  while (!queue.empty()) {
    StateId s = queue.front();
    queue.pop_front();
    // ---- compute_trs(s): compose_fst_op.rs:406-418
    const ComposeStateTuple tuple = table.tuples[s];
    const StateId s1 = tuple.s1, s2 = tuple.s2;
    const State& st1 = fst1.states[s1];
    const State& st2 = fst2.states[s2];
    FilterCtx fc;
    fc.kind = kind; fc.fs = tuple.fs;
    {  // set_state: sequence_compose_filter.rs:134-148, alt_sequence:139-153, match:132-161
      size_t na1 = st1.trs.size(), ne1 = st1.noepsilons;
      size_t na2 = st2.trs.size(), ne2 = st2.niepsilons;
      fc.alleps1 = (na1 == ne1) && !st1.has_final; fc.noeps1 = (ne1 == 0);
      fc.alleps2 = (na2 == ne2) && !st2.has_final; fc.noeps2 = (ne2 == 0);
    }
    // match_input: compose_fst_op.rs:199-219; SortedMatcher priority = num_trs (sorted_matcher.rs:91-93)
    bool match_input;
    if (mt == MatchInput) match_input = true;
    else if (mt == MatchOutput) match_input = false;
    else {
      // SigmaMatcher::priority (sigma_matcher.rs:134-146): REQUIRE_PRIORITY when the state has a sigma arc
      const bool req1 = state_has_sigma(st1.trs, true, cfg.sigma1), req2 = state_has_sigma(st2.trs, false, cfg.sigma2);
      if (req1 && req2) throw std::runtime_error("Both sides can't require match");
      if (req1) match_input = false;
      else if (req2) match_input = true;
      else match_input = st1.trs.size() <= st2.trs.size();
    }

    std::vector<Tr> trs;
    // ordered_expand: compose_fst_op.rs:221-265; match_tr:324-353; match_tr_selected:287-322
    auto emit = [&](const Tr& arc1, const Tr& arc2) {
      uint32_t fs = filter_tr(fc, arc1, arc2);
      if (fs == FS_NO_STATE) return;
      // add_tr: compose_fst_op.rs:267-285
      ComposeStateTuple nt{fs, arc1.nextstate, arc2.nextstate};
      float w = w_times(arc1.weight, arc2.weight);
      trs.push_back(Tr{arc1.ilabel, arc2.olabel, w, table.find_id(nt)});
    };
    if (match_input) {
      // iterate fst1 at s1 (loop first), search fst2 at s2 by ilabel
      auto match_one = [&](const Tr& a1) {
        Label label = a1.olabel;
        bool current_loop = (label == EPS_LABEL);                 // sorted_matcher.rs:124-155
        Label ml = (label == NO_LABEL) ? EPS_LABEL : label;
        size_t pos = 0;
        if (!current_loop) {
          pos = std::lower_bound(st2.trs.begin(), st2.trs.end(), ml,
                                 [](const Tr& x, Label l) { return x.ilabel < l; }) - st2.trs.begin();
        }
        if (cfg.sigma2.enabled) {  // IteratorSigmaMatcher::new (sigma_matcher.rs:196-246) over matcher2
          const SigmaConfig& sc = cfg.sigma2;
          if (label == sc.sigma_label && sc.sigma_label != NO_LABEL)
            throw std::runtime_error("SigmaMatcher::Find: bad label (sigma)");
          const bool normal_nonempty = current_loop || (pos < st2.trs.size() && st2.trs[pos].ilabel == ml);
          if (!normal_nonempty) {
            if (state_has_sigma(st2.trs, false, sc) && label != EPS_LABEL && label != NO_LABEL &&
                (!sc.has_allowed || sc.allowed.count(label))) {
              size_t sp, se;
              sorted_run(st2.trs, false, sc.sigma_label, &sp, &se);
              for (size_t q = sp; q < se; q++) {  // value_openfst (sigma_matcher.rs:249-276): relabel sigma -> label
                Tr t = st2.trs[q];
                if (rewrite_both2) { if (t.ilabel == sc.sigma_label) t.ilabel = label; if (t.olabel == sc.sigma_label) t.olabel = label; }
                else t.ilabel = label;
                emit(a1, t);
              }
            }
            return;  // next_openfst never switches from normal to sigma matches (r.is_none() is never true there)
          }
        }
        if (current_loop) {  // IterItemMatcher::EpsLoop -> matchers/mod.rs:98-105 (MatchInput)
          Tr loop2{NO_LABEL, EPS_LABEL, W_ONE, s2};
          emit(a1, loop2);
        }
        while (pos < st2.trs.size() && st2.trs[pos].ilabel == ml) { emit(a1, st2.trs[pos]); pos++; }
      };
      Tr loop1{EPS_LABEL, NO_LABEL, W_ONE, s1};  // compose_fst_op.rs:229-231
      match_one(loop1);
      for (const Tr& a1 : st1.trs) match_one(a1);
      if (stats) stats->arcs_iterated += st1.trs.size();
    } else {
      auto match_one = [&](const Tr& a2) {
        Label label = a2.ilabel;
        bool current_loop = (label == EPS_LABEL);
        Label ml = (label == NO_LABEL) ? EPS_LABEL : label;
        size_t pos = 0;
        if (!current_loop) {
          pos = std::lower_bound(st1.trs.begin(), st1.trs.end(), ml,
                                 [](const Tr& x, Label l) { return x.olabel < l; }) - st1.trs.begin();
        }
        if (cfg.sigma1.enabled) {  // sigma matcher on fst1 (MatchOutput)
          const SigmaConfig& sc = cfg.sigma1;
          if (label == sc.sigma_label && sc.sigma_label != NO_LABEL)
            throw std::runtime_error("SigmaMatcher::Find: bad label (sigma)");
          const bool normal_nonempty = current_loop || (pos < st1.trs.size() && st1.trs[pos].olabel == ml);
          if (!normal_nonempty) {
            if (state_has_sigma(st1.trs, true, sc) && label != EPS_LABEL && label != NO_LABEL &&
                (!sc.has_allowed || sc.allowed.count(label))) {
              size_t sp, se;
              sorted_run(st1.trs, true, sc.sigma_label, &sp, &se);
              for (size_t q = sp; q < se; q++) {
                Tr t = st1.trs[q];
                if (rewrite_both1) { if (t.ilabel == sc.sigma_label) t.ilabel = label; if (t.olabel == sc.sigma_label) t.olabel = label; }
                else t.olabel = label;
                emit(t, a2);
              }
            }
            return;
          }
        }
        if (current_loop) {  // eps_loop(MatchOutput)
          Tr loop1{EPS_LABEL, NO_LABEL, W_ONE, s1};
          emit(loop1, a2);
        }
        while (pos < st1.trs.size() && st1.trs[pos].olabel == ml) { emit(st1.trs[pos], a2); pos++; }
      };
      Tr loop2{NO_LABEL, EPS_LABEL, W_ONE, s2};  // compose_fst_op.rs:232-233
      match_one(loop2);
      for (const Tr& a2 : st2.trs) match_one(a2);
      if (stats) stats->arcs_iterated += st2.trs.size();
    }
    if (stats) { stats->states_expanded++; stats->arcs_emitted += trs.size(); }

    // ---- lazy_fst.rs:243-259
    for (const Tr& tr : trs) {
      if (tr.nextstate >= visited.size()) visited.resize((size_t)tr.nextstate + 1, 0);
      if (!visited[tr.nextstate]) { queue.push_back(tr.nextstate); visited[tr.nextstate] = 1; }
      size_t n = out.num_states();
      if (tr.nextstate >= n) out.add_states((size_t)tr.nextstate - n + 1);
    }
    out.set_trs_unchecked(s, std::move(trs));
    // compute_final_weight: compose_fst_op.rs:420-449
    if (st1.has_final && st2.has_final) {
      float fw = w_times(st1.final_weight, st2.final_weight);
      if (!w_is_zero(fw)) out.set_final(s, fw);
    }
  }
  out.props = cprops;  // lazy_fst.rs:260 set_properties(self.properties())
  if (cfg.connect) connect(out);  // compose_static.rs:293-295
  return out;
}

// ---------------------------------------------------------------------------------------------
// shortest_path — algorithms/shortest_path.rs, algorithms/queues/*
// ---------------------------------------------------------------------------------------------
struct Queue {  // algorithms/queue.rs
  virtual ~Queue() = default;
  virtual void enqueue(StateId s) = 0;
  virtual bool dequeue(StateId* out) = 0;
  virtual void update(StateId) {}
  virtual bool is_empty() const = 0;
  virtual void clear() = 0;
};
struct FifoQueue : Queue {  // queues/fifo_queue.rs
  std::deque<StateId> q;
  void enqueue(StateId s) override { q.push_back(s); }
  bool dequeue(StateId* o) override { if (q.empty()) return false; *o = q.front(); q.pop_front(); return true; }
  bool is_empty() const override { return q.empty(); }
  void clear() override { q.clear(); }
};
struct LifoQueue : Queue {  // queues/lifo_queue.rs
  std::vector<StateId> q;
  void enqueue(StateId s) override { q.push_back(s); }
  bool dequeue(StateId* o) override { if (q.empty()) return false; *o = q.back(); q.pop_back(); return true; }
  bool is_empty() const override { return q.empty(); }
  void clear() override { q.clear(); }
};
struct TrivialQueue : Queue {  // queues/trivial_queue.rs
  bool has = false; StateId st = 0;
  void enqueue(StateId s) override { has = true; st = s; }
  bool dequeue(StateId* o) override { if (!has) return false; *o = st; has = false; return true; }
  bool is_empty() const override { return !has; }
  void clear() override { has = false; }
};
struct StateOrderQueue : Queue {  // queues/state_order_queue.rs
  size_t front = 0; bool has_back = false; size_t back = 0;
  std::vector<uint8_t> enq;
  void enqueue(StateId s) override {
    size_t state = s;
    if (!has_back || front > back) { front = state; back = state; has_back = true; }
    else if (state > back) back = state;
    else if (state < front) front = state;
    while (enq.size() <= state) enq.push_back(0);
    enq[state] = 1;
  }
  bool is_empty() const override { return has_back ? front > back : true; }
  bool dequeue(StateId* o) override {
    if (is_empty()) return false;
    *o = (StateId)front;
    enq[front] = 0;
    while (front <= back && !enq[front]) front++;
    return true;
  }
  void clear() override {
    if (has_back) for (size_t i = front; i <= back && i < enq.size(); i++) enq[i] = 0;
    front = 0; has_back = false;
  }
};
struct TopOrderQueue : Queue {  // queues/top_order_queue.rs:20-42,45-70
  std::vector<StateId> order;
  std::vector<int64_t> state;  // -1 = None
  StateId front = 0; bool has_back = false; StateId back = 0;
  explicit TopOrderQueue(std::vector<StateId> ord) : order(std::move(ord)), state(order.size(), -1) {}
  void enqueue(StateId s) override {
    StateId o = order[s];
    if (!has_back || front > back) { front = o; back = o; has_back = true; }
    else if (o > back) back = o;
    else if (o < front) front = o;
    state[o] = s;
  }
  bool is_empty() const override { return has_back ? front > back : true; }
  bool dequeue(StateId* out) override {
    if (is_empty()) return false;
    int64_t old_head = state[front];
    state[front] = -1;
    while (front <= back && state[front] < 0) front++;
    if (old_head < 0) return false;
    *out = (StateId)old_head;
    return true;
  }
  void clear() override {
    if (has_back) for (StateId s = front; s <= back && s < state.size(); s++) state[s] = -1;
    front = 0; has_back = false;
  }
};
struct SccQueue : Queue {  // queues/scc_queue.rs:16-62
  int64_t front = 0, back = -1;
  std::vector<std::unique_ptr<Queue>> queues;
  std::vector<StateId> sccs;
  SccQueue(std::vector<std::unique_ptr<Queue>> q, std::vector<StateId> s) : queues(std::move(q)), sccs(std::move(s)) {}
  void update_front() { while (front <= back && queues[front]->is_empty()) front++; }
  void enqueue(StateId s) override {
    int64_t c = sccs[s];
    if (front > back) { front = c; back = c; }
    else if (c > back) back = c;
    else if (c < front) front = c;
    queues[c]->enqueue(s);
  }
  bool is_empty() const override {
    if (front < back) return false;
    if (front > back) return true;
    return queues[front]->is_empty();
  }
  bool dequeue(StateId* o) override {
    if (is_empty()) return false;
    update_front();
    return queues[front]->dequeue(o);
  }
  void update(StateId s) override { queues[sccs[s]]->update(s); }
  void clear() override {
    for (int64_t i = front; i <= back; i++) queues[i]->clear();
    front = 0; back = -1;
  }
};

// algorithms/top_sort.rs:12-61 TopOrderVisitor
struct TopOrderVisitor : Visitor {
  std::vector<StateId> order, finish;
  bool acyclic = true;
  bool init_state(StateId, StateId) override { return true; }
  bool tree_tr(StateId, const Tr&) override { return true; }
  bool back_tr(StateId, const Tr&) override { acyclic = false; return false; }
  bool forward_or_cross_tr(StateId, const Tr&) override { return true; }
  void finish_state(StateId s, bool, StateId) override { finish.push_back(s); }
  void finish_visit() override {
    if (acyclic) {
      order.assign(finish.size(), 0);
      for (size_t s = 0; s < finish.size(); s++) order[finish[finish.size() - s - 1]] = (StateId)s;
    }
  }
};

// algorithms/state_sort.rs:16-78.  The reference permutes in place along the cycles of `order` (old state s becomes
// state order[s], arcs keep their order, nextstates are mapped) and finally overwrites the property word with the
// stored word restricted to statesort_properties() (properties.rs:319-349), so the mutations in between leave no trace.
inline void state_sort(Fst& fst, const std::vector<StateId>& order) {
  if (order.size() != fst.num_states())
    throw std::runtime_error("StateSort : Bad order vector size : " + std::to_string(order.size()) + ". Expected " +
                             std::to_string(fst.num_states()));
  if (!fst.has_start) return;
  const uint64_t statesort_mask = P::ALL & ~(P::TOP_SORTED | P::NOT_TOP_SORTED | P::STRING | P::NOT_STRING);
  const uint64_t props = fst.props & statesort_mask;
  std::vector<State> ns(fst.states.size());
  for (size_t s = 0; s < fst.states.size(); s++) {
    State st = fst.states[s];
    for (Tr& tr : st.trs) tr.nextstate = order[tr.nextstate];
    ns[order[s]] = std::move(st);
  }
  fst.states.swap(ns);
  fst.start = order[fst.start];
  fst.set_properties_with_mask(props, P::ALL);
}

// algorithms/top_sort.rs:75-95
inline void top_sort(Fst& fst) {
  TopOrderVisitor v;
  dfs_visit(fst, v, false);
  if (v.acyclic) {
    state_sort(fst, v.order);
    const uint64_t p = P::ACYCLIC | P::INITIAL_ACYCLIC | P::TOP_SORTED;
    fst.set_properties_with_mask(p, p);
  } else {
    const uint64_t p = P::CYCLIC | P::NOT_TOP_SORTED;
    fst.set_properties_with_mask(p, p);
  }
}

enum QueueKind { QK_STATE_ORDER, QK_TOP_ORDER, QK_LIFO, QK_SCC };

// queues/auto_queue.rs:23-99 (distance = None => less = None) and :101-157 scc_queue_type
inline std::unique_ptr<Queue> make_auto_queue(const Fst& fst, QueueKind* kind_out = nullptr) {
  uint64_t props = fst.props;
  if ((props & P::TOP_SORTED) || !fst.has_start) {
    if (kind_out) *kind_out = QK_STATE_ORDER;
    return std::make_unique<StateOrderQueue>();
  }
  if (props & P::ACYCLIC) {
    TopOrderVisitor v;
    dfs_visit(fst, v, false);
    if (!v.acyclic) throw std::runtime_error("Unexpectted Acyclic FST for TopOprerQueue");
    // note: with an incomplete DFS order.len() may be < num_states; the reference indexes order[state]
    if (kind_out) *kind_out = QK_TOP_ORDER;
    return std::make_unique<TopOrderQueue>(std::move(v.order));
  }
  if (props & P::UNWEIGHTED) {  // TropicalWeight is IDEMPOTENT
    if (kind_out) *kind_out = QK_LIFO;
    return std::make_unique<LifoQueue>();
  }
  SccVisitor sv(fst);
  dfs_visit(fst, sv, false);
  std::vector<StateId> sccs(sv.scc.begin(), sv.scc.end());
  size_t n_sccs = (size_t)sv.nscc;
  enum QT { Trivial, Fifo, Lifo, ShortestFirst };
  std::vector<int> qt(n_sccs, Trivial);
  bool all_trivial = true, unweighted = true;
  for (size_t state = 0; state < fst.num_states(); state++) {
    for (const Tr& tr : fst.states[state].trs) {
      if (sccs[state] == sccs[tr.nextstate]) {
        // compare.is_none() => FifoQueue (auto_queue.rs:128-131)
        qt[sccs[state]] = Fifo;
        if (qt[sccs[state]] != Trivial) all_trivial = false;
      }
      if (!w_is_zero(tr.weight) && !w_is_one(tr.weight)) unweighted = false;
    }
  }
  if (unweighted) {
    if (kind_out) *kind_out = QK_LIFO;
    return std::make_unique<LifoQueue>();
  }
  if (all_trivial) {
    if (kind_out) *kind_out = QK_TOP_ORDER;
    return std::make_unique<TopOrderQueue>(std::move(sccs));
  }
  std::vector<std::unique_ptr<Queue>> queues;
  for (size_t i = 0; i < n_sccs; i++) {
    if (qt[i] == Trivial) queues.push_back(std::make_unique<TrivialQueue>());
    else queues.push_back(std::make_unique<FifoQueue>());
  }
  if (kind_out) *kind_out = QK_SCC;
  return std::make_unique<SccQueue>(std::move(queues), std::move(sccs));
}

struct ShortestPathConfig {  // shortest_path.rs:24-60
  float delta = 1e-6f;
  size_t nshortest = 1;
  bool unique = false;
};
struct SsspStats { uint64_t arcs_relaxed = 0, states_dequeued = 0; };

// fst_properties/mutate_properties.rs:662-672
inline uint64_t shortest_path_properties(uint64_t props, bool tree) {
  uint64_t out = props | P::ACYCLIC | P::INITIAL_ACYCLIC | P::ACCESSIBLE | P::UNWEIGHTED_CYCLES;
  if (!tree) out |= P::COACCESSIBLE;
  return out;
}

// ---- n > 1: shortest_distance + reverse + heap n-best -------------------------------------------------------
// semirings/utils_float.rs:1-3 via TropicalWeight::approx_equal (tropical_weight.rs:72-74).  NOT the KDELTA `==`:
// |inf - inf| is NaN, so two infinite weights are *not* approx_equal (the relaxation below relies on it verbatim).
inline bool w_approx_equal(float w1, float w2, float delta) { return std::fabs(w1 - w2) <= delta; }

// algorithms/shortest_distance.rs:153-237 with reverse = false (:318-323), AnyTrFilter, first_path = false,
// retain = false.  The returned vector grows on demand (ensure_distance_index_is_valid, :137-144): its length is
// 1 + the largest state index touched, not num_states().
inline std::vector<float> shortest_distance(const Fst& fst, float delta) {
  std::vector<float> distance, adder, radder;
  std::vector<uint8_t> enqueued;
  if (!fst.has_start) return distance;  // :158-161
  std::unique_ptr<Queue> queue = make_auto_queue(fst);  // :319
  auto ensure = [&](size_t index) {
    while (distance.size() <= index) {
      distance.push_back(W_ZERO); enqueued.push_back(0); adder.push_back(W_ZERO); radder.push_back(W_ZERO);
    }
  };
  queue->clear();
  size_t source = fst.start;
  ensure(source);
  distance[source] = W_ONE; adder[source] = W_ONE; radder[source] = W_ONE;
  enqueued[source] = 1;
  queue->enqueue((StateId)source);
  StateId st;
  while (queue->dequeue(&st)) {
    size_t state = st;
    enqueued[state] = 0;
    float r = radder[state];
    radder[state] = W_ZERO;
    for (const Tr& tr : fst.states[state].trs) {
      size_t nextstate = tr.nextstate;
      ensure(nextstate);
      float weight = w_times(r, tr.weight);
      if (!w_approx_equal(distance[nextstate], w_plus(distance[nextstate], weight), delta)) {  // :217
        adder[nextstate] = w_plus(adder[nextstate], weight);
        distance[nextstate] = adder[nextstate];
        radder[nextstate] = w_plus(radder[nextstate], weight);
        if (!enqueued[state]) {  // :224 — tests `state`, not `nextstate` (reference quirk, kept)
          queue->enqueue((StateId)nextstate);
          enqueued[nextstate] = 1;
        } else {
          queue->update((StateId)nextstate);
        }
      }
    }
  }
  return distance;
}

// fst_properties/mutate_properties.rs:622-638
inline uint64_t reverse_properties(uint64_t inprops, bool has_superinitial) {
  uint64_t out = (P::ACCEPTOR | P::NOT_ACCEPTOR | P::EPSILONS | P::I_EPSILONS | P::O_EPSILONS | P::UNWEIGHTED |
                  P::CYCLIC | P::ACYCLIC | P::WEIGHTED_CYCLES | P::UNWEIGHTED_CYCLES) & inprops;
  if (has_superinitial) out |= P::WEIGHTED & inprops;
  return out;
}

// algorithms/reverse.rs:33-87 (TropicalWeight: reverse() is the identity), incl. the property word of :78-83.
// Pinned by rustfst-python/tests/algorithms/test_reverse.py (tests/test_oracle_kat.py).
inline Fst reverse(const Fst& ifst) {
  Fst ofst;
  StateId ostart = ofst.add_state();
  ofst.add_states(ifst.num_states());
  std::vector<std::vector<Tr>> states_trs(ifst.num_states() + 1);
  for (size_t is = 0; is < ifst.num_states(); is++) {
    StateId os = (StateId)is + 1;
    if (ifst.has_start && ifst.start == is) ofst.set_final(os, W_ONE);
    const State& st = ifst.states[is];
    if (st.has_final) states_trs[0].push_back(Tr{EPS_LABEL, EPS_LABEL, st.final_weight, os});
    for (const Tr& itr : st.trs) states_trs[itr.nextstate + 1].push_back(Tr{itr.ilabel, itr.olabel, itr.weight, os});
  }
  for (size_t s = 0; s < states_trs.size(); s++) ofst.set_trs_unchecked((StateId)s, std::move(states_trs[s]));
  ofst.set_start(ostart);
  ofst.set_properties_with_mask(reverse_properties(ifst.props, true) | ofst.props, P::ALL);
  return ofst;
}

inline bool natural_less(float w1, float w2) {  // shortest_path.rs:284-286 (`==`/`!=` are the KDELTA-approximate ones)
  return w_eq(w_plus(w1, w2), w1) && !w_eq(w1, w2);
}

// shortest_path.rs:409-518 n_shortest_path over the reversed machine, with ShortestPathCompare (:288-339) and the
// hand-rolled binary heap (:341-407) restated verbatim (the pop order among ties decides state numbering).
inline Fst n_shortest_path(const Fst& ifst, const std::vector<float>& distance, size_t nshortest, float delta) {
  Fst ofst;
  if (nshortest == 0) return ofst;
  if (!ifst.has_start || distance.size() <= ifst.start || w_is_zero(distance[ifst.start])) return ofst;  // :427-434
  StateId istart = ifst.start;
  StateId ostart = ofst.add_state();
  ofst.set_start(ostart);
  StateId final_state = ofst.add_state();
  ofst.set_final(final_state, W_ONE);
  struct Pair { bool some; StateId state; float w; };
  std::vector<Pair> pairs(final_state + 1, Pair{false, 0, W_ZERO});
  pairs[final_state] = Pair{true, istart, W_ONE};

  auto pweight = [&](const Pair& p) -> float {  // :309-321
    if (p.some) return p.state < distance.size() ? distance[p.state] : W_ZERO;
    return W_ONE;
  };
  auto compare = [&](StateId x, StateId y) -> bool {  // :323-338
    const Pair& px = pairs[x];
    const Pair& py = pairs[y];
    float wx = w_times(pweight(px), px.w);
    float wy = w_times(pweight(py), py.w);
    if (!px.some && py.some) return natural_less(wy, wx) || w_approx_equal(wx, wy, delta);
    if (px.some && !py.some) return natural_less(wy, wx) && !w_approx_equal(wx, wy, delta);
    return natural_less(wy, wx);
  };
  std::vector<StateId> heap;
  std::function<void(size_t)> sift_up = [&](size_t idx) {  // :361-369
    if (idx > 0) {
      size_t parent_idx = (idx - 1) / 2;
      if (compare(heap[parent_idx], heap[idx])) { std::swap(heap[idx], heap[parent_idx]); sift_up(parent_idx); }
    }
  };
  std::function<void(size_t)> sift_down = [&](size_t idx) {  // :374-392
    StateId cur_val = heap[idx];
    size_t c1 = 2 * idx + 1, c2 = 2 * idx + 2, big;
    if (c1 >= heap.size() && c2 >= heap.size()) return;
    else if (c1 < heap.size() && c2 >= heap.size()) big = c1;
    else if (compare(heap[c1], heap[c2])) big = c2;
    else big = c1;
    if (!compare(heap[big], cur_val)) { std::swap(heap[idx], heap[big]); sift_down(big); }
  };
  auto push = [&](StateId v) { heap.push_back(v); sift_up(heap.size() - 1); };
  auto pop = [&]() -> StateId {  // :393-402
    StateId top = heap[0];
    if (heap.size() == 1) heap.clear();
    else { heap[0] = heap.back(); heap.pop_back(); sift_down(0); }
    return top;
  };
  push(final_state);
  float limit = w_times(distance[istart], W_ZERO);  // weight_threshold = zero (:448-449)
  std::vector<size_t> r;
  while (!heap.empty()) {
    StateId state = pop();
    Pair p = pairs[state];
    size_t p_first_real = p.some ? (size_t)p.state + 1 : 0;
    float d = p.some ? (p.state < distance.size() ? distance[p.state] : W_ZERO) : W_ONE;
    if (natural_less(limit, w_times(d, p.w))) continue;
    while (r.size() <= p_first_real) r.push_back(0);
    r[p_first_real] += 1;
    if (!p.some) ofst.add_tr(ofst.start, Tr{0, 0, W_ONE, state});
    if (!p.some && r[p_first_real] == nshortest) break;
    if (r[p_first_real] > nshortest) continue;
    if (!p.some) continue;
    for (const Tr& rarc : ifst.states[p.state].trs) {
      Tr tr{rarc.ilabel, rarc.olabel, rarc.weight, rarc.nextstate};
      float weight = w_times(p.w, tr.weight);
      StateId next = ofst.add_state();
      pairs.push_back(Pair{true, tr.nextstate, weight});
      tr.nextstate = state;
      ofst.add_tr(next, tr);
      push(next);
    }
    const State& ist = ifst.states[p.state];
    if (ist.has_final && !w_is_zero(ist.final_weight)) {
      float weight = w_times(p.w, ist.final_weight);
      StateId next = ofst.add_state();
      pairs.push_back(Pair{false, 0, weight});
      ofst.add_tr(next, Tr{0, 0, ist.final_weight, state});
      push(next);
    }
  }
  connect(ofst);
  ofst.set_properties_with_mask(shortest_path_properties(ofst.props, false), P::ALL);
  return ofst;
}

// determinize_with_distance — algorithms/determinize/determinize_static.rs:24-39 over DeterminizeFsaOp
// (determinize_fsa_op.rs:44-54 start, :56-101 arcs of a subset state, :103-120 final weight, :147-178 norm_tr), the state
// table (state_table.rs:20-35 out_dist of a new subset, :76-96 ids in order of first lookup), DefaultCommonDivisor
// (divisors.rs:16-22: plus), quantisation (semiring.rs:132-145), division (tropical_weight.rs:127-132: plain
// subtraction) and the breadth-first materialisation of a lazy FST (lazy/lazy_fst.rs:226-259).
// Subsets are kept sorted by state (see the header of this file for why the reference's own order is not reproducible).
struct DetElement { StateId state; float weight; };
struct DetTuple { std::vector<DetElement> subset; StateId filter_state; };
struct DetResult { Fst fst; std::vector<float> out_dist; };
inline float w_quantize(float v, float delta) {  // semiring.rs:135-142
  if (std::isinf(v)) return v;
  return std::floor((v / delta) + 0.5f) * delta;
}
inline DetResult determinize_with_distance(const Fst& ifst, const std::vector<float>& in_dist, float delta) {
  if (!(ifst.props & P::ACCEPTOR))  // determinize_fsa_op.rs:137-139
    throw std::runtime_error("DeterminizeFsaImpl : expected acceptor as argument");
  DetResult res;
  std::vector<DetTuple> tuples;
  std::map<std::vector<uint64_t>, StateId> ids;
  auto find_state = [&](const DetTuple& t) -> StateId {  // state_table.rs:76-96
    std::vector<uint64_t> key;
    key.reserve(t.subset.size() + 1);
    key.push_back(t.filter_state);
    for (const DetElement& e : t.subset) {
      float w = e.weight == 0.0f ? 0.0f : e.weight;  // -0.0 and 0.0 hash alike (OrderedFloat)
      uint32_t bits;
      std::memcpy(&bits, &w, 4);
      key.push_back(((uint64_t)e.state << 32) | bits);
    }
    auto it = ids.find(key);
    if (it != ids.end()) return it->second;
    StateId id = (StateId)tuples.size();
    ids.emplace(std::move(key), id);
    tuples.push_back(t);
    float outd = W_ZERO;  // state_table.rs:20-35
    for (const DetElement& e : t.subset)
      outd = w_plus(outd, w_times(e.weight, e.state < in_dist.size() ? in_dist[e.state] : W_ZERO));
    res.out_dist.push_back(outd);
    return id;
  };
  if (!ifst.has_start) return res;  // lazy_fst.rs:227-232
  StateId start = find_state(DetTuple{{DetElement{ifst.start, W_ONE}}, ifst.start});  // determinize_fsa_op.rs:44-54
  Fst& out = res.fst;
  out.add_states((size_t)start + 1);
  out.set_start(start);
  std::deque<StateId> queue;
  std::vector<bool> visited((size_t)start + 1, false);
  visited[start] = true;
  queue.push_back(start);
  struct DetTr { Label label; float weight; DetTuple dest; };
  while (!queue.empty()) {
    StateId s = queue.front();
    queue.pop_front();
    const DetTuple src = tuples[s];  // a copy: find_state below grows `tuples`
    std::map<Label, DetTr> label_map;  // BTreeMap: arcs leave in label order (:57-82)
    for (const DetElement& se : src.subset)
      for (const Tr& tr : ifst.states[se.state].trs) {
        auto it = label_map.find(tr.ilabel);
        if (it == label_map.end()) it = label_map.emplace(tr.ilabel, DetTr{tr.ilabel, W_ZERO, DetTuple{{}, 0}}).first;
        it->second.dest.subset.push_back(DetElement{tr.nextstate, w_times(se.weight, tr.weight)});
      }
    std::vector<Tr> trs;
    for (auto& kv : label_map) {  // norm_tr (:147-178)
      DetTr& d = kv.second;
      std::stable_sort(d.dest.subset.begin(), d.dest.subset.end(),
                       [](const DetElement& x, const DetElement& y) { return x.state < y.state; });
      for (const DetElement& e : d.dest.subset) d.weight = w_plus(d.weight, e.weight);
      std::vector<DetElement> merged;  // one element per state, weights combined with plus; ascending state
      for (const DetElement& e : d.dest.subset) {
        if (!merged.empty() && merged.back().state == e.state) merged.back().weight = w_plus(merged.back().weight, e.weight);
        else merged.push_back(e);
      }
      for (DetElement& e : merged) e.weight = w_quantize(e.weight - d.weight, delta);
      d.dest.subset = std::move(merged);
    }
    for (auto& kv : label_map) {  // :92-99
      const DetTr& d = kv.second;
      trs.push_back(Tr{d.label, d.label, d.weight, find_state(d.dest)});
    }
    for (const Tr& tr : trs) {  // lazy_fst.rs:242-254
      if (tr.nextstate >= visited.size()) visited.resize((size_t)tr.nextstate + 1, false);
      if (!visited[tr.nextstate]) { queue.push_back(tr.nextstate); visited[tr.nextstate] = true; }
      if (tr.nextstate >= out.num_states()) out.add_states((size_t)tr.nextstate - out.num_states() + 1);
    }
    out.set_trs_unchecked(s, std::move(trs));
    float fw = W_ZERO;  // determinize_fsa_op.rs:103-120
    for (const DetElement& e : src.subset) {
      const State& st = ifst.states[e.state];
      fw = w_plus(fw, w_times(e.weight, st.has_final ? st.final_weight : W_ZERO));
    }
    if (!w_is_zero(fw)) out.set_final(s, fw);
  }
  out.props = 0;  // lazy_fst.rs:260 with DeterminizeFsaOp::properties() = empty (:122-125); nobody reads it
  return res;
}

// shortest_path.rs:107-171 (dispatch), :173-239 single_shortest_path, :241-282 backtrace
inline Fst shortest_path(const Fst& ifst, const ShortestPathConfig& cfg, SsspStats* stats = nullptr,
                         std::vector<float>* distance_out = nullptr) {
  if (cfg.nshortest == 0) return Fst();
  if (cfg.nshortest != 1) {  // :135-170
    std::vector<float> distance = shortest_distance(ifst, cfg.delta);
    Fst rfst = reverse(ifst);
    float d = W_ZERO;
    for (const Tr& rarc : rfst.states[0].trs) {
      size_t state = rarc.nextstate - 1;
      if (state < distance.size()) d = w_plus(d, w_times(rarc.weight, distance[state]));
    }
    std::vector<float> distance_2;
    distance_2.reserve(distance.size() + 1);
    distance_2.push_back(d);
    distance_2.insert(distance_2.end(), distance.begin(), distance.end());
    if (distance_out) *distance_out = distance;
    if (cfg.unique) {  // :156-165 (reverse weights of the tropical semiring are the weights themselves)
      DetResult det = determinize_with_distance(rfst, distance_2, cfg.delta);
      return n_shortest_path(det.fst, det.out_dist, cfg.nshortest, cfg.delta);
    }
    return n_shortest_path(rfst, distance_2, cfg.nshortest, cfg.delta);
  }

  std::vector<float> distance;
  std::vector<int64_t> parent_state;  // -1 = None
  std::vector<size_t> parent_pos;
  bool has_f_parent = false;
  StateId f_parent = 0;

  if (ifst.has_start) {
    std::unique_ptr<Queue> queue = make_auto_queue(ifst);
    StateId source = ifst.start;
    float f_distance = W_ZERO;
    queue->clear();
    size_t n = ifst.num_states();
    distance.assign(n, W_ZERO);
    std::vector<uint8_t> enqueued(n, 0);
    parent_state.assign(n, -1);
    parent_pos.assign(n, 0);
    distance[source] = W_ONE;
    enqueued[source] = 1;
    queue->enqueue(source);
    StateId s;
    while (queue->dequeue(&s)) {
      enqueued[s] = 0;
      float sd = distance[s];
      if (stats) stats->states_dequeued++;
      const State& st = ifst.states[s];
      if (st.has_final) {
        float plus = w_plus(f_distance, w_times(sd, st.final_weight));
        if (!w_eq(f_distance, plus)) { f_distance = plus; f_parent = s; has_f_parent = true; }
      }
      for (size_t pos = 0; pos < st.trs.size(); pos++) {
        const Tr& tr = st.trs[pos];
        float& nd = distance[tr.nextstate];
        float weight = w_times(sd, tr.weight);
        float p = w_plus(nd, weight);
        if (!w_eq(nd, p)) {
          nd = p;
          parent_state[tr.nextstate] = s;
          parent_pos[tr.nextstate] = pos;
          if (!enqueued[tr.nextstate]) { queue->enqueue(tr.nextstate); enqueued[tr.nextstate] = 1; }
          else queue->update(tr.nextstate);
        }
      }
      if (stats) stats->arcs_relaxed += st.trs.size();
    }
  }
  if (distance_out) *distance_out = distance;

  // backtrace
  Fst ofst;
  bool has_sp = false, has_dp = false, has_d = false;
  StateId s_p = 0, d_p = 0, d = 0;
  bool has_next = has_f_parent;
  StateId nextstate = f_parent;
  while (has_next) {
    StateId state = nextstate;
    d_p = s_p; has_dp = has_sp;
    s_p = ofst.add_state(); has_sp = true;
    if (has_d) {
      size_t pos = parent_pos[d];
      Tr tr = ifst.states[state].trs[pos];
      (void)has_dp;
      tr.nextstate = d_p;
      ofst.add_tr(s_p, tr);
    } else if (ifst.states[f_parent].has_final) {
      ofst.set_final(s_p, ifst.states[f_parent].final_weight);
    }
    d = state; has_d = true;
    if (parent_state[state] >= 0) { nextstate = (StateId)parent_state[state]; has_next = true; }
    else has_next = false;
  }
  if (has_sp) ofst.set_start(s_p);
  ofst.set_properties_with_mask(shortest_path_properties(ofst.props, true), P::ALL);
  return ofst;
}

// ---------------------------------------------------------------------------------------------
// OpenFst binary I/O — parsers/bin_fst/fst_header.rs:71-137, vector_fst/serializable_fst.rs:29-88,129-168,
// const_fst/serializable_fst.rs (read only)
// ---------------------------------------------------------------------------------------------
constexpr int32_t FST_MAGIC = 2125659606;

struct Reader {
  const uint8_t* p; size_t n; size_t off = 0;
  template <class T> T get() {
    if (off + sizeof(T) > n) throw std::runtime_error("Error while parsing binary VectorFst: truncated");
    T v; std::memcpy(&v, p + off, sizeof(T)); off += sizeof(T); return v;
  }
  std::string str() {
    int32_t len = get<int32_t>();
    if (len < 0 || off + (size_t)len > n) throw std::runtime_error("bad string");
    std::string s((const char*)p + off, (size_t)len); off += (size_t)len; return s;
  }
  void align(size_t a) { size_t r = off % a; if (r) off += a - r; }
};

inline void skip_symt(Reader& r) {  // parsers/bin_symt/nom_parser.rs (table is dropped by the oracle)
  int32_t magic = r.get<int32_t>(); (void)magic;
  r.str();
  r.get<int64_t>();            // available key
  int64_t n = r.get<int64_t>();
  for (int64_t i = 0; i < n; i++) { r.str(); r.get<int64_t>(); }
}

inline Fst fst_from_bytes(const uint8_t* data, size_t len, bool allow_const = false) {
  Reader r{data, len};
  if (r.get<int32_t>() != FST_MAGIC) throw std::runtime_error("bad magic number");
  std::string fst_type = r.str();
  std::string arc_type = r.str();
  if (arc_type != "standard") throw std::runtime_error("arc type is not standard");
  int32_t version = r.get<int32_t>();
  uint32_t flags = r.get<uint32_t>();
  uint64_t props = r.get<uint64_t>();
  int64_t start = r.get<int64_t>();
  int64_t num_states = r.get<int64_t>();
  int64_t num_arcs = r.get<int64_t>();
  if (flags & 1) skip_symt(r);
  if (flags & 2) skip_symt(r);
  Fst f;
  f.props = props & P::TRINARY;  // FstProperties::from_bits_truncate
  f.has_start = start != -1;
  f.start = (StateId)start;
  f.states.resize((size_t)num_states);
  if (fst_type == "vector") {
    if (version < 2) throw std::runtime_error("vector version < 2");
    for (int64_t s = 0; s < num_states; s++) {
      State& st = f.states[s];
      float fw = r.get<float>();
      // utils_parsing.rs:17-26 parse_final_weight: None iff approx-equal to zero()
      if (!w_eq(fw, W_ZERO)) { st.has_final = true; st.final_weight = fw; }
      int64_t narcs = r.get<int64_t>();
      st.trs.resize((size_t)narcs);
      for (int64_t a = 0; a < narcs; a++) {
        Tr& t = st.trs[a];
        t.ilabel = (Label)r.get<int32_t>(); t.olabel = (Label)r.get<int32_t>();
        t.weight = r.get<float>(); t.nextstate = (StateId)r.get<int32_t>();
        if (t.ilabel == EPS_LABEL) st.niepsilons++;
        if (t.olabel == EPS_LABEL) st.noepsilons++;
      }
    }
  } else if (fst_type == "const" && allow_const) {
    bool aligned = (version == 1);  // const_fst/serializable_fst.rs:210 (CONST_ALIGNED_FILE_VERSION)
    if (aligned) r.align(16);
    struct CS { float fw; int32_t pos, narcs, nie, noe; };
    std::vector<CS> cs((size_t)num_states);
    for (auto& c : cs) { c.fw = r.get<float>(); c.pos = r.get<int32_t>(); c.narcs = r.get<int32_t>();
                         c.nie = r.get<int32_t>(); c.noe = r.get<int32_t>(); }
    if (aligned) r.align(16);
    std::vector<Tr> arcs((size_t)num_arcs);
    for (auto& t : arcs) { t.ilabel = (Label)r.get<int32_t>(); t.olabel = (Label)r.get<int32_t>();
                           t.weight = r.get<float>(); t.nextstate = (StateId)r.get<int32_t>(); }
    for (int64_t s = 0; s < num_states; s++) {
      State& st = f.states[s];
      if (!w_eq(cs[s].fw, W_ZERO)) { st.has_final = true; st.final_weight = cs[s].fw; }
      st.trs.assign(arcs.begin() + cs[s].pos, arcs.begin() + cs[s].pos + cs[s].narcs);
      for (auto& t : st.trs) { if (t.ilabel == 0) st.niepsilons++; if (t.olabel == 0) st.noepsilons++; }
    }
  } else {
    throw std::runtime_error("fst type is not vector");
  }
  return f;
}

struct Writer {
  std::vector<uint8_t> buf;
  template <class T> void put(T v) { size_t o = buf.size(); buf.resize(o + sizeof(T)); std::memcpy(buf.data() + o, &v, sizeof(T)); }
  void str(const std::string& s) { put<int32_t>((int32_t)s.size()); buf.insert(buf.end(), s.begin(), s.end()); }
};

// vector_fst/serializable_fst.rs:46-88
inline std::vector<uint8_t> fst_to_bytes(const Fst& f) {
  Writer w;
  w.put<int32_t>(FST_MAGIC);
  w.str("vector"); w.str("standard");
  w.put<int32_t>(2);
  w.put<uint32_t>(0);
  w.put<uint64_t>(f.props | P::EXPANDED | P::MUTABLE);
  w.put<int64_t>(f.has_start ? (int64_t)f.start : -1);
  w.put<int64_t>((int64_t)f.num_states());
  w.put<int64_t>((int64_t)f.num_trs_total());
  for (const State& st : f.states) {
    w.put<float>(st.has_final ? st.final_weight : W_ZERO);
    w.put<int64_t>((int64_t)st.trs.size());
    for (const Tr& t : st.trs) { w.put<int32_t>((int32_t)t.ilabel); w.put<int32_t>((int32_t)t.olabel);
                                 w.put<float>(t.weight); w.put<int32_t>((int32_t)t.nextstate); }
  }
  return std::move(w.buf);
}

inline std::vector<uint8_t> read_file(const std::string& path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) throw std::runtime_error("cannot open " + path);
  return std::vector<uint8_t>((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
}
inline void write_file(const std::string& path, const std::vector<uint8_t>& b) {
  std::ofstream out(path, std::ios::binary);
  if (!out) throw std::runtime_error("cannot write " + path);
  out.write((const char*)b.data(), (std::streamsize)b.size());
}

}  // namespace oracle
