// queue_plan.cpp — host-side mirror of AutoQueue::new: which queue discipline rustfst's single_shortest_path would
// run with for a given FST, decided from the STORED property bits exactly as the reference does
// (rustfst/src/algorithms/queues/auto_queue.rs:23-99, scc_queue_type :101-157), plus the DFS-derived orders it
// needs (rustfst/src/algorithms/top_sort.rs:12-61 TopOrderVisitor, rustfst/src/algorithms/visitors/scc_visitors.rs,
// rustfst/src/algorithms/dfs_visit.rs:97-187).  This is queue *construction* (control plane): a sequential,
// lexicographic DFS whose visiting order defines the tie-breaking of the reference.  The relaxation itself runs on
// the device (sssp.cu).
#include <chrono>

#include "algos.h"

namespace b200 {
namespace {

// Iterative DFS over the CSR in the reference's order: root = start, then every still-white state 0,1,2,...;
// arcs in stored order.  Produces the finish order and, when asked, Tarjan SCC numbers (completion order).
struct DfsResult {
  std::vector<uint32_t> finish;  // states in finish order
  std::vector<int32_t> scc;      // completion-order SCC number per state (want_scc)
  int32_t nscc = 0;
  bool acyclic = true;
};

DfsResult dfs(const CsrFst& f, bool want_scc, bool stop_on_back_arc) {
  const size_t n = f.num_states();
  DfsResult r;
  r.finish.reserve(n);
  if (!f.has_start) return r;
  enum : uint8_t { kWhite = 0, kGrey = 1, kBlack = 2 };
  std::vector<uint8_t> color(n, kWhite);
  std::vector<int32_t> dfnumber, lowlink;
  std::vector<uint8_t> onstack;
  std::vector<uint32_t> scc_stack;
  if (want_scc) { r.scc.assign(n, -1); dfnumber.assign(n, -1); lowlink.assign(n, -1); onstack.assign(n, 0); }
  int32_t nvisited = 0;
  struct Frame { uint32_t s, pos; };
  std::vector<Frame> stack;
  auto init_state = [&](uint32_t s) {
    if (want_scc) { scc_stack.push_back(s); dfnumber[s] = lowlink[s] = nvisited; onstack[s] = 1; }
    nvisited++;
  };
  auto finish_state = [&](uint32_t s, bool has_parent, uint32_t parent) {
    r.finish.push_back(s);
    if (!want_scc) return;
    if (dfnumber[s] == lowlink[s]) {
      uint32_t t;
      do { t = scc_stack.back(); r.scc[t] = r.nscc; onstack[t] = 0; scc_stack.pop_back(); } while (t != s);
      r.nscc++;
    }
    if (has_parent && lowlink[s] < lowlink[parent]) lowlink[parent] = lowlink[s];
  };
  bool go = true;
  size_t root = f.start;
  while (go && root < n) {
    color[root] = kGrey;
    stack.push_back({(uint32_t)root, f.offsets[root]});
    init_state((uint32_t)root);
    while (!stack.empty()) {
      Frame& fr = stack.back();
      uint32_t s = fr.s;
      if (!go || fr.pos >= f.offsets[s + 1]) {
        color[s] = kBlack;
        stack.pop_back();
        if (!stack.empty()) { finish_state(s, true, stack.back().s); stack.back().pos++; }
        else finish_state(s, false, 0);
        continue;
      }
      uint32_t t = f.arcs[fr.pos].nextstate;
      if (color[t] == kWhite) {
        color[t] = kGrey;
        stack.push_back({t, f.offsets[t]});
        init_state(t);
      } else if (color[t] == kGrey) {  // back arc
        r.acyclic = false;
        if (stop_on_back_arc) go = false;  // TopOrderVisitor::back_tr returns false (top_sort.rs:40-43)
        if (want_scc && dfnumber[t] < lowlink[s]) lowlink[s] = dfnumber[t];
        stack.back().pos++;
      } else {  // forward or cross arc
        if (want_scc && dfnumber[t] < dfnumber[s] && onstack[t] && dfnumber[t] < lowlink[s]) lowlink[s] = dfnumber[t];
        stack.back().pos++;
      }
    }
    root = (root == f.start) ? 0 : root + 1;
    while (root < n && color[root] != kWhite) root++;
  }
  return r;
}

// The same traversal specialised for the TopOrderVisitor (no SCC bookkeeping) and tuned for multi-million-state
// lattices, where it is the one sequential host step of a shortest-path call: the walk is a chain of dependent cache
// misses (offsets of the state -> its arcs -> colour of every target), so the colours of all targets of a state are
// prefetched when the state is first visited and the frame keeps the end of the arc range (C4 lattice, 5M states /
// 50M arcs: 2.05 s -> 1.15 s on the development container).  Tried and dropped: prefetching the targets' offsets as
// well (no gain), colour and offset fused in one 64-bit word + look-ahead prefetch of the next target's arcs (slower:
// the byte-sized colour array stays cache resident, the 8-byte words do not).  Identical visiting order, hence
// identical finish order.
DfsResult dfs_top_order(const CsrFst& f) {
  const size_t n = f.num_states();
  DfsResult r;
  r.finish.reserve(n);
  if (!f.has_start) return r;
  enum : uint8_t { kWhite = 0, kGrey = 1, kBlack = 2 };
  std::vector<uint8_t> color(n, kWhite);
  struct Frame { uint32_t s, pos, end; };
  std::vector<Frame> stack;
  stack.reserve(1024);
  const uint32_t* off = f.offsets.data();
  const Tr* arcs = f.arcs.data();
  uint8_t* col = color.data();
  auto push = [&](uint32_t t) {
    col[t] = kGrey;
    const uint32_t lo = off[t], hi = off[t + 1];
    for (uint32_t e = lo; e < hi; e++) __builtin_prefetch(&col[arcs[e].nextstate], 0, 1);
    stack.push_back({t, lo, hi});
  };
  size_t root = f.start;
  while (root < n) {
    push((uint32_t)root);
    while (!stack.empty()) {
      Frame& fr = stack.back();
      if (fr.pos >= fr.end) {
        col[fr.s] = kBlack;
        r.finish.push_back(fr.s);
        stack.pop_back();
        if (!stack.empty()) stack.back().pos++;
        continue;
      }
      const uint32_t t = arcs[fr.pos].nextstate;
      const uint8_t c = col[t];
      if (c == kWhite) {
        push(t);
      } else if (c == kGrey) {  // back arc: TopOrderVisitor::back_tr stops the visit (top_sort.rs:40-43)
        r.acyclic = false;
        return r;
      } else {
        fr.pos++;
      }
    }
    root = (root == f.start) ? 0 : root + 1;
    while (root < n && col[root] != kWhite) root++;
  }
  return r;
}

}  // namespace

// TopOrderVisitor of top_sort.rs:12-61 for fst_top_sort: false when a back arc is met, otherwise order[state] = position
// of the state in reverse DFS finish order.
bool top_order(const CsrFst& f, std::vector<uint32_t>& order) {
  DfsResult r = dfs_top_order(f);
  if (!r.acyclic) return false;
  order.assign(r.finish.size(), 0);
  for (size_t i = 0; i < r.finish.size(); i++) order[r.finish[r.finish.size() - 1 - i]] = (uint32_t)i;
  return true;
}

QueueKind queue_kind_from_props(uint64_t p, bool has_start) {
  if ((p & props::kTopSorted) || !has_start) return kStateOrderQueue;
  if (p & props::kAcyclic) return kTopOrderQueue;
  if (p & props::kUnweighted) return kLifoQueue;
  return kSccQueue;
}

QueuePlan build_queue_plan(const CsrFst& f, bool defer_acyclic_order) {
  auto t0 = std::chrono::steady_clock::now();
  QueuePlan plan;
  const size_t n = f.num_states();
  const uint64_t p = f.props;
  auto done = [&]() {
    plan.host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return plan;
  };
  if ((p & props::kTopSorted) || !f.has_start) { plan.kind = kStateOrderQueue; return done(); }
  if (p & props::kAcyclic) {
    if (defer_acyclic_order) { plan.kind = kTopOrderQueue; plan.deferred = true; return done(); }
    DfsResult r = dfs_top_order(f);
    if (!r.acyclic) throw FstError("Unexpectted Acyclic FST for TopOprerQueue");  // top_order_queue.rs:25-27 (panic)
    plan.kind = kTopOrderQueue;
    plan.order.assign(n, 0);  // top_sort.rs:52-59: order[finish[len-1-s]] = s
    for (size_t i = 0; i < r.finish.size(); i++) plan.order[r.finish[r.finish.size() - 1 - i]] = (uint32_t)i;
    return done();
  }
  if (p & props::kUnweighted) { plan.kind = kLifoQueue; return done(); }  // TropicalWeight is idempotent
  DfsResult r = dfs(f, true, false);
  std::vector<uint32_t> scc(n);
  for (size_t s = 0; s < n; s++) scc[s] = (uint32_t)(r.nscc - 1 - r.scc[s]);  // scc_visitors.rs:172-179
  // scc_queue_type with compare = None: any arc inside a component makes it a FifoQueue
  std::vector<uint8_t> is_fifo((size_t)r.nscc, 0);
  bool all_trivial = true, unweighted = true;
  for (size_t s = 0; s < n; s++) {
    for (uint32_t i = f.offsets[s]; i < f.offsets[s + 1]; i++) {
      const Tr& tr = f.arcs[i];
      if (scc[s] == scc[tr.nextstate]) { is_fifo[scc[s]] = 1; all_trivial = false; }
      if (!w_is_zero(tr.weight) && !w_is_one(tr.weight)) unweighted = false;
    }
  }
  if (unweighted) { plan.kind = kLifoQueue; return done(); }
  if (all_trivial) { plan.kind = kTopOrderQueue; plan.order = std::move(scc); return done(); }
  plan.kind = kSccQueue;
  plan.scc = std::move(scc);
  plan.scc_is_fifo = std::move(is_fifo);
  return done();
}

}  // namespace b200
