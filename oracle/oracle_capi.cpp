// oracle/oracle_capi.cpp — TEST INFRASTRUCTURE ONLY: a flat C API over oracle.hpp so that tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg can drive the CPU oracle via
// ctypes.  The product library (rustfst_b200/csrc) never links or calls this.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include "oracle.hpp"

using namespace oracle;
static thread_local std::string g_err;

#define GUARD(expr)                         \
  try { expr; return 0; }                   \
  catch (const std::exception& e) { g_err = e.what(); return 1; }

extern "C" {
const char* oracle_last_error() { return g_err.c_str(); }
void* oracle_fst_new() { return new Fst(); }
void oracle_fst_free(void* f) { delete (Fst*)f; }
int oracle_fst_add_state(void* f, uint32_t* out) { GUARD(*out = ((Fst*)f)->add_state()); }
int oracle_fst_set_start(void* f, uint32_t s) { GUARD(((Fst*)f)->set_start(s)); }
int oracle_fst_set_final(void* f, uint32_t s, float w) { GUARD(((Fst*)f)->set_final(s, w)); }
int oracle_fst_add_tr(void* f, uint32_t s, uint32_t il, uint32_t ol, float w, uint32_t ns) {
  GUARD(((Fst*)f)->add_tr(s, Tr{il, ol, w, ns}));
}
int oracle_fst_tr_sort(void* f, int ilabel) { GUARD(tr_sort(*(Fst*)f, ilabel != 0)); }
int oracle_fst_connect(void* f) { GUARD(connect(*(Fst*)f)); }
int oracle_fst_top_sort(void* f) { GUARD(top_sort(*(Fst*)f)); }
int oracle_fst_reverse(void* f, void** out) { GUARD(*out = new Fst(reverse(*(Fst*)f))); }
int oracle_fst_compute_props(void* f) { GUARD(((Fst*)f)->props = compute_fst_properties_all(*(Fst*)f)); }
uint64_t oracle_fst_props(void* f) { return ((Fst*)f)->props; }
void oracle_fst_set_props(void* f, uint64_t p) { ((Fst*)f)->props = p; }
uint64_t oracle_fst_num_states(void* f) { return ((Fst*)f)->num_states(); }
uint64_t oracle_fst_num_trs(void* f) { return ((Fst*)f)->num_trs_total(); }
int64_t oracle_fst_start(void* f) { return ((Fst*)f)->has_start ? (int64_t)((Fst*)f)->start : -1; }
int oracle_fst_equal(void* a, void* b) { return fst_equal(*(Fst*)a, *(Fst*)b) ? 1 : 0; }

int oracle_fst_from_bytes(const uint8_t* data, uint64_t len, void** out) {
  GUARD(*out = new Fst(fst_from_bytes(data, (size_t)len, true)));
}
// Two-call protocol: returns the size; copies when buf != NULL and cap >= size.
int oracle_fst_to_bytes(void* f, uint8_t* buf, uint64_t cap, uint64_t* size) {
  GUARD({
    auto b = fst_to_bytes(*(Fst*)f);
    *size = b.size();
    if (buf && cap >= b.size()) std::memcpy(buf, b.data(), b.size());
  });
}
// CSR export for bulk comparisons: offsets[N+1], arcs[A] (16-byte Tr), finals[N] (+inf = non final)
int oracle_fst_to_csr(void* fp, uint64_t* offsets, Tr* arcs, float* finals) {
  GUARD({
    Fst& f = *(Fst*)fp;
    uint64_t o = 0;
    for (size_t s = 0; s < f.states.size(); s++) {
      offsets[s] = o;
      finals[s] = f.states[s].has_final ? f.states[s].final_weight : W_ZERO;
      for (auto& t : f.states[s].trs) arcs[o++] = t;
    }
    offsets[f.states.size()] = o;
  });
}
// Bulk build from CSR (props taken verbatim, like a header read).
int oracle_fst_from_csr(uint64_t nstates, const uint64_t* offsets, const Tr* arcs, const float* finals,
                        int64_t start, uint64_t props, void** out) {
  GUARD({
    auto* f = new Fst();
    f->states.resize(nstates);
    for (uint64_t s = 0; s < nstates; s++) {
      State& st = f->states[s];
      if (!w_eq(finals[s], W_ZERO)) { st.has_final = true; st.final_weight = finals[s]; }
      st.trs.assign(arcs + offsets[s], arcs + offsets[s + 1]);
      for (auto& t : st.trs) { if (t.ilabel == 0) st.niepsilons++; if (t.olabel == 0) st.noepsilons++; }
    }
    f->has_start = start >= 0; f->start = (StateId)start; f->props = props & P::TRINARY;
    *out = f;
  });
}

// sigma: optional 2 x {enabled, sigma_label, rewrite_mode, n_allowed, allowed...} flattened; NULL = no sigma matcher
static void fill_sigma(SigmaConfig& sc, const uint32_t*& p) {
  sc.enabled = *p++ != 0;
  sc.sigma_label = *p++;
  sc.rewrite_mode = (int)*p++;
  uint32_t n = *p++;
  sc.has_allowed = n > 0;
  for (uint32_t i = 0; i < n; i++) sc.allowed.insert(*p++);
}
// stats: [states_expanded, arcs_iterated, arcs_emitted]; seconds: wall time of the algorithm only
int oracle_compose_sigma(void* a, void* b, int filter, int do_connect, const uint32_t* sigma, void** out) {
  GUARD({
    ComposeConfig cfg; cfg.filter = filter; cfg.connect = do_connect != 0;
    const uint32_t* p = sigma;
    fill_sigma(cfg.sigma1, p);
    fill_sigma(cfg.sigma2, p);
    *out = new Fst(compose(*(Fst*)a, *(Fst*)b, cfg, nullptr));
  });
}
// number of successful paths of an acyclic FST (fst_traits/paths_iterator.rs semantics: start -> any final state)
int oracle_count_paths(void* fp, uint64_t* count) {
  try {
    Fst& f = *(Fst*)fp;
    *count = 0;
    if (!f.has_start) return 0;
    std::vector<uint64_t> memo(f.num_states(), ~0ull);
    std::vector<std::pair<StateId, size_t>> stack;
    std::vector<uint8_t> onstack(f.num_states(), 0);
    // iterative post-order DP; throws on cycles
    stack.push_back({f.start, 0});
    onstack[f.start] = 1;
    std::vector<uint64_t> acc(f.num_states(), 0);
    while (!stack.empty()) {
      auto& fr = stack.back();
      StateId s = fr.first;
      if (fr.second < f.states[s].trs.size()) {
        StateId t = f.states[s].trs[fr.second].nextstate;
        if (memo[t] != ~0ull) { acc[s] += memo[t]; fr.second++; }
        else if (onstack[t]) throw std::runtime_error("cyclic");
        else { onstack[t] = 1; stack.push_back({t, 0}); }
      } else {
        memo[s] = acc[s] + (f.states[s].has_final ? 1 : 0);
        onstack[s] = 0;
        stack.pop_back();
        if (!stack.empty()) { acc[stack.back().first] += memo[s]; stack.back().second++; }
      }
    }
    *count = memo[f.start];
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
int oracle_compose(void* a, void* b, int filter, int do_connect, void** out, uint64_t* stats, double* seconds) {
  GUARD({
    ComposeConfig cfg; cfg.filter = filter; cfg.connect = do_connect != 0;
    ComposeStats st;
    auto t0 = std::chrono::steady_clock::now();
    Fst r = compose(*(Fst*)a, *(Fst*)b, cfg, &st);
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    if (stats) { stats[0] = st.states_expanded; stats[1] = st.arcs_iterated; stats[2] = st.arcs_emitted; }
    *out = new Fst(std::move(r));
  });
}
// The queue AutoQueue::new builds for `a` (auto_queue.rs:23-99): kind 0 StateOrder, 1 TopOrder, 2 Lifo, 3 Scc;
// order_or_scc[n] = order[state] (TopOrder) or scc[state] (Scc); is_fifo[n_scc].
int oracle_queue_plan(void* a, int32_t* kind, uint32_t* order_or_scc, uint8_t* is_fifo, uint32_t* n_scc) {
  GUARD({
    Fst& f = *(Fst*)a;
    QueueKind k;
    std::unique_ptr<Queue> q = make_auto_queue(f, &k);
    *kind = (int32_t)k; *n_scc = 0;
    if (k == QK_TOP_ORDER) {
      auto* t = dynamic_cast<TopOrderQueue*>(q.get());
      for (size_t i = 0; i < t->order.size(); i++) order_or_scc[i] = t->order[i];
    } else if (k == QK_SCC) {
      auto* sq = dynamic_cast<SccQueue*>(q.get());
      for (size_t i = 0; i < sq->sccs.size(); i++) order_or_scc[i] = sq->sccs[i];
      *n_scc = (uint32_t)sq->queues.size();
      for (size_t c = 0; c < sq->queues.size(); c++) is_fifo[c] = dynamic_cast<FifoQueue*>(sq->queues[c].get()) ? 1 : 0;
    }
  });
}
// stats: [arcs_relaxed, states_dequeued]; distance (optional, N floats)
int oracle_shortest_path(void* a, uint64_t nshortest, int unique, float delta, void** out, uint64_t* stats,
                         double* seconds, float* distance) {
  GUARD({
    ShortestPathConfig cfg; cfg.nshortest = (size_t)nshortest; cfg.unique = unique != 0; cfg.delta = delta;
    SsspStats st;
    std::vector<float> dist;
    auto t0 = std::chrono::steady_clock::now();
    Fst r = shortest_path(*(Fst*)a, cfg, &st, distance ? &dist : nullptr);
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    if (stats) { stats[0] = st.arcs_relaxed; stats[1] = st.states_dequeued; }
    if (distance) std::memcpy(distance, dist.data(), dist.size() * sizeof(float));
    *out = new Fst(std::move(r));
  });
}
}
