// algos.h — device algorithms behind fst_compose* / fst_connect / fst_shortest_path* (see include/rustfst_b200.h).
#pragma once
#include <vector>

#include "device_common.cuh"

namespace b200 {

// rustfst-ffi/src/algorithms/compose.rs:20-33 (values of the size_t passed to fst_compose_config_new)
enum ComposeFilter : int {
  kAutoFilter = 0, kNullFilter = 1, kTrivialFilter = 2, kSequenceFilter = 3, kAltSequenceFilter = 4,
  kMatchFilter = 5, kNoMatchFilter = 6
};

// rustfst/src/algorithms/compose/matchers/sigma_matcher.rs + compose_static.rs:34-58 (MatcherConfig)
struct SigmaSpec {
  bool enabled = false;
  uint32_t sigma_label = kNoLabel;
  int rewrite_mode = 0;            // 0 Auto (rewrite both labels iff the FST is an acceptor), 1 Always, 2 Never
  std::vector<uint32_t> allowed;   // empty = every label may match sigma
};

struct ComposeOptions {  // rustfst/src/algorithms/compose/compose_static.rs:80-97
  int filter = kAutoFilter;
  bool connect = true;
  SigmaSpec sigma1, sigma2;        // matcher1_config (fst1, MatchOutput) / matcher2_config (fst2, MatchInput)
};

struct ComposeStats {
  uint64_t states_expanded = 0;  // S   product states expanded
  uint64_t arcs_iterated = 0;    // A_it arcs scanned on the iterated side (excludes the implicit eps loops)
  uint64_t arcs_emitted = 0;     // A_out arcs emitted before trimming
  uint64_t waves = 0;            // BFS levels
  uint64_t states_out = 0, arcs_out = 0;  // after connect (== S, A_out when connect is off)
  float ms_expand = 0, ms_connect = 0;    // device time (CUDA events on the call's stream)
  float ms_emit_kernel = 0;               // summed duration of the emit kernel (roofline numerator's time)
  uint64_t emit_launches = 0;
  uint64_t kernel_launches = 0;           // kernels of this library launched by the call
  float ms_phase[4] = {0, 0, 0, 0};       // persistent back end: device time inside phases A (match), B (emit),
                                          // C (rank), D (resolve), from %globaltimer
};

// a must be olabel-sorted and/or b ilabel-sorted as recorded in their property words (A.1 of SURVEY.md);
// throws FstError with the reference's message otherwise.
DevFst compose_device(const DevFst& a, const DevFst& b, const ComposeOptions& opt, ComposeStats* stats,
                      cudaStream_t s);
// Back end 1: one kernel per phase and wave, launch sizes read back by the host (grows every buffer on demand).
DevFst compose_device_waves(const DevFst& a, const DevFst& b, const ComposeOptions& opt, ComposeStats* stats,
                            cudaStream_t s);
// Back end 2: one persistent cooperative kernel runs the whole BFS (no host round trips); returns false when a
// pre-sized buffer overflowed, in which case the caller falls back to back end 1.
// Batched mode: `a` is the disjoint union of many acceptors; the BFS starts from the n tuples (starts1[i], start(b))
// (product ids 0..n-1) and the result is the union of the individual compositions.  out_s1[id] = fst1 state of every
// result state (identifies its acceptor), out_start_map[i] = result id of start tuple i or 0xFFFFFFFF if trimmed.
struct BatchStarts {
  const uint32_t* d_starts1 = nullptr;
  uint32_t n = 0;
  DevBuf<uint32_t>* out_s1 = nullptr;
  DevBuf<uint32_t>* out_start_map = nullptr;
};
bool compose_device_coop(const DevFst& a, const DevFst& b, const ComposeOptions& opt, ComposeStats* stats,
                         cudaStream_t s, DevFst* out, const BatchStarts* batch = nullptr);

// Back end 3 (default): the persistent warp-stream kernel of compose_ws.cu — one exchange and one grid barrier per BFS
// wave.  Returns 0 on success, otherwise the overflow flags (compose_match.cuh) of the capacity that was too small;
// `caps` carries the capacities of the attempt in and out (0 = derive from the operands).
struct WsCaps { size_t states = 0, arcs = 0, items = 0, runs = 0, waves = 0; };
int compose_device_ws(const DevFst& a, const DevFst& b, const ComposeOptions& opt, ComposeStats* stats, cudaStream_t s,
                      DevFst* out, const BatchStarts* batch = nullptr, WsCaps* caps = nullptr);
// The persistent back ends with growth on overflow (false = not representable, use back end 1).
bool compose_device_persistent(const DevFst& a, const DevFst& b, const ComposeOptions& opt, ComposeStats* stats,
                               cudaStream_t s, DevFst* out, const BatchStarts* batch = nullptr);

// The n results of a batched composition in one block (no per-result allocation): result i owns the states
// [state_off[i], state_off[i+1]) and the arcs [arc_off[i], arc_off[i+1]); `offsets` are per-state arc offsets into the
// block's arc array, arc next states and `starts` are local to their result (-1 = no start state).
struct PackedBatch {
  uint64_t n = 0;
  std::vector<uint32_t> state_off, arc_off;
  std::vector<int32_t> starts;
  std::vector<uint64_t> props;
  PoolVec<uint32_t> offsets;
  PoolVec<float> finals;
  PoolVec<Tr> arcs;
  size_t byte_size() const;
  void serialize(uint8_t* dst) const;                              // byte_size() bytes
  static PackedBatch deserialize(const uint8_t* src, size_t len);  // validates every index
  CsrFst result(size_t i) const;
  void append(const CsrFst& c);
};
// Stable partition of a batched composition result by component, on the device (batch.cu).
void split_batch_device(const DevFst& r, const uint32_t* d_tag, const uint32_t* d_start_map,
                        const std::vector<uint32_t>& base_state, PackedBatch& out, uint64_t* launches, cudaStream_t s);

// Trim: keep states that are accessible and coaccessible, order-preserving renumbering
// (rustfst/src/algorithms/connect.rs:51-66, rustfst/src/fst_impls/vector_fst/mutable_fst.rs:132-189).
// assume_accessible skips the forward pass (true for a freshly composed FST: every state was reached by the BFS).
DevFst connect_device(const DevFst& in, bool assume_accessible, uint64_t* launches, cudaStream_t s);
// Same result for a freshly composed FST whose BFS wave boundaries are known (wave k = ids [wave_lo[k], wave_lo[k+1]),
// device array of n_waves + 1 entries, start state = 0): one persistent kernel, coaccessibility pulled in reverse
// wave order.
struct TrimExtras {  // optional per-state payload carried through the compaction (batched mode)
  const unsigned long long* tuples = nullptr;  // packed (s1, s2, fs) of every input state
  uint32_t n_starts = 1;
  DevBuf<uint32_t>* out_tag = nullptr;         // s1 of every KEPT state, in new-id order
  DevBuf<uint32_t>* out_start_map = nullptr;   // new id of input states 0..n_starts-1 (0xFFFFFFFF = deleted)
};
// `pa` (optional): the arcs of `in` are not in in.arcs but still in the provisional runs of compose_ws.cu; in.offsets,
// in.finals and in.num_arcs describe the canonical machine, pa->next its resolved next states in canonical order.
struct ProvArcs {
  const Tr* prov = nullptr;
  const uint2* st_first = nullptr;   // per state: (run, index in the run) of its first arc
  const uint32_t* run_src = nullptr; // first provisional arc of every run
  const uint32_t* run_cnt = nullptr; // arcs of every run
  const uint32_t* run_dst = nullptr; // canonical index of every run's first arc (n_runs + 1 entries)
  uint32_t n_runs = 0;
  const uint32_t* next = nullptr;    // num_arcs resolved next states, canonical order
};
DevFst connect_waves_device(const DevFst& in, const uint32_t* d_wave_lo, uint32_t n_waves, uint64_t* launches,
                            cudaStream_t s, const TrimExtras* extras = nullptr, const ProvArcs* pa = nullptr);

// isomorphic() on the device (iso.cu; isomorphic.rs:49-160 with delta = KDELTA): 1 = isomorphic, 0 = not, -1 = undecided
// (the reference's answer would depend on its visiting order: take the host restatement, host_fst.h isomorphic()).
int isomorphic_device(const DevFst& a, const DevFst& b, float delta, cudaStream_t s);

// Stable per-state arc sort by input or output label, in place on the device (algorithms/tr_sort.rs:51-62).
void tr_sort_device(DevFst& f, bool ilabel, cudaStream_t s);

// ---- shortest path (n = 1)
enum QueueKind : int { kStateOrderQueue = 0, kTopOrderQueue = 1, kLifoQueue = 2, kSccQueue = 3 };

// Host-built mirror of AutoQueue::new (rustfst/src/algorithms/queues/auto_queue.rs:23-99): which discipline the
// reference would pick from the stored property bits, plus the DFS-derived orders it needs.
struct QueuePlan {
  QueueKind kind = kStateOrderQueue;
  std::vector<uint32_t> order;      // kTopOrderQueue: order[state] (reverse DFS finish order or reversed SCC number)
  std::vector<uint32_t> scc;        // kSccQueue: scc[state]
  std::vector<uint8_t> scc_is_fifo; // kSccQueue: per component, 1 = FifoQueue, 0 = TrivialQueue
  double host_ms = 0;               // time spent on the host DFS
  bool deferred = false;            // kTopOrderQueue of an ACYCLIC machine whose order has not been computed yet
  const uint32_t* d_order = nullptr; // kTopOrderQueue: the order, already on the device (dag_order.cu); else `order`
  float device_ms = 0;              // device time spent computing d_order
};
// defer_acyclic_order: for a machine whose properties say ACYCLIC, only record the decision (the caller computes the
// order on the device with dag_top_order_device and falls back to the full host plan if that declines).
QueuePlan build_queue_plan(const CsrFst& fst, bool defer_acyclic_order = false);
// The decision alone, from a property word (auto_queue.rs:28-45,68-76); kSccQueue = "needs the host DFS".
QueueKind queue_kind_from_props(uint64_t props, bool has_start);
// order[s] = position of s in the reverse finish order of the reference's DFS (top_sort.rs:12-61) for an acyclic
// machine, computed on the device.  false = cyclic, or deeper than the device path handles: use the host DFS.
bool dag_top_order_device(const DevFst& f, DevBuf<uint32_t>& order, float* ms, uint64_t* launches, cudaStream_t s);
// Topological order of an acyclic machine as the reference's TopOrderVisitor numbers it (false = cyclic).
bool top_order(const CsrFst& fst, std::vector<uint32_t>& order);

struct SsspStats {
  uint64_t arcs_relaxed = 0;   // E
  uint64_t states_settled = 0; // N touched
  uint64_t waves = 0;
  int path = 0;                // 0 = parallel relaxation path, 1 = order-faithful serial kernel
  float ms_device = 0;         // device time of the whole call
  float ms_relax_kernel = 0;   // summed duration of the relaxation kernel
  uint64_t relax_launches = 0;
  uint64_t kernel_launches = 0;
  double plan_host_ms = 0;
  float order_device_ms = 0;   // device time of the TopOrderQueue order (dag_order.cu)
  bool order_on_device = false;
  bool sweep = false;          // the waves went over their visit budget: distances came from the in-order sweep
};

// Returns the single-shortest-path FST exactly as rustfst's single_shortest_path + backtrace would
// (rustfst/src/algorithms/shortest_path.rs:173-282): state 0 = final-most state, start = last state.
CsrFst shortest_path_device(const DevFst& fst, const QueuePlan& plan, SsspStats* stats, cudaStream_t s,
                            bool force_serial = false);

// Forward shortest distances from the start state with the semantics of shortest_distance_with_config(fst, false,
// delta) (rustfst/src/algorithms/shortest_distance.rs:153-237,312-323): out[s] for s < num_states, +inf = unreached
// (the reference's vector may be shorter; missing entries read as zero() everywhere it is used).
void shortest_distance_device(const DevFst& fst, const QueuePlan& plan, float delta, DevBuf<float>& out,
                              SsspStats* stats, cudaStream_t s, bool force_serial = false);

// reverse() of rustfst/src/algorithms/reverse.rs:33-87 as a device CSR: state 0 = superinitial with one epsilon arc
// per final state (in state order, weight = final weight), state t + 1 = in-arcs of t in (source state, arc
// position) order with nextstate = source + 1; final: start + 1 with weight one().
DevFst reverse_device(const DevFst& fst, cudaStream_t s, uint64_t* launches = nullptr);
// The same as a host FST with the reference's property word (fst_reverse).
CsrFst reverse_fst_device(const DevFst& fst, cudaStream_t s);

// n > 1 shortest paths (shortest_path.rs:135-170, n_shortest_path :409-518; unique = true walks the reversed machine
// determinized on demand, :156-165): distances and the
// reversed machine are built on the device, the n-best heap search runs on the host over rows fetched on demand,
// the result tree is trimmed on the device.  (Machines with `Some(+inf)` final weights are refused by upload().)
struct NShortestStats {
  SsspStats distance;
  uint64_t heap_pops = 0, rows_fetched = 0, rows_cached = 0, arcs_fetched = 0, states_before_trim = 0;
  uint64_t det_states_expanded = 0;  // unique = true: subset states whose arcs were built
  float ms_distance = 0, ms_reverse = 0, ms_search_host = 0, ms_total = 0;
};
CsrFst n_shortest_paths_device(const DevFst& fst, const QueuePlan& plan,
                               size_t nshortest, float delta, NShortestStats* stats, cudaStream_t s,
                               bool force_serial = false, bool unique = false);

}  // namespace b200
