// compose.cu — WFST composition on B200 as a level-synchronous frontier expansion.
//
// Replaces (paths relative to /root/reference):
//   rustfst/src/algorithms/compose/compose_static.rs:198-298   compose_with_config
//   rustfst/src/algorithms/compose/compose_fst_op.rs:169-449   match_type / compute_start / compute_trs /
//                                                              ordered_expand / match_tr / add_tr / final weight
//   rustfst/src/algorithms/compose/matchers/sorted_matcher.rs:124-184  binary search + run of equal labels
//   rustfst/src/algorithms/compose/compose_filters/*.rs        the six epsilon filters
//   rustfst/src/algorithms/lazy/{lazy_fst.rs:226-269,state_table.rs:49-59}  FIFO BFS + tuple->id table
//
// The reference numbers product states in first-emission order under a FIFO BFS.  That order is reproduced
// exactly on a parallel machine: BFS wave k+1 is precisely the set of states first emitted while expanding wave
// k (FIFO), and inside a wave the emission order is (source state id, iterated-arc position, match position).
// Every emitted arc therefore gets a canonical wave-local emission index from two prefix sums; a new tuple's id
// is  first_id_of_next_wave + rank(min emission index over the arcs that reach it).
//
// Per wave (frontier = product ids [lo, hi)):
//   k_state_setup   1 thread / frontier state : which side is iterated, #items (= 1 eps-loop + its arcs), final weight
//   scan            items per state -> item offsets
//   k_item_match    1 thread / item           : binary-search the sorted side, apply the filter -> (pos, count)
//   scan            arcs per item -> canonical emission index of every output arc
//   k_emit          1 thread / output arc     : gather both arcs (128-bit loads), w1 (x) w2, write the 16-byte arc
//                                               coalesced, find-or-insert the destination tuple in the
//                                               open-addressed table, atomicMin its first emission index
//   k_mark_first    1 thread / output arc     : ballot the arcs that are the first emission of a new tuple
//   scan            popcounts of the ballot words (E/32 values)
//   k_resolve       1 thread / output arc     : rank -> state id, patch nextstate, owners publish id + tuple
#include <vector>

#include "compose_common.cuh"

namespace b200 {
namespace composeimpl {

__global__ void k_count_eps(const uint32_t* off, const Tr* arcs, uint32_t n, int by_olabel, uint32_t* neps) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  uint32_t c = 0;
  for (uint32_t i = off[s]; i < off[s + 1]; i++) {
    Label l = by_olabel ? __ldg(&arcs[i].olabel) : __ldg(&arcs[i].ilabel);
    c += (l == kEps);
  }
  neps[s] = c;
}

void launch_count_eps(const uint32_t* off, const Tr* arcs, uint32_t n, int by_olabel, uint32_t* neps, cudaStream_t s) {
  if (n) k_count_eps<<<blocks_for(n), kThreads, 0, s>>>(off, arcs, n, by_olabel, neps);
}

}  // namespace composeimpl

namespace {
using namespace composeimpl;

__global__ void k_init_table(Slot* slots, uint32_t mask, unsigned long long* tuples, unsigned long long key0) {
  uint32_t h = hash_key(key0) & mask;
  slots[h].key = key0; slots[h].id = 0; slots[h].emin = 0;
  tuples[0] = key0;
}

__global__ void k_rehash(const Slot* old_slots, uint32_t old_cap, Slot* slots, uint32_t mask) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= old_cap) return;
  Slot s = old_slots[i];
  if (s.key == kEmptyKey) return;
  uint32_t h = hash_key(s.key) & mask;
  while (true) {
    unsigned long long prev = atomicCAS(&slots[h].key, kEmptyKey, s.key);
    if (prev == kEmptyKey) { slots[h].id = s.id; slots[h].emin = s.emin; return; }
    h = (h + 1) & mask;
  }
}

// compose_fst_op.rs:199-219 match_input, :406-418 compute_trs prologue, :420-449 compute_final_weight
__global__ void k_state_setup(FstView a, FstView b, const unsigned long long* __restrict__ tuples, uint32_t lo,
                              uint32_t F, int match_side, uint32_t* __restrict__ st_items,
                              uint32_t* __restrict__ st_side, float* __restrict__ out_finals) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > F) return;
  if (i == F) { st_items[F] = 0; return; }
  uint32_t fs, s1, s2;
  unpack_key(tuples[lo + i], fs, s1, s2);
  uint32_t d1 = a.off[s1 + 1] - a.off[s1], d2 = b.off[s2 + 1] - b.off[s2];
  bool match_input = match_side == kMatchInput || (match_side == kMatchBoth && d1 <= d2);
  st_items[i] = 1u + (match_input ? d1 : d2);
  st_side[i] = match_input ? 1u : 0u;
  float fw = w_times(a.fin[s1], b.fin[s2]);  // either side non-final (+inf) => +inf => None
  out_finals[lo + i] = w_is_zero(fw) ? w_zero() : fw;
}

// item record: x = first matching arc (absolute index in the searched side), y = run length,
// z = frontier-local state index, w = flags.  The item's position j inside its state (j = 0 is the implicit
// epsilon loop of the iterated side, j >= 1 the (j-1)-th stored arc) is t - item_off[z].
__device__ __forceinline__ uint32_t pack_item_w(bool match_input, bool loop_ok, uint32_t fs_loop, bool real_ok,
                                                uint32_t fs_real) {
  return (loop_ok ? 1u : 0u) | ((fs_loop & 3u) << 1) | (real_ok ? 1u << 3 : 0u) | ((fs_real & 3u) << 4) |
         (match_input ? 1u << 6 : 0u);
}

// ordered_expand / match_tr (compose_fst_op.rs:221-265,324-353) + IteratorSortedMatcher::new (sorted_matcher.rs:124-155)
__global__ void k_item_match(FstView a, FstView b, const unsigned long long* __restrict__ tuples, uint32_t lo,
                             uint32_t F, const uint32_t* __restrict__ item_off, const uint32_t* __restrict__ st_side,
                             uint32_t T, int kind, uint4* __restrict__ item_rec, uint32_t* __restrict__ item_cnt) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > T) return;
  if (t == T) { item_cnt[T] = 0; return; }
  uint32_t i = find_segment(item_off, F, t);
  uint32_t j = t - item_off[i];
  uint32_t fs, s1, s2;
  unpack_key(tuples[lo + i], fs, s1, s2);
  bool match_input = st_side[i] != 0;
  FsFlags ff = state_flags(a, b, s1, s2);

  Label label;  // label of the iterated arc on the matched side
  if (j == 0) label = kNoLabel;  // loop1 = (eps, NO_LABEL) when iterating fst1; loop2 = (NO_LABEL, eps) for fst2
  else label = match_input ? __ldg(&a.arcs[a.off[s1] + j - 1].olabel) : __ldg(&b.arcs[b.off[s2] + j - 1].ilabel);

  bool has_loop = (label == kEps);                    // current_loop
  Label key = (label == kNoLabel) ? kEps : label;     // NO_LABEL matches epsilon arcs, without the loop
  uint32_t pos, end, fs_loop, fs_real;
  if (match_input) {  // search fst2 at s2 by ilabel
    uint32_t blo = b.off[s2], bhi = b.off[s2 + 1];
    pos = has_loop ? blo : lower_bound_label<false>(b.arcs, blo, bhi, key);
    end = run_end<false>(b.arcs, pos, bhi, key);
    fs_loop = has_loop ? filter_eval(kind, fs, ff, label, kNoLabel) : kNoFs;  // (a1, loop2): loop2.ilabel = NO_LABEL
    fs_real = filter_eval(kind, fs, ff, label, key);
  } else {  // search fst1 at s1 by olabel
    uint32_t alo = a.off[s1], ahi = a.off[s1 + 1];
    pos = has_loop ? alo : lower_bound_label<true>(a.arcs, alo, ahi, key);
    end = run_end<true>(a.arcs, pos, ahi, key);
    fs_loop = has_loop ? filter_eval(kind, fs, ff, kNoLabel, label) : kNoFs;  // (loop1, a2): loop1.olabel = NO_LABEL
    fs_real = filter_eval(kind, fs, ff, key, label);
  }
  uint32_t cnt = end - pos;
  bool loop_ok = has_loop && fs_loop != kNoFs;
  bool real_ok = fs_real != kNoFs && cnt > 0;
  item_rec[t] = make_uint4(pos, cnt, i, pack_item_w(match_input, loop_ok, fs_loop, real_ok, fs_real));
  item_cnt[t] = (loop_ok ? 1u : 0u) + (real_ok ? cnt : 0u);
}

__global__ void k_state_offsets(const uint32_t* __restrict__ item_off, const uint32_t* __restrict__ arc_off,
                                uint32_t lo, uint32_t F, uint32_t base, uint32_t* __restrict__ out_offsets) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F) return;
  out_offsets[lo + i] = base + arc_off[item_off[i]];
}

// add_tr (compose_fst_op.rs:267-285) + StateTable::find_id (state_table.rs:49-59,115-118)
__global__ void __launch_bounds__(kThreads)
k_emit(FstView a, FstView b, const unsigned long long* __restrict__ tuples, uint32_t lo,
       const uint32_t* __restrict__ item_off, const uint32_t* __restrict__ arc_off,
       const uint4* __restrict__ item_rec, uint32_t T, uint32_t E,
       Tr* __restrict__ out_arcs /* already offset by base */, Slot* __restrict__ slots, uint32_t mask) {
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  uint32_t t = find_segment(arc_off, T, e);
  uint32_t k = e - __ldg(&arc_off[t]);
  uint4 rec = __ldg(&item_rec[t]);
  uint32_t j = t - __ldg(&item_off[rec.z]);
  bool loop_ok = rec.w & 1u;
  bool match_input = (rec.w >> 6) & 1u;
  uint32_t fs, s1, s2;
  unpack_key(__ldg(&tuples[lo + rec.z]), fs, s1, s2);

  Tr it;
  if (j == 0) {
    it = match_input ? Tr{kEps, kNoLabel, 0.0f, s1} : Tr{kNoLabel, kEps, 0.0f, s2};
  } else {
    it = match_input ? load_tr(&a.arcs[a.off[s1] + j - 1]) : load_tr(&b.arcs[b.off[s2] + j - 1]);
  }
  Tr cand;
  uint32_t fsn;
  if (loop_ok && k == 0) {
    cand = match_input ? Tr{kNoLabel, kEps, 0.0f, s2} : Tr{kEps, kNoLabel, 0.0f, s1};
    fsn = (rec.w >> 1) & 3u;
  } else {
    uint32_t idx = rec.x + k - (loop_ok ? 1u : 0u);
    cand = match_input ? load_tr(&b.arcs[idx]) : load_tr(&a.arcs[idx]);
    fsn = (rec.w >> 4) & 3u;
  }
  const Tr& arc1 = match_input ? it : cand;
  const Tr& arc2 = match_input ? cand : it;
  Tr out;
  out.ilabel = arc1.ilabel;
  out.olabel = arc2.olabel;
  out.weight = w_times(arc1.weight, arc2.weight);

  // find-or-insert (fs', n1, n2)
  unsigned long long key = pack_key(fsn, arc1.nextstate, arc2.nextstate);
  uint32_t h = hash_key(key) & mask;
  while (true) {
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&slots[h].key);
    if (cur == key) break;
    if (cur == kEmptyKey) {
      unsigned long long prev = atomicCAS(&slots[h].key, kEmptyKey, key);
      if (prev == kEmptyKey || prev == key) break;
    }
    h = (h + 1) & mask;
  }
  uint32_t id = *reinterpret_cast<volatile uint32_t*>(&slots[h].id);  // only written between waves
  if (id != kUnassigned) {
    out.nextstate = id;
  } else {
    atomicMin(&slots[h].emin, e);
    out.nextstate = kPendingBit | h;
  }
  store_tr(&out_arcs[e], out);
}

__global__ void k_mark_first(const Tr* __restrict__ out_arcs, const Slot* __restrict__ slots, uint32_t E,
                             uint32_t W, uint32_t* __restrict__ first_bits, uint32_t* __restrict__ first_cnt) {
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  bool owner = false;
  if (e < E) {
    uint32_t ns = __ldg(&out_arcs[e].nextstate);
    if (ns & kPendingBit) owner = (slots[ns & ~kPendingBit].emin == e);
  }
  uint32_t word = __ballot_sync(0xFFFFFFFFu, owner);
  uint32_t w = e >> 5;
  if ((threadIdx.x & 31) == 0) {
    if (w < W) { first_bits[w] = word; first_cnt[w] = __popc(word); }
    else if (w == W) first_cnt[W] = 0;
  }
}

__global__ void k_resolve(Tr* __restrict__ out_arcs, Slot* __restrict__ slots, uint32_t E,
                          const uint32_t* __restrict__ first_bits, const uint32_t* __restrict__ first_pre,
                          uint32_t next_base, unsigned long long* __restrict__ tuples) {
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  uint32_t ns = out_arcs[e].nextstate;
  if (!(ns & kPendingBit)) return;
  uint32_t h = ns & ~kPendingBit;
  uint32_t emin = slots[h].emin;
  uint32_t w = emin >> 5;
  uint32_t rank = __ldg(&first_pre[w]) + __popc(__ldg(&first_bits[w]) & ((1u << (emin & 31)) - 1u));
  uint32_t id = next_base + rank;
  out_arcs[e].nextstate = id;
  if (emin == e) {  // the first emission publishes the id; nobody reads slot.id inside this kernel
    slots[h].id = id;
    tuples[id] = slots[h].key;
  }
}

struct EventPair { cudaEvent_t a, b; };

}  // namespace

DevFst compose_device_waves(const DevFst& fa, const DevFst& fb, const ComposeOptions& opt, ComposeStats* stats,
                            cudaStream_t s) {
  int kind = opt.filter == kAutoFilter ? kSequenceFilter : opt.filter;  // compose_fst.rs:58-92
  if (kind < kNullFilter || kind > kNoMatchFilter) throw FstError("EnumConversionError");

  if (opt.sigma1.enabled || opt.sigma2.enabled)
    throw FstError("sigma matcher configurations are only supported by the persistent compose back end "
                   "(the result exceeded its pre-sized buffers)");
  int side = resolve_match_side(fa.props, fb.props);
  if (fa.num_states >= 0x7FFFFFFFu || fb.num_states >= 0x7FFFFFFFu)
    throw FstError("compose: operands with >= 2^31 states are not supported");

  ComposeStats local;
  ComposeStats& st = stats ? *stats : local;
  st = ComposeStats();
  cudaEvent_t ev0, ev1, ev2;
  B200_CUDA(cudaEventCreate(&ev0)); B200_CUDA(cudaEventCreate(&ev1)); B200_CUDA(cudaEventCreate(&ev2));
  std::vector<EventPair> emit_events;
  B200_CUDA(cudaEventRecord(ev0, s));

  DevFst out(s);
  out.props = props::of_compose(fa.props, fb.props);  // mutate_properties.rs:151-184, lazy_fst.rs:260

  // compute_start (compose_fst_op.rs:389-404): no start on either side => empty FST
  if (!fa.has_start || !fb.has_start) {
    out.offsets.reserve_discard(1);
    B200_CUDA(cudaMemsetAsync(out.offsets.p, 0, 4, s));
    out.props = props::kNull;  // VectorFst::new(); lazy_fst.rs:229-232 returns before set_properties
    if (opt.connect) out.props = props::after_connect(out.props);
    B200_CUDA(cudaEventRecord(ev1, s)); B200_CUDA(cudaEventRecord(ev2, s));
    B200_CUDA(cudaStreamSynchronize(s));
    cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev2);
    return out;
  }

  // ---- per-state epsilon counts on the matched sides (only the epsilon-aware filters read them)
  DevBuf<uint32_t> neps1(s), neps2(s);
  bool need_eps = (kind == kSequenceFilter || kind == kAltSequenceFilter || kind == kMatchFilter);
  FstView va{fa.offsets.p, fa.arcs.p, fa.finals.p, nullptr, fa.num_states};
  FstView vb{fb.offsets.p, fb.arcs.p, fb.finals.p, nullptr, fb.num_states};
  if (need_eps && !(fa.props & props::kNoOEpsilons)) {
    neps1.reserve_discard(fa.num_states);
    launch_count_eps(va.off, va.arcs, va.n, 1, neps1.p, s);
    va.neps = neps1.p; st.kernel_launches++;
  }
  if (need_eps && !(fb.props & props::kNoIEpsilons)) {
    neps2.reserve_discard(fb.num_states);
    launch_count_eps(vb.off, vb.arcs, vb.n, 0, neps2.p, s);
    vb.neps = neps2.p; st.kernel_launches++;
  }

  // ---- state table + output arrays (sized for 180 GB of HBM: start generous, double on demand)
  size_t guess_states = (size_t)fa.num_states + fb.num_states + 1024;
  uint32_t table_cap = 1u << 16;
  while (table_cap < 4 * guess_states && table_cap < (1u << 30)) table_cap <<= 1;
  DevBuf<Slot> slots(s, table_cap);
  B200_CUDA(cudaMemsetAsync(slots.p, 0xFF, (size_t)table_cap * sizeof(Slot), s));
  uint32_t table_mask = table_cap - 1;

  DevBuf<unsigned long long> tuples(s, guess_states);
  DevBuf<uint32_t> out_offsets(s, guess_states + 1);
  DevBuf<float> out_finals(s, guess_states);
  DevBuf<Tr> out_arcs(s, (size_t)fa.num_arcs + fb.num_arcs + 1024);
  auto ensure_states = [&](size_t n, size_t keep) {
    tuples.reserve_keep(n, keep);
    out_offsets.reserve_keep(n + 1, keep + 1);
    out_finals.reserve_keep(n, keep);
  };

  uint32_t start_fs = (kind == kNullFilter || kind == kTrivialFilter || kind == kNoMatchFilter) ? 1u : 0u;
  k_init_table<<<1, 1, 0, s>>>(slots.p, table_mask, tuples.p, pack_key(start_fs, fa.start, fb.start));
  st.kernel_launches++;

  DevBuf<uint32_t> st_items(s), st_side(s), item_cnt(s), first_bits(s), first_cnt(s);
  DevBuf<uint4> item_rec(s);
  DevBuf<uint8_t> scan_tmp(s);

  uint32_t lo = 0, hi = 1;  // frontier = product ids [lo, hi)
  uint64_t total_arcs = 0;
  while (lo < hi) {
    uint32_t F = hi - lo;
    st.waves++;
    // (1) per-state setup + item offsets
    st_items.reserve_discard((size_t)F + 1); st_side.reserve_discard((size_t)F + 1);
    k_state_setup<<<blocks_for((size_t)F + 1), kThreads, 0, s>>>(va, vb, tuples.p, lo, F, side, st_items.p,
                                                                 st_side.p, out_finals.p);
    exclusive_sum_u32(st_items.p, st_items.p, (size_t)F + 1, scan_tmp, s);  // st_items becomes item_off
    uint32_t T = read_u32(st_items.p + F, s);
    // (2) per-item matching + canonical emission offsets
    item_rec.reserve_discard(T); item_cnt.reserve_discard((size_t)T + 1);
    k_item_match<<<blocks_for((size_t)T + 1), kThreads, 0, s>>>(va, vb, tuples.p, lo, F, st_items.p, st_side.p, T,
                                                                kind, item_rec.p, item_cnt.p);
    exclusive_sum_u32(item_cnt.p, item_cnt.p, (size_t)T + 1, scan_tmp, s);  // item_cnt becomes arc_off
    uint32_t E = read_u32(item_cnt.p + T, s);
    st.kernel_launches += 4;
    st.states_expanded += F;
    st.arcs_iterated += (uint64_t)T - F;
    st.arcs_emitted += E;
    if (total_arcs + E > 0xFFFFFFF0ull) throw FstError("compose: result has more than 2^32 transitions");
    uint32_t base = (uint32_t)total_arcs;
    k_state_offsets<<<blocks_for(F), kThreads, 0, s>>>(st_items.p, item_cnt.p, lo, F, base, out_offsets.p);
    st.kernel_launches++;
    uint32_t n_new = 0;
    if (E > 0) {
      // capacity: arcs, and a table load factor <= 1/2 even if every arc discovers a new state
      out_arcs.reserve_keep((size_t)base + E, base);
      if (((size_t)hi + E) * 2 > table_cap) {
        uint32_t new_cap = table_cap;
        while (((size_t)hi + E) * 4 > new_cap) {
          if (new_cap >= (1u << 30)) throw FstError("compose: state table would exceed 2^30 slots");
          new_cap <<= 1;
        }
        DevBuf<Slot> bigger(s, new_cap);
        B200_CUDA(cudaMemsetAsync(bigger.p, 0xFF, (size_t)new_cap * sizeof(Slot), s));
        k_rehash<<<blocks_for(table_cap), kThreads, 0, s>>>(slots.p, table_cap, bigger.p, new_cap - 1);
        st.kernel_launches++;
        slots = std::move(bigger);
        table_cap = new_cap; table_mask = new_cap - 1;
      }
      // (3) emit
      EventPair ep;
      B200_CUDA(cudaEventCreate(&ep.a)); B200_CUDA(cudaEventCreate(&ep.b));
      B200_CUDA(cudaEventRecord(ep.a, s));
      k_emit<<<blocks_for(E), kThreads, 0, s>>>(va, vb, tuples.p, lo, st_items.p, item_cnt.p, item_rec.p, T, E,
                                                out_arcs.p + base, slots.p, table_mask);
      B200_CUDA(cudaEventRecord(ep.b, s));
      emit_events.push_back(ep);
      st.emit_launches++;
      // (4) rank first emissions -> ids of the next wave
      uint32_t W = (E + 31) / 32;
      first_bits.reserve_discard(W); first_cnt.reserve_discard((size_t)W + 1);
      k_mark_first<<<blocks_for(((size_t)W + 1) * 32), kThreads, 0, s>>>(out_arcs.p + base, slots.p, E, W,
                                                                         first_bits.p, first_cnt.p);
      exclusive_sum_u32(first_cnt.p, first_cnt.p, (size_t)W + 1, scan_tmp, s);
      n_new = read_u32(first_cnt.p + W, s);
      if ((uint64_t)hi + n_new >= 0x7FFFFFFFull) throw FstError("compose: result has >= 2^31 states");
      ensure_states((size_t)hi + n_new, hi);
      k_resolve<<<blocks_for(E), kThreads, 0, s>>>(out_arcs.p + base, slots.p, E, first_bits.p, first_cnt.p, hi,
                                                   tuples.p);
      st.kernel_launches += 4;
    }
    total_arcs += E;
    lo = hi;
    hi += n_new;
  }
  uint32_t S = hi;
  uint32_t A = (uint32_t)total_arcs;
  B200_CUDA(cudaMemcpyAsync(out_offsets.p + S, &A, 4, cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaStreamSynchronize(s));  // A lives on the stack

  out.offsets = std::move(out_offsets);
  out.arcs = std::move(out_arcs);
  out.finals = std::move(out_finals);
  out.num_states = S; out.num_arcs = A;
  out.has_start = true; out.start = 0;
  B200_CUDA(cudaEventRecord(ev1, s));

  if (opt.connect) {
    uint64_t launches = 0;
    DevFst trimmed = connect_device(out, /*assume_accessible=*/true, &launches, s);
    st.kernel_launches += launches;
    out = std::move(trimmed);
  }
  st.states_out = out.num_states; st.arcs_out = out.num_arcs;
  B200_CUDA(cudaEventRecord(ev2, s));
  B200_CUDA(cudaStreamSynchronize(s));
  B200_CUDA(cudaEventElapsedTime(&st.ms_expand, ev0, ev1));
  B200_CUDA(cudaEventElapsedTime(&st.ms_connect, ev1, ev2));
  for (auto& ep : emit_events) {
    float ms = 0;
    B200_CUDA(cudaEventElapsedTime(&ms, ep.a, ep.b));
    st.ms_emit_kernel += ms;
    cudaEventDestroy(ep.a); cudaEventDestroy(ep.b);
  }
  cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev2);
  return out;
}

}  // namespace b200
