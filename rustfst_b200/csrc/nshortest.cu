// nshortest.cu — n > 1 shortest paths over the tropical semiring on B200.
//
// Replaces (paths relative to /root/reference):
//   rustfst/src/algorithms/shortest_path.rs:135-170   the nshortest > 1 branch of shortest_path_with_config
//   rustfst/src/algorithms/reverse.rs:33-87           reverse (superinitial state, stable in-arc order)
//   rustfst/src/algorithms/shortest_path.rs:288-518   ShortestPathCompare, Heap, n_shortest_path
//
// Work split.  The two passes over the whole machine are data parallel and run on the device:
//   * forward distances (sssp.cu: shortest_distance_device);
//   * the reversed machine: ONE stable radix sort of (reversed-row id) over a list holding first the N "final weight"
//     pseudo-arcs and then the A arcs in CSR order, followed by one gather.  A stable sort keeps CSR order inside a
//     row, which is exactly the (source state, arc position) push order of reverse.rs:62-71, and puts the
//     superinitial state's arcs (row 0) in state order (reverse.rs:57-60).
// The n-best search itself is a best-first search driven by a binary heap whose pop order (ties included) decides
// the numbering of the result states: inherently sequential, O(n * path length * degree) steps that touch a few
// thousand arcs.  It runs on the host (as SURVEY.md §8f ranks it) over rows of the reversed machine fetched on
// demand from HBM — the reversed machine never leaves the device.  The search tree is a host structure (one arc per
// state, all chains ending in the final state), so its `connect` is a walk along the n result chains.
// With unique = true (shortest_path.rs:156-165) the search walks the determinized reversed machine, whose subset states
// are built when the search pops them (see n_shortest_paths_device).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <unordered_map>
#include <vector>

#include "algos.h"

namespace b200 {
namespace {

// keys[0..n)   : row of the final-weight pseudo arc of state s (0 = superinitial row, n + 1 = "none" sentinel row)
// keys[n..n+A) : row of arc e = nextstate + 1
// vals[i] = i; src_of[e] = source state of arc e; rows[r] += 1 (row sizes; rows has n + 2 entries, zeroed)
__global__ void __launch_bounds__(kThreads)
k_rev_keys(const uint32_t* __restrict__ off, const Tr* __restrict__ arcs, const float* __restrict__ fin, uint32_t n,
           unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ src_of,
           uint32_t* __restrict__ rows) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  const bool is_final = s < n && fin[s] != w_zero();
  const uint32_t m = __ballot_sync(0xFFFFFFFFu, is_final);
  if (m && (threadIdx.x & 31) == (uint32_t)(__ffs(m) - 1)) atomicAdd(&rows[0], (uint32_t)__popc(m));
  if (s >= n) return;
  keys[s] = is_final ? 0ull : (unsigned long long)n + 1ull;
  vals[s] = s;
  for (uint32_t e = off[s]; e < off[s + 1]; e++) {
    const uint32_t t = __ldg(&arcs[e].nextstate);
    keys[(size_t)n + e] = (unsigned long long)t + 1ull;
    vals[(size_t)n + e] = n + e;
    src_of[e] = s;
    atomicAdd(&rows[t + 1], 1u);
  }
}

// out[k] = the k-th entry of the sorted list, rewritten as an arc of the reversed machine.
__global__ void __launch_bounds__(kThreads)
k_rev_gather(const Tr* __restrict__ arcs, const float* __restrict__ fin, uint32_t n,
             const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ src_of, uint32_t count,
             Tr* __restrict__ out) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const uint32_t i = sorted[k];
  Tr tr;
  if (i < n) {  // reverse.rs:57-60
    tr.ilabel = kEps; tr.olabel = kEps; tr.weight = fin[i]; tr.nextstate = i + 1;
  } else {      // reverse.rs:62-67
    const uint32_t e = i - n;
    const int4 v = __ldg(reinterpret_cast<const int4*>(&arcs[e]));
    tr.ilabel = (uint32_t)v.x; tr.olabel = (uint32_t)v.y; tr.weight = __int_as_float(v.z);
    tr.nextstate = src_of[e] + 1;
  }
  *reinterpret_cast<int4*>(&out[k]) = *reinterpret_cast<const int4*>(&tr);
}

// OR of the add_tr_properties events (fst_types.h: arc_events) of every arc of a CSR machine.
__global__ void __launch_bounds__(kThreads)
k_arc_events(const uint32_t* __restrict__ off, const Tr* __restrict__ arcs, uint32_t n, uint32_t* __restrict__ out) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t ev = 0;
  if (s < n) {
    Tr prev;
    bool has_prev = false;
    for (uint32_t e = off[s]; e < off[s + 1]; e++) {
      const int4 v = __ldg(reinterpret_cast<const int4*>(&arcs[e]));
      Tr tr; tr.ilabel = (uint32_t)v.x; tr.olabel = (uint32_t)v.y; tr.weight = __int_as_float(v.z); tr.nextstate = (uint32_t)v.w;
      ev |= props::arc_events(s, tr, has_prev ? &prev : nullptr);
      prev = tr; has_prev = true;
    }
  }
  ev = __reduce_or_sync(0xFFFFFFFFu, ev);
  if ((threadIdx.x & 31) == 0 && ev) atomicOr(out, ev);
}

__global__ void k_rev_finals(float* __restrict__ fin, uint32_t n1, uint32_t final_state, bool has_final) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n1) fin[s] = (has_final && s == final_state) ? 0.0f : w_zero();
}

// One row of the reversed machine, packed for the host search: out[0] = {#arcs, 0, 0, 0}; when the row fits `cap`,
// out[1 + k] = arc k and dists[k] = forward distance of its source state (distance_2[nextstate], i.e. dist[nextstate
// - 1]).  The host never holds the distance array or the row offsets of a multi-million-state machine.
// One hop ahead: a best-first search pops, more often than not, a state it has just pushed, i.e. the target of one of
// these arcs.  Blocks 1.. therefore also pack the rows of the first kPrefetchKids targets (those with at most
// kKidCap arcs) behind the main row: kid slot k = kid_out + k * (1 + kKidCap) entries, header {#arcs or ~0 = not
// packed, state, 0, 0}, distances likewise in kid_dists.  The host keeps them in a cache keyed by state.
constexpr uint32_t kPrefetchKids = 64, kKidCap = 32;
__global__ void __launch_bounds__(kThreads)
k_fetch_row(const uint32_t* __restrict__ roff, const Tr* __restrict__ rarcs, const float* __restrict__ dist,
            uint32_t n, uint32_t q, uint32_t cap, int4* __restrict__ out, float* __restrict__ dists,
            int4* __restrict__ kid_out, float* __restrict__ kid_dists, uint32_t main_blocks) {
  const uint32_t b = roff[q], deg = roff[q + 1] - b;
  if (blockIdx.x >= main_blocks) {  // ---- kid rows: one warp per kid
    const uint32_t kid = (blockIdx.x - main_blocks) * (kThreads / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
    if (kid >= kPrefetchKids) return;
    int4* const slot = kid_out + (size_t)kid * (1 + kKidCap);
    if (deg > kPrefetchKids || kid >= deg) { if (lane == 0) slot[0] = make_int4(-1, 0, 0, 0); return; }
    const uint32_t cs = __ldg(&rarcs[b + kid].nextstate);
    const uint32_t cb = roff[cs], cdeg = roff[cs + 1] - cb;
    if (cdeg > kKidCap) { if (lane == 0) slot[0] = make_int4(-1, (int)cs, 0, 0); return; }
    if (lane == 0) slot[0] = make_int4((int)cdeg, (int)cs, 0, 0);
    if (lane < cdeg) {
      const int4 v = __ldg(reinterpret_cast<const int4*>(&rarcs[cb + lane]));
      slot[1 + lane] = v;
      const uint32_t src = (uint32_t)v.w - 1u;
      kid_dists[(size_t)kid * kKidCap + lane] = (dist && src < n) ? dist[src] : w_zero();
    }
    return;
  }
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid == 0) out[0] = make_int4((int)deg, 0, 0, 0);
  if (deg > cap) return;
  for (uint32_t k = tid; k < deg; k += main_blocks * blockDim.x) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(&rarcs[b + k]));
    out[1 + k] = v;
    const uint32_t src = (uint32_t)v.w - 1u;  // rows never point at the superinitial state
    dists[k] = (dist && src < n) ? dist[src] : w_zero();
  }
}

// Rows of MANY states of the reversed machine at once (unique = true: one subset state of the determinized machine can
// hold thousands of states): degrees, then — after an exclusive scan — the arcs and the forward distances of their
// targets, packed row after row in the order of `states`.
__global__ void __launch_bounds__(kThreads)
k_rows_deg(const uint32_t* __restrict__ roff, const uint32_t* __restrict__ states, uint32_t m, uint32_t* __restrict__ deg) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) { const uint32_t q = states[i]; deg[i] = roff[q + 1] - roff[q]; }
  else if (i == m) deg[i] = 0u;
}
__global__ void __launch_bounds__(kThreads)
k_rows_gather(const uint32_t* __restrict__ roff, const Tr* __restrict__ rarcs, const float* __restrict__ dist, uint32_t n,
              const uint32_t* __restrict__ states, uint32_t m, const uint32_t* __restrict__ off,
              int4* __restrict__ out, float* __restrict__ dists) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
  if (w >= m) return;
  const uint32_t q = states[w], b = roff[q], deg = roff[q + 1] - b, o = off[w];
  for (uint32_t k = lane; k < deg; k += 32u) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(&rarcs[b + k]));
    out[o + k] = v;
    const uint32_t src = (uint32_t)v.w - 1u;
    dists[o + k] = (dist && src < n) ? dist[src] : w_zero();
  }
}

// Mapped page-locked staging area: [header + cap arcs as int4][cap distances].
struct Staging {
  void* base = nullptr;
  size_t cap = 0;
  int4* arcs() const { return static_cast<int4*>(base); }
  float* dists() const { return reinterpret_cast<float*>(static_cast<char*>(base) + (cap + 1) * 16); }
  int4* kid_arcs() const { return reinterpret_cast<int4*>(static_cast<char*>(base) + (cap + 1) * 16 + cap * 4); }
  float* kid_dists() const { return reinterpret_cast<float*>(kid_arcs() + (size_t)kPrefetchKids * (1 + kKidCap)); }
  void ensure(size_t want) {
    if (want <= cap) return;
    if (base) cudaFreeHost(base);
    base = nullptr; cap = 0;
    size_t c = 4096;
    while (c < want) c <<= 1;
    const size_t bytes = (c + 1) * 16 + c * 4 + (size_t)kPrefetchKids * (1 + kKidCap) * 16 + (size_t)kPrefetchKids * kKidCap * 4;
    B200_CUDA(cudaHostAlloc(&base, bytes, cudaHostAllocMapped | cudaHostAllocPortable));
    cap = c;
  }
  ~Staging() { if (base) cudaFreeHost(base); }
};
Staging& thread_staging() {
  thread_local Staging st;
  return st;
}

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// shortest_path.rs:284-286; `==` / `!=` on TropicalWeight are the KDELTA-approximate ones (semiring.rs:159-168)
inline bool natural_less(float w1, float w2) { return w_approx_eq(w_plus(w1, w2), w1) && !w_approx_eq(w1, w2); }
// TropicalWeight::approx_equal(.., delta) (tropical_weight.rs:72-74, utils_float.rs:1-3)
inline bool approx_equal(float a, float b, float delta) { return std::fabs(a - b) <= delta; }

}  // namespace

DevFst reverse_device(const DevFst& f, cudaStream_t s, uint64_t* launches) {
  const uint32_t n = f.num_states, A = f.num_arcs;
  if ((size_t)n + A >= 0xFFFFFFF0ull) throw FstError("reverse: machine too large for 32-bit arc ids");
  DevFst r(s);
  r.num_states = n + 1;
  r.has_start = true; r.start = 0;
  const size_t total = (size_t)n + A;
  DevBuf<unsigned long long> k_in(s, total ? total : 1), k_out(s, total ? total : 1);
  DevBuf<uint32_t> v_in(s, total ? total : 1), v_out(s, total ? total : 1), src_of(s, A ? A : 1);
  DevBuf<uint8_t> tmp(s);
  r.offsets.reserve_discard((size_t)n + 3);  // rows 0 .. n (+ the sentinel row n + 1 during construction)
  B200_CUDA(cudaMemsetAsync(r.offsets.p, 0, ((size_t)n + 3) * 4, s));
  uint64_t nl = 0;
  if (n) {
    k_rev_keys<<<blocks_for(n), kThreads, 0, s>>>(f.offsets.p, f.arcs.p, f.finals.p, n, k_in.p, v_in.p, src_of.p,
                                                  r.offsets.p);
    nl++;
  }
  exclusive_sum_u32(r.offsets.p, r.offsets.p, (size_t)n + 2, tmp, s);  // offsets[n + 1] = #real arcs
  int bits = 1;
  while (((unsigned long long)n + 1ull) >> bits) bits++;
  if (total) sort_pairs_u64_u32(k_in.p, k_out.p, v_in.p, v_out.p, total, bits, tmp, s);
  uint32_t count = read_u32(r.offsets.p + n + 1, s);
  r.num_arcs = count;
  r.arcs.reserve_discard(count ? count : 1);
  if (count) {
    k_rev_gather<<<blocks_for(count), kThreads, 0, s>>>(f.arcs.p, f.finals.p, n, v_out.p, src_of.p, count, r.arcs.p);
    nl++;
  }
  r.finals.reserve_discard((size_t)n + 1);
  k_rev_finals<<<blocks_for((size_t)n + 1), kThreads, 0, s>>>(r.finals.p, n + 1, f.start + 1, f.has_start);
  nl++;
  r.props = 0;  // the n-best route never reads it; reverse_fst_device() below computes it (reverse.rs:78-83)
  if (launches) *launches += nl + 2;  // + scan + sort (library passes counted once each)
  return r;
}

// Property word of reverse(f): replay reverse.rs' mutation sequence on VectorFst::new() — add_state, add_states(n),
// set_final of the old start, every set_trs_unchecked (an order-independent function of the arcs' events, one
// reduction on the device), set_start — then OR in reverse_properties(input, true)  (reverse.rs:40-83).
static uint64_t reversed_props(const DevFst& f, const DevFst& r, cudaStream_t s) {
  DevBuf<uint32_t> ev(s, 1);
  B200_CUDA(cudaMemsetAsync(ev.p, 0, 4, s));
  k_arc_events<<<blocks_for(r.num_states), kThreads, 0, s>>>(r.offsets.p, r.arcs.p, r.num_states, ev.p);
  const uint32_t events = read_u32(ev.p, s);
  uint64_t p = props::kNull;
  p = props::on_add_state(p);
  p &= props::kKeepOnAddState;
  const float one = 0.0f;
  if (f.has_start) p = props::on_set_final(p, nullptr, &one);
  p = props::apply_arc_events(p, events, r.num_arcs > 0);
  p = props::on_set_start(p);
  return (props::of_reverse(f.props, true) | p) & props::kTrinary;
}

CsrFst reverse_fst_device(const DevFst& f, cudaStream_t s) {
  DevFst r = reverse_device(f, s, nullptr);
  r.props = reversed_props(f, r, s);
  return download(r, s);
}

CsrFst n_shortest_paths_device(const DevFst& f, const QueuePlan& plan,
                               size_t nshortest, float delta, NShortestStats* stats, cudaStream_t s,
                               bool force_serial, bool unique) {
  NShortestStats local;
  NShortestStats& st = stats ? *stats : local;
  st = NShortestStats();
  const double t_begin = now_ms();
  CsrFst empty;  // FO::new(): shortest_path.rs:419-434 return the untouched new FST (null_properties)
  if (nshortest == 0) return empty;
  const uint32_t n = f.num_states;

  // ---- device: forward distances (shortest_path.rs:139-140) and the reversed machine (:142)
  DevBuf<float> d_dist(s);
  double t0 = now_ms();
  shortest_distance_device(f, plan, delta, d_dist, &st.distance, s, force_serial);
  B200_CUDA(cudaStreamSynchronize(s));
  st.ms_distance = (float)(now_ms() - t0);
  t0 = now_ms();
  DevFst r = reverse_device(f, s, &st.distance.kernel_launches);
  const size_t r_row0_len = read_u32(r.offsets.p + 1, s);
  if (unique && !(reversed_props(f, r, s) & props::kAcceptor))  // determinize_fsa_op.rs:137-139
    throw FstError("DeterminizeFsaImpl : expected acceptor as argument");
  st.ms_reverse = (float)(now_ms() - t0);
  t0 = now_ms();

  // Rows of the reversed machine are fetched on demand into a page-locked staging buffer: one small kernel gathers
  // the row's arcs and the forward distances of their source states (k_fetch_row).
  const float* d_dist_p = (n && f.has_start) ? d_dist.p : nullptr;
  // Staging buffer in mapped page-locked host memory, written by the kernel itself (zero copy): a row costs one
  // launch and one stream synchronisation.  It is cached per host thread and sized to the longest row seen.
  Staging& stage = thread_staging();
  stage.ensure(std::max<size_t>(r_row0_len, 4096));
  std::vector<Tr> row;
  std::vector<float> row_dist;
  struct CachedRow { std::vector<Tr> arcs; std::vector<float> dists; };
  std::unordered_map<uint32_t, CachedRow> row_cache;  // rows that arrived one hop ahead of their pop
  auto fetch_row = [&](uint32_t q) {
    auto hit = row_cache.find(q);
    if (hit != row_cache.end()) {  // rows are immutable; a state can be expanded up to nshortest times
      row = hit->second.arcs; row_dist = hit->second.dists;
      st.rows_cached++;
      return;
    }
    while (true) {
      const uint32_t cap = (uint32_t)stage.cap;
      const size_t hint = q == 0 ? std::max<size_t>(r_row0_len, 1) : 256;  // rows are short; the kernel strides
      const uint32_t main_blocks = blocks_for(std::min<size_t>(cap, hint));
      const uint32_t kid_blocks = q == 0 ? 0u : kPrefetchKids / (kThreads / 32);
      k_fetch_row<<<main_blocks + kid_blocks, kThreads, 0, s>>>(r.offsets.p, r.arcs.p, d_dist_p, n, q, cap, stage.arcs(),
                                                                stage.dists(), stage.kid_arcs(), stage.kid_dists(),
                                                                main_blocks);
      st.distance.kernel_launches++;
      B200_CUDA(cudaStreamSynchronize(s));
      const uint32_t deg = (uint32_t)stage.arcs()[0].x;
      if (deg <= cap) {
        row.resize(deg); row_dist.resize(deg);
        if (deg) {
          std::memcpy(row.data(), stage.arcs() + 1, (size_t)deg * 16);
          std::memcpy(row_dist.data(), stage.dists(), (size_t)deg * 4);
        }
        if (q != 0 && deg <= 4096) { CachedRow& cr = row_cache[q]; cr.arcs = row; cr.dists = row_dist; }
        if (kid_blocks) {
          for (uint32_t k = 0; k < kPrefetchKids && k < deg; k++) {
            const int4* slot = stage.kid_arcs() + (size_t)k * (1 + kKidCap);
            if (slot[0].x < 0) continue;
            const uint32_t cdeg = (uint32_t)slot[0].x, cs = (uint32_t)slot[0].y;
            CachedRow& cr = row_cache[cs];
            cr.arcs.resize(cdeg); cr.dists.resize(cdeg);
            if (cdeg) {
              std::memcpy(cr.arcs.data(), slot + 1, (size_t)cdeg * 16);
              std::memcpy(cr.dists.data(), stage.kid_dists() + (size_t)k * kKidCap, (size_t)cdeg * 4);
            }
          }
        }
        break;
      }
      stage.ensure(deg);  // grow to the row and fetch again
    }
    st.rows_fetched++; st.arcs_fetched += row.size();
  };
  // distance_2 of shortest_path.rs:153-154: index 0 = the superinitial state (d0), index q = distance[q - 1]
  float d0 = w_zero();
  fetch_row(0);
  const std::vector<Tr> row0 = row;
  const std::vector<float> row0_dist = row_dist;
  for (size_t k = 0; k < row0.size(); k++) d0 = w_plus(d0, w_times(row0[k].weight, row0_dist[k]));  // :144-150


  const uint32_t rfinal = f.has_start ? f.start + 1 : kNoState;
  // ---- unique = true (shortest_path.rs:156-165): the search below walks the DETERMINIZED reversed machine.  The
  // reference materialises determinize_with_distance(rfst, distance_2) completely first (lazy_fst.rs:226-259); the
  // search only ever looks at the subset states it pops, so here a subset state gets its arcs the first time it is
  // popped (determinize_fsa_op.rs:56-101 + norm_tr :147-178) from rows of the reversed machine fetched as above, and
  // its distance (state_table.rs:20-35) when it is first seen.  Subsets are kept sorted by state — the reference's own
  // element order comes out of a RandomState HashMap (:158-171) and is not reproducible; paths and weights do not
  // depend on it.  Identity of a subset: its filter state (the start state for the start subset, 0 otherwise), its
  // states and the bit patterns of its quantised residual weights (Hash of the reference's tuple).
  struct DetElem { uint32_t state; float w; float d2; };  // d2 = distance_2[state], delivered with the arc that led here
  struct DetRow { std::vector<Tr> arcs; std::vector<float> dists; float fin = w_zero(); bool built = false; };
  std::vector<std::vector<DetElem>> det_subset;
  std::vector<float> det_dist;
  std::vector<DetRow> det_rows;
  std::map<std::vector<uint64_t>, uint32_t> det_ids;
  auto det_find = [&](const std::vector<DetElem>& sub, uint32_t filter_state) -> uint32_t {
    std::vector<uint64_t> key;
    key.reserve(sub.size() + 1);
    key.push_back(filter_state);
    for (const DetElem& e : sub) {
      const float w = e.w == 0.0f ? 0.0f : e.w;  // -0.0 and 0.0 hash alike
      uint32_t bits;
      std::memcpy(&bits, &w, 4);
      key.push_back(((uint64_t)e.state << 32) | bits);
    }
    auto it = det_ids.find(key);
    if (it != det_ids.end()) return it->second;
    const uint32_t id = (uint32_t)det_subset.size();
    det_ids.emplace(std::move(key), id);
    det_subset.push_back(sub);
    float outd = w_zero();
    for (const DetElem& e : sub) outd = w_plus(outd, w_times(e.w, e.d2));
    det_dist.push_back(outd);
    det_rows.emplace_back();
    return id;
  };
  auto quantize = [&](float v) -> float {  // semiring.rs:135-142
    if (std::isinf(v)) return v;
    return std::floor((v / delta) + 0.5f) * delta;
  };
  std::vector<Tr> many_arcs;
  std::vector<float> many_dist;
  std::vector<uint32_t> many_off;
  auto fetch_rows = [&](const std::vector<uint32_t>& states) {  // rows of `states`, packed: row i = [many_off[i], many_off[i + 1])
    const uint32_t m = (uint32_t)states.size();
    DevBuf<uint32_t> d_states(s, m), d_deg(s, (size_t)m + 1), d_off(s, (size_t)m + 1);
    DevBuf<uint8_t> tmp(s);
    B200_CUDA(cudaMemcpyAsync(d_states.p, states.data(), (size_t)m * 4, cudaMemcpyHostToDevice, s));
    k_rows_deg<<<blocks_for((size_t)m + 1), kThreads, 0, s>>>(r.offsets.p, d_states.p, m, d_deg.p);
    exclusive_sum_u32(d_deg.p, d_off.p, (size_t)m + 1, tmp, s);
    many_off.resize((size_t)m + 1);
    B200_CUDA(cudaMemcpyAsync(many_off.data(), d_off.p, ((size_t)m + 1) * 4, cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaStreamSynchronize(s));
    const uint32_t total = many_off[m];
    many_arcs.resize(total); many_dist.resize(total);
    if (total) {
      DevBuf<int4> d_out(s, total);
      DevBuf<float> d_dists(s, total);
      k_rows_gather<<<blocks_for((size_t)m * 32), kThreads, 0, s>>>(r.offsets.p, r.arcs.p, d_dist_p, n, d_states.p, m, d_off.p,
                                                                    d_out.p, d_dists.p);
      B200_CUDA(cudaMemcpyAsync(many_arcs.data(), d_out.p, (size_t)total * 16, cudaMemcpyDeviceToHost, s));
      B200_CUDA(cudaMemcpyAsync(many_dist.data(), d_dists.p, (size_t)total * 4, cudaMemcpyDeviceToHost, s));
      B200_CUDA(cudaStreamSynchronize(s));
    }
    st.distance.kernel_launches += 3;
    st.rows_fetched += m; st.arcs_fetched += total;
  };
  auto det_row = [&](uint32_t q) -> const DetRow& {
    if (det_rows[q].built) return det_rows[q];
    const std::vector<DetElem> src = det_subset[q];  // a copy: det_find below grows the tables
    std::map<uint32_t, std::vector<DetElem>> label_map;  // BTreeMap: arcs leave in label order
    float fin = w_zero();
    constexpr size_t kOneByOne = 4;  // small subsets go through the per-row fetch and its one-hop cache
    std::vector<uint32_t> want;
    if (src.size() > kOneByOne) {
      for (const DetElem& e : src) if (e.state != 0) want.push_back(e.state);
      fetch_rows(want);
    }
    size_t wi = 0;
    for (const DetElem& e : src) {
      const Tr* prow = row0.data();
      const float* pdist = row0_dist.data();
      size_t len = row0.size();
      if (e.state != 0) {
        if (src.size() > kOneByOne) {
          prow = many_arcs.data() + many_off[wi]; pdist = many_dist.data() + many_off[wi];
          len = many_off[wi + 1] - many_off[wi];
          wi++;
        } else {
          fetch_row(e.state);
          prow = row.data(); pdist = row_dist.data(); len = row.size();
        }
      }
      for (size_t k = 0; k < len; k++) {
        const Tr& tr = prow[k];
        label_map[tr.ilabel].push_back(DetElem{tr.nextstate, w_times(e.w, tr.weight), pdist[k]});
      }
      // final weight (:103-120): the reversed machine has one final state, the old start, with weight one()
      fin = w_plus(fin, w_times(e.w, e.state == rfinal ? 0.0f : w_zero()));
    }
    DetRow built;
    built.fin = fin;
    for (auto& kv : label_map) {
      std::vector<DetElem>& d = kv.second;
      std::stable_sort(d.begin(), d.end(), [](const DetElem& x, const DetElem& y) { return x.state < y.state; });
      float w = w_zero();
      for (const DetElem& e : d) w = w_plus(w, e.w);  // DefaultCommonDivisor: plus (divisors.rs:16-22)
      std::vector<DetElem> merged;
      for (const DetElem& e : d) {
        if (!merged.empty() && merged.back().state == e.state) merged.back().w = w_plus(merged.back().w, e.w);
        else merged.push_back(e);
      }
      for (DetElem& e : merged) e.w = quantize(e.w - w);  // divide (tropical_weight.rs:127-132), then quantize
      const uint32_t dest = det_find(merged, 0);
      built.arcs.push_back(Tr{kv.first, kv.first, w, dest});
      built.dists.push_back(det_dist[dest]);
    }
    built.built = true;
    det_rows[q] = std::move(built);
    st.det_states_expanded++;
    return det_rows[q];
  };
  // ---- host: n_shortest_path (shortest_path.rs:409-518) over the reversed machine
  if (w_is_zero(d0)) { st.ms_total = (float)(now_ms() - t_begin); return empty; }  // :427-434 (istart = 0 always exists)
  // The result tree is written straight in CSR form: state 0 = start (arcs appended as complete paths are popped),
  // state 1 = final, every later state carries exactly one arc; the property word replays the reference's mutation
  // sequence (add_state / set_start / set_final / add_tr) in its order.
  uint64_t pw = props::kNull;
  pw = props::on_add_state(pw);                 // ostart
  pw = props::on_set_start(pw);
  pw = props::on_add_state(pw);                 // final_state
  { const float one = 0.0f; pw = props::on_set_final(pw, nullptr, &one); }
  const StateId ostart = 0, final_state = 1;
  std::vector<Tr> start_arcs, single_arc;       // single_arc[i] = the arc of state i + 2
  struct Pair { bool some; uint32_t state; float w; };
  std::vector<Pair> pairs(final_state + 1, Pair{false, 0, w_zero()});
  pairs[final_state] = Pair{true, 0, 0.0f};
  // Keys of ShortestPathCompare (:323-338) are pure functions of a pair, fixed at push time: cache them.
  struct Key { float w; bool some; };
  std::vector<Key> keys(pairs.size(), Key{w_zero(), false});
  keys[final_state] = Key{w_times(d0, 0.0f), true};
  auto compare = [&](StateId x, StateId y) -> bool {
    const Key &kx = keys[x], &ky = keys[y];
    if (!kx.some && ky.some) return natural_less(ky.w, kx.w) || approx_equal(kx.w, ky.w, delta);
    if (kx.some && !ky.some) return natural_less(ky.w, kx.w) && !approx_equal(kx.w, ky.w, delta);
    return natural_less(ky.w, kx.w);
  };
  std::vector<StateId> heap;  // Heap of :341-407, restated (iteratively) with the same comparisons in the same order
  auto sift_up = [&](size_t idx) {
    while (idx > 0) {
      const size_t parent = (idx - 1) / 2;
      if (!compare(heap[parent], heap[idx])) break;
      std::swap(heap[idx], heap[parent]);
      idx = parent;
    }
  };
  auto sift_down = [&](size_t idx) {
    while (true) {
      const StateId cur = heap[idx];
      const size_t c1 = 2 * idx + 1, c2 = 2 * idx + 2;
      size_t big;
      if (c1 >= heap.size() && c2 >= heap.size()) return;
      else if (c1 < heap.size() && c2 >= heap.size()) big = c1;
      else if (compare(heap[c1], heap[c2])) big = c2;
      else big = c1;
      if (compare(heap[big], cur)) return;
      std::swap(heap[idx], heap[big]);
      idx = big;
    }
  };
  auto push = [&](StateId v) { heap.push_back(v); sift_up(heap.size() - 1); };
  auto pop = [&]() -> StateId {
    const StateId top = heap[0];
    if (heap.size() == 1) heap.clear();
    else { heap[0] = heap.back(); heap.pop_back(); sift_down(0); }
    return top;
  };
  auto new_state = [&](const Tr& tr, const Pair& pr, float dist_of_pair_state) -> StateId {
    const StateId next = (StateId)(single_arc.size() + 2);
    pw = props::on_add_state(pw);
    pairs.push_back(pr);
    keys.push_back(Key{w_times(pr.some ? dist_of_pair_state : 0.0f, pr.w), pr.some});
    single_arc.push_back(tr);
    pw = props::on_add_tr(pw, next, tr, nullptr);
    return next;
  };
  pairs.reserve(r_row0_len + 1024); keys.reserve(r_row0_len + 1024); single_arc.reserve(r_row0_len + 1024);
  heap.reserve(r_row0_len + 1024);
  if (unique) {
    const uint32_t det_start = det_find(std::vector<DetElem>{DetElem{0u, 0.0f, d0}}, 0u);  // {(start, one)}, filter state = start = 0
    (void)det_start;                                                                     // = 0 = istart of the search below
  }
  push(final_state);
  const float limit = w_times(d0, w_zero());  // weight_threshold = zero(): :448-449
  // r of :451 — the reference grows a dense vector up to the largest popped state id (tens of MB on a multi-million-
  // state lattice for a few hundred pops); a hash map holds the same counters
  std::unordered_map<uint32_t, size_t> seen;
  while (!heap.empty()) {
    const StateId state = pop();
    st.heap_pops++;
    const Pair p = pairs[state];
    const uint32_t first_real = p.some ? p.state + 1 : 0;
    // d (x) p.1 of :463-474 is exactly the cached heap key
    if (natural_less(limit, keys[state].w)) continue;
    const size_t n_seen = ++seen[first_real];
    if (!p.some) {
      const Tr tr{0, 0, 0.0f, state};
      pw = props::on_add_tr(pw, ostart, tr, start_arcs.empty() ? nullptr : &start_arcs.back());
      start_arcs.push_back(tr);
    }
    if (!p.some && n_seen == nshortest) break;
    if (n_seen > nshortest) continue;
    if (!p.some) continue;
    const std::vector<Tr>* prow = &row0;
    const std::vector<float>* pdist = &row0_dist;
    float fin = p.state == rfinal ? 0.0f : w_zero();  // the only final state of the reversed machine: reverse.rs:54-56
    if (unique) {
      const DetRow& dr = det_row(p.state);
      prow = &dr.arcs; pdist = &dr.dists; fin = dr.fin;
    } else if (p.state != 0) {
      fetch_row(p.state); prow = &row; pdist = &row_dist;
    }
    for (size_t k = 0; k < prow->size(); k++) {
      const Tr& rarc = (*prow)[k];
      const StateId next = new_state(Tr{rarc.ilabel, rarc.olabel, rarc.weight, state},
                                     Pair{true, rarc.nextstate, w_times(p.w, rarc.weight)}, (*pdist)[k]);
      push(next);
    }
    if (!w_is_zero(fin)) {  // :494-505
      const StateId next = new_state(Tr{0, 0, fin, state}, Pair{false, 0, w_times(p.w, fin)}, 0.0f);
      push(next);
    }
  }
  st.ms_search_host = (float)(now_ms() - t0);
  t0 = now_ms();
  // ---- connect (:511) + shortest_path_properties(.., false) (:512-515).  The search tree never leaves the host: every
  // state >= 2 has exactly one arc, towards the state it was expanded from, and those chains all end in the final
  // state 1, so every state is coaccessible as soon as the start state has an arc, and the accessible states are
  // exactly the chains hanging off the start state's arcs.  del_states (mutable_fst.rs:132-189) keeps them in id order.
  const size_t ns = single_arc.size() + 2;
  st.states_before_trim = ns;
  CsrFst out;
  out.props = props::of_shortest_path(props::after_connect(pw & props::kTrinary), false) & props::kTrinary;
  if (!start_arcs.empty()) {
    std::vector<uint8_t> keep(ns, 0);
    keep[ostart] = 1; keep[final_state] = 1;
    for (const Tr& a : start_arcs)
      for (StateId x = a.nextstate; x >= 2 && !keep[x]; x = single_arc[x - 2].nextstate) keep[x] = 1;
    std::vector<uint32_t> new_id(ns, 0);
    uint32_t nk = 0;
    for (size_t x = 0; x < ns; x++) if (keep[x]) new_id[x] = nk++;
    out.offsets.resize((size_t)nk + 1);
    out.finals.assign(nk, w_zero());
    out.arcs.resize(start_arcs.size() + (nk - 2));
    size_t o = 0;
    out.offsets[0] = 0;
    for (const Tr& a : start_arcs) { Tr t = a; t.nextstate = new_id[a.nextstate]; out.arcs[o++] = t; }
    out.offsets[1] = (uint32_t)o;          // new id of the start state is 0
    out.offsets[2] = (uint32_t)o;          // the final state (new id 1) has no arcs
    out.finals[1] = 0.0f;
    for (size_t x = 2; x < ns; x++) {
      if (!keep[x]) continue;
      Tr t = single_arc[x - 2];
      t.nextstate = new_id[t.nextstate];
      out.arcs[o++] = t;
      out.offsets[(size_t)new_id[x] + 1] = (uint32_t)o;
    }
    out.has_start = true; out.start = 0;
  }
  const float ms_trim = (float)(now_ms() - t0);
  st.ms_total = (float)(now_ms() - t_begin);
  if (std::getenv("B200_NSHORTEST_TRACE"))
    std::fprintf(stderr, "[nshortest] n=%zu distance %.2f ms (path %d) reverse %.2f ms search %.2f ms (%llu pops, %llu rows fetched + %llu from the one-hop cache, "
                 "%llu arcs fetched, %llu tree states) trim %.2f ms total %.2f ms\n", nshortest, st.ms_distance,
                 st.distance.path, st.ms_reverse, st.ms_search_host, (unsigned long long)st.heap_pops,
                 (unsigned long long)st.rows_fetched, (unsigned long long)st.rows_cached, (unsigned long long)st.arcs_fetched,
                 (unsigned long long)st.states_before_trim, ms_trim, st.ms_total);
  return out;
}

}  // namespace b200
