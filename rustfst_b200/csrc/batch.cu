// batch.cu — per-acceptor split of a batched composition, on the device.
//
// The batched mode (BASELINE.json configs[4]) composes the disjoint union of many acceptors with one shared transducer
// in ONE device BFS; the result is the union of the individual compositions (compose_static.rs:198-298 applied per
// acceptor).  States of one component keep their relative order in the union, and that relative order is exactly the
// numbering of the stand-alone composition (the same first-emission BFS restricted to the component), so the split
// is a STABLE partition of the states by component: one radix sort of (component, state), a prefix sum of the
// out-degrees in the new order and one gather that rewrites next states to component-local ids.  The results leave
// the device as one packed block (PackedBatch): no per-result allocation, copy or handle.
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "algos.h"

namespace b200 {
namespace {
double wall_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// component of every result state: the acceptor whose union-state range contains its fst1 component
__global__ void k_split_keys(const uint32_t* __restrict__ tag, uint32_t n_states, const uint32_t* __restrict__ base_state,
                             uint32_t n_acc, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_states) return;
  const uint32_t t = __ldg(&tag[s]);
  uint32_t lo = 0, hi = n_acc;  // largest i with base_state[i] <= t
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(&base_state[mid]) <= t) lo = mid; else hi = mid; }
  keys[s] = lo;
  vals[s] = s;
}
// inverse permutation + out-degrees in the new order
__global__ void k_split_inverse(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ off, uint32_t n_states,
                                uint32_t* __restrict__ inv, uint32_t* __restrict__ ndeg) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n_states) return;
  if (k == n_states) { ndeg[k] = 0; return; }
  const uint32_t s = __ldg(&perm[k]);
  inv[s] = k;
  ndeg[k] = __ldg(&off[s + 1]) - __ldg(&off[s]);
}
// first state of every component in the new order (lower bound of i in the sorted keys), i in [0, n_acc]
__global__ void k_split_bounds(const unsigned long long* __restrict__ sorted, uint32_t n_states, uint32_t n_acc,
                               uint32_t* __restrict__ state_off) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_acc) return;
  uint32_t lo = 0, hi = n_states;  // first k with sorted[k] >= i
  while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (sorted[mid] < i) lo = mid + 1; else hi = mid; }
  state_off[i] = lo;
}
__global__ void k_split_gather(const uint32_t* __restrict__ perm, const unsigned long long* __restrict__ sorted,
                               const uint32_t* __restrict__ inv, const uint32_t* __restrict__ state_off,
                               const uint32_t* __restrict__ off, const Tr* __restrict__ arcs, const float* __restrict__ fin,
                               const uint32_t* __restrict__ noff, uint32_t n_states, Tr* __restrict__ narcs,
                               float* __restrict__ nfin) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_states) return;
  const uint32_t s = __ldg(&perm[k]);
  const uint32_t first = __ldg(&state_off[(uint32_t)sorted[k]]);
  nfin[k] = __ldg(&fin[s]);
  uint32_t o = __ldg(&noff[k]);
  for (uint32_t e = __ldg(&off[s]); e < __ldg(&off[s + 1]); e++) {
    int4 v = __ldg(reinterpret_cast<const int4*>(&arcs[e]));
    v.w = (int)(__ldg(&inv[(uint32_t)v.w]) - first);  // arcs never leave their component
    *reinterpret_cast<int4*>(&narcs[o++]) = v;
  }
}
__global__ void k_split_starts(const uint32_t* __restrict__ start_map, const uint32_t* __restrict__ inv,
                               const uint32_t* __restrict__ state_off, const uint32_t* __restrict__ noff, uint32_t n_acc,
                               int32_t* __restrict__ starts, uint32_t* __restrict__ arc_off) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_acc) return;
  arc_off[i] = __ldg(&noff[__ldg(&state_off[i])]);
  if (i == n_acc) return;
  const uint32_t m = __ldg(&start_map[i]);
  starts[i] = m == 0xFFFFFFFFu ? -1 : (int32_t)(__ldg(&inv[m]) - __ldg(&state_off[i]));
}

}  // namespace

// `r` = result of the batched composition (union of the components), tag[s] = fst1 state of result state s,
// start_map[i] = result id of the start tuple of acceptor i (0xFFFFFFFF = trimmed away), base_state[i] = first union
// state of acceptor i (n_acc + 1 entries, host).  Fills every array of `out` except props.
void split_batch_device(const DevFst& r, const uint32_t* d_tag, const uint32_t* d_start_map,
                        const std::vector<uint32_t>& base_state, PackedBatch& out, uint64_t* launches, cudaStream_t s) {
  const uint32_t n_acc = (uint32_t)base_state.size() - 1, ns = r.num_states, na = r.num_arcs;
  const bool trace = std::getenv("B200_BATCH_TRACE") != nullptr;
  const double t0 = wall_ms();
  out.n = n_acc;
  out.state_off.assign((size_t)n_acc + 1, 0);
  out.arc_off.assign((size_t)n_acc + 1, 0);
  out.starts.assign(n_acc, -1);
  out.offsets.resize((size_t)ns + 1);
  out.finals.resize(ns);
  out.arcs.resize(na);
  if (ns == 0) { out.offsets[0] = 0; return; }
  DevBuf<uint32_t> d_base(s, (size_t)n_acc + 1), v_in(s, ns), perm(s, ns), inv(s, ns), ndeg(s, (size_t)ns + 1);
  DevBuf<uint32_t> noff(s, (size_t)ns + 1), d_state_off(s, (size_t)n_acc + 1), d_arc_off(s, (size_t)n_acc + 1);
  DevBuf<int32_t> d_starts(s, n_acc);
  DevBuf<unsigned long long> k_in(s, ns), k_out(s, ns);
  DevBuf<Tr> narcs(s, na ? na : 1);
  DevBuf<float> nfin(s, ns);
  DevBuf<uint8_t> tmp(s);
  const double t1 = wall_ms();
  B200_CUDA(cudaMemcpyAsync(d_base.p, base_state.data(), ((size_t)n_acc + 1) * 4, cudaMemcpyHostToDevice, s));
  k_split_keys<<<blocks_for(ns), kThreads, 0, s>>>(d_tag, ns, d_base.p, n_acc, k_in.p, v_in.p);
  int bits = 1;
  while ((1ull << bits) < (unsigned long long)n_acc + 1) bits++;
  sort_pairs_u64_u32(k_in.p, k_out.p, v_in.p, perm.p, ns, bits, tmp, s);  // stable: relative order inside a component stays
  k_split_inverse<<<blocks_for((size_t)ns + 1), kThreads, 0, s>>>(perm.p, r.offsets.p, ns, inv.p, ndeg.p);
  exclusive_sum_u32(ndeg.p, noff.p, (size_t)ns + 1, tmp, s);
  k_split_bounds<<<blocks_for((size_t)n_acc + 1), kThreads, 0, s>>>(k_out.p, ns, n_acc, d_state_off.p);
  k_split_gather<<<blocks_for(ns), kThreads, 0, s>>>(perm.p, k_out.p, inv.p, d_state_off.p, r.offsets.p, r.arcs.p,
                                                     r.finals.p, noff.p, ns, narcs.p, nfin.p);
  k_split_starts<<<blocks_for((size_t)n_acc + 1), kThreads, 0, s>>>(d_start_map, inv.p, d_state_off.p, noff.p, n_acc,
                                                                    d_starts.p, d_arc_off.p);
  if (launches) *launches += 7;
  double t2 = t1;
  if (trace) { B200_CUDA(cudaStreamSynchronize(s)); t2 = wall_ms(); }
  // the three big arrays land in page-locked memory (PoolVec) and are queued first; the three small ones are ordinary
  // vectors — an asynchronous copy into pageable memory queued behind running kernels took milliseconds here, so they
  // are fetched after the stream has drained
  B200_CUDA(cudaMemcpyAsync(out.offsets.data(), noff.p, ((size_t)ns + 1) * 4, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaMemcpyAsync(out.finals.data(), nfin.p, (size_t)ns * 4, cudaMemcpyDeviceToHost, s));
  if (na) B200_CUDA(cudaMemcpyAsync(out.arcs.data(), narcs.p, (size_t)na * sizeof(Tr), cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  B200_CUDA(cudaMemcpy(out.state_off.data(), d_state_off.p, ((size_t)n_acc + 1) * 4, cudaMemcpyDeviceToHost));
  B200_CUDA(cudaMemcpy(out.arc_off.data(), d_arc_off.p, ((size_t)n_acc + 1) * 4, cudaMemcpyDeviceToHost));
  B200_CUDA(cudaMemcpy(out.starts.data(), d_starts.p, (size_t)n_acc * 4, cudaMemcpyDeviceToHost));
  if (trace)
    std::fprintf(stderr, "[split] host arrays + device buffers %.2f ms, kernels %.2f ms, download %.2f ms (%u states, %u arcs)\n",
                 t1 - t0, t2 - t1, wall_ms() - t2, ns, na);
}

// ---- PackedBatch <-> one contiguous byte block (what travels over NCCL to rank 0)
namespace {
constexpr uint64_t kPackMagic = 0x4B434150'30303242ull;  // "B200PACK"
template <class T>
void put(uint8_t*& w, const T* p, size_t n) {
  const size_t bytes = n * sizeof(T);
  const uint8_t* src = reinterpret_cast<const uint8_t*>(p);
  uint8_t* dst = w;
  // large arrays (the arcs of a batch are tens of MB) are copied by several host threads
  parallel_ranges(bytes, [&](size_t lo, size_t hi) { std::memcpy(dst + lo, src + lo, hi - lo); }, (size_t)4 << 20);
  w += bytes;
}
template <class T>
void get(const uint8_t*& r, const uint8_t* end, T* p, size_t n) {
  if ((size_t)(end - r) < n * sizeof(T)) throw FstError("packed batch: truncated block");
  if (n) std::memcpy(p, r, n * sizeof(T));
  r += n * sizeof(T);
}
}  // namespace

size_t PackedBatch::byte_size() const {
  return 4 * 8 + (state_off.size() + arc_off.size()) * 4 + starts.size() * 4 + props.size() * 8 + offsets.size() * 4 +
         finals.size() * 4 + arcs.size() * sizeof(Tr);
}
void PackedBatch::serialize(uint8_t* dst) const {
  uint8_t* w = dst;
  const uint64_t hdr[4] = {kPackMagic, n, finals.size(), arcs.size()};
  put(w, hdr, 4);
  put(w, state_off.data(), state_off.size());
  put(w, arc_off.data(), arc_off.size());
  put(w, starts.data(), starts.size());
  put(w, props.data(), props.size());
  put(w, offsets.data(), offsets.size());
  put(w, finals.data(), finals.size());
  put(w, arcs.data(), arcs.size());
}
PackedBatch PackedBatch::deserialize(const uint8_t* src, size_t len) {
  const uint8_t* r = src;
  const uint8_t* end = src + len;
  uint64_t hdr[4];
  get(r, end, hdr, 4);
  if (hdr[0] != kPackMagic) throw FstError("packed batch: bad magic");
  if (hdr[1] > 0x7FFFFFFFull || hdr[2] > 0x7FFFFFFFull || hdr[3] > 0xFFFFFFF0ull) throw FstError("packed batch: bad header");
  PackedBatch b;
  b.n = hdr[1];
  b.state_off.resize(b.n + 1); b.arc_off.resize(b.n + 1); b.starts.resize(b.n); b.props.resize(b.n);
  b.offsets.resize(hdr[2] + 1); b.finals.resize(hdr[2]); b.arcs.resize(hdr[3]);
  get(r, end, b.state_off.data(), b.state_off.size());
  get(r, end, b.arc_off.data(), b.arc_off.size());
  get(r, end, b.starts.data(), b.starts.size());
  get(r, end, b.props.data(), b.props.size());
  get(r, end, b.offsets.data(), b.offsets.size());
  get(r, end, b.finals.data(), b.finals.size());
  get(r, end, b.arcs.data(), b.arcs.size());
  // the block comes from another process: check every index before anything walks it
  if (b.state_off[b.n] != hdr[2] || b.arc_off[b.n] != hdr[3] || b.offsets[hdr[2]] != hdr[3]) throw FstError("packed batch: inconsistent totals");
  for (size_t i = 0; i < b.n; i++) {
    if (b.state_off[i] > b.state_off[i + 1] || b.arc_off[i] > b.arc_off[i + 1]) throw FstError("packed batch: offsets not monotone");
    const uint32_t ns = b.state_off[i + 1] - b.state_off[i];
    if (b.starts[i] >= 0 && (uint32_t)b.starts[i] >= ns) throw FstError("packed batch: start out of range");
    if (b.offsets[b.state_off[i]] != b.arc_off[i]) throw FstError("packed batch: arc offsets inconsistent");
    for (uint32_t e = b.arc_off[i]; e < b.arc_off[i + 1]; e++)
      if (b.arcs[e].nextstate >= ns) throw FstError("packed batch: transition out of range");
  }
  for (size_t s = 0; s < hdr[2]; s++)
    if (b.offsets[s] > b.offsets[s + 1]) throw FstError("packed batch: offsets not monotone");
  return b;
}
CsrFst PackedBatch::result(size_t i) const {
  if (i >= n) throw FstError("packed batch: index out of range");
  CsrFst c;
  const uint32_t s0 = state_off[i], s1 = state_off[i + 1], a0 = arc_off[i], a1 = arc_off[i + 1];
  c.offsets.resize((size_t)(s1 - s0) + 1);
  for (uint32_t s = s0; s <= s1; s++) c.offsets[s - s0] = offsets[s] - a0;
  c.finals.assign(finals.data() + s0, finals.data() + s1);
  c.arcs.assign(arcs.data() + a0, arcs.data() + a1);
  c.has_start = starts[i] >= 0;
  c.start = c.has_start ? (StateId)starts[i] : 0;
  c.props = props[i];
  return c;
}
void PackedBatch::append(const CsrFst& c) {  // host-side packing (heterogeneous batches composed one by one)
  if (n == 0 && state_off.empty()) { state_off.push_back(0); arc_off.push_back(0); offsets.push_back(0); }
  const uint32_t a0 = arc_off.back();
  offsets.pop_back();
  for (size_t s = 0; s <= c.num_states(); s++) offsets.push_back(a0 + c.offsets[s]);
  for (size_t s = 0; s < c.num_states(); s++) finals.push_back(c.finals[s]);
  for (const Tr& t : c.arcs) arcs.push_back(t);
  state_off.push_back(state_off.back() + (uint32_t)c.num_states());
  arc_off.push_back(a0 + (uint32_t)c.arcs.size());
  starts.push_back(c.has_start ? (int32_t)c.start : -1);
  props.push_back(c.props);
  n++;
}

}  // namespace b200
