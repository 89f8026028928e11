// iso.cu — isomorphic() as a device-side verifier.
//
// Replaces the success path of rustfst/src/algorithms/isomorphic.rs:49-160 (breadth-first pairing of states from the
// two start states; the arcs of a pair are compared after sorting by (ilabel, olabel, weight, nextstate); weights and
// final weights approx-equal with KDELTA).  Parallel restatement:
//   * the per-pair sorts become ONE stable LSD radix sort of all arcs of each machine by (state, ilabel, olabel, weight,
//     nextstate) — three passes over (key, index) pairs — done once;
//   * the pairing runs level by level: every pair of the current level compares final weights, degrees and its two
//     sorted arc rows; `pair[s1] <- s2` is a compare-and-swap, the winner appends the new pair to the next level.
// When every check passes the pairing is forced (it does not depend on the order in which pairs are visited), so the
// device answer `true` is the reference's answer.  A failed check is the reference's `false` unless the reference
// would have raised its "Non-determinism as an unweighted automaton" error, which depends on the visiting order: the
// device reports `false` only when no two neighbouring arcs of any visited row were equal as an unweighted automaton,
// and "undecided" otherwise (the caller then takes the sequential host restatement, host_fst.h isomorphic()).
#include "algos.h"

namespace b200 {
namespace {

constexpr uint32_t kNone = 0xFFFFFFFFu;
enum IsoFlag : uint32_t { kIsoFail = 1, kIsoNonDet = 2 };

__device__ __forceinline__ uint32_t ord_f32(float f) {  // monotone float -> uint32 (-0.0 and +0.0 coincide)
  if (f == 0.0f) f = 0.0f;
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void k_iso_src(const uint32_t* __restrict__ off, uint32_t n, uint32_t* __restrict__ src) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  for (uint32_t e = off[s]; e < off[s + 1]; e++) src[e] = s;
}
// key of pass `pass` for the arc at position idx[i]: 0 = (weight, nextstate), 1 = (ilabel, olabel), 2 = source state
__global__ void k_iso_keys(const Tr* __restrict__ arcs, const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx,
                           uint32_t a, int pass, unsigned long long* __restrict__ keys) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a) return;
  const uint32_t e = idx ? idx[i] : i;
  const int4 v = __ldg(reinterpret_cast<const int4*>(&arcs[e]));
  unsigned long long k;
  if (pass == 0) k = ((unsigned long long)ord_f32(__int_as_float(v.z)) << 32) | (uint32_t)v.w;
  else if (pass == 1) k = ((unsigned long long)(uint32_t)v.x << 32) | (uint32_t)v.y;
  else k = __ldg(&src[e]);
  keys[i] = k;
}
__global__ void k_iso_iota(uint32_t* __restrict__ idx, uint32_t a) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a) idx[i] = i;
}
__global__ void k_iso_gather(const Tr* __restrict__ arcs, const uint32_t* __restrict__ idx, uint32_t a, Tr* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a) *reinterpret_cast<int4*>(&out[i]) = __ldg(reinterpret_cast<const int4*>(&arcs[idx[i]]));
}

struct IsoParams {
  const uint32_t* off1; const Tr* arcs1; const float* fin1;
  const uint32_t* off2; const Tr* arcs2; const float* fin2;
  uint32_t* pair;            // pair[s1] = s2 or kNone
  uint32_t* q1; uint32_t* q2;  // paired states in visiting order
  uint32_t* ctl;             // [0] queue tail, [1] flags
  float delta;
};
__device__ __forceinline__ bool iso_approx(float x, float y, float delta) { return fabsf(x - y) <= delta; }

// one warp per pair of the level [lo, hi): lanes stride the two sorted rows
__global__ void __launch_bounds__(kThreads)
k_iso_level(IsoParams P, uint32_t lo, uint32_t hi) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
  if (lo + w >= hi) return;
  const uint32_t s1 = P.q1[lo + w], s2 = P.q2[lo + w];
  uint32_t flags = 0;
  const float f1 = P.fin1[s1], f2 = P.fin2[s2];
  const bool h1 = f1 != w_zero(), h2 = f2 != w_zero();
  if (h1 != h2 || (h1 && !iso_approx(f1, f2, P.delta))) flags |= kIsoFail;
  const uint32_t a1 = P.off1[s1], d1 = P.off1[s1 + 1] - a1, a2 = P.off2[s2], d2 = P.off2[s2 + 1] - a2;
  if (d1 != d2) flags |= kIsoFail;
  if (!flags) {
    for (uint32_t i = lane; i < d1; i += 32) {
      const Tr x = *reinterpret_cast<const Tr*>(&P.arcs1[a1 + i]), y = *reinterpret_cast<const Tr*>(&P.arcs2[a2 + i]);
      if (x.ilabel != y.ilabel || x.olabel != y.olabel || !iso_approx(x.weight, y.weight, P.delta)) { flags |= kIsoFail; continue; }
      if (i > 0) {
        const Tr p = *reinterpret_cast<const Tr*>(&P.arcs1[a1 + i - 1]);
        if (p.ilabel == x.ilabel && p.olabel == x.olabel && iso_approx(p.weight, x.weight, P.delta)) flags |= kIsoNonDet;
      }
      const uint32_t old = atomicCAS(&P.pair[x.nextstate], kNone, y.nextstate);
      if (old == kNone) {
        const uint32_t t = atomicAdd(&P.ctl[0], 1u);
        P.q1[t] = x.nextstate; P.q2[t] = y.nextstate;
      } else if (old != y.nextstate) {
        flags |= kIsoFail;
      }
    }
  }
  if (flags) atomicOr(&P.ctl[1], flags);
}

// arcs of `f` sorted per state by (ilabel, olabel, weight, nextstate): stable LSD passes over an index permutation
void sorted_rows(const DevFst& f, DevBuf<Tr>& out, cudaStream_t s) {
  const uint32_t a = f.num_arcs, n = f.num_states;
  out.reserve_discard(a ? a : 1);
  if (!a) return;
  DevBuf<uint32_t> src(s, a), idx_a(s, a), idx_b(s, a);
  DevBuf<unsigned long long> k_in(s, a), k_out(s, a);
  DevBuf<uint8_t> tmp(s);
  k_iso_src<<<blocks_for(n), kThreads, 0, s>>>(f.offsets.p, n, src.p);
  k_iso_iota<<<blocks_for(a), kThreads, 0, s>>>(idx_a.p, a);
  int state_bits = 1;
  while ((1ull << state_bits) < (unsigned long long)n) state_bits++;
  uint32_t* in = idx_a.p;
  uint32_t* outp = idx_b.p;
  for (int pass = 0; pass < 3; pass++) {
    k_iso_keys<<<blocks_for(a), kThreads, 0, s>>>(f.arcs.p, src.p, in, a, pass, k_in.p);
    sort_pairs_u64_u32(k_in.p, k_out.p, in, outp, a, pass == 2 ? state_bits : 64, tmp, s);
    std::swap(in, outp);
  }
  k_iso_gather<<<blocks_for(a), kThreads, 0, s>>>(f.arcs.p, in, a, out.p);
}

}  // namespace

// 1 = isomorphic, 0 = not isomorphic, -1 = undecided (a check failed after rows with equal neighbouring arcs were seen:
// the reference's answer — false or its non-determinism error — depends on the visiting order; ask the host).
int isomorphic_device(const DevFst& a, const DevFst& b, float delta, cudaStream_t s) {
  if (!a.has_start && !b.has_start) return 1;  // isomorphic.rs:57-63
  if (!a.has_start || !b.has_start) return 0;
  DevBuf<Tr> rows1(s), rows2(s);
  sorted_rows(a, rows1, s);
  sorted_rows(b, rows2, s);
  const uint32_t n1 = a.num_states;
  DevBuf<uint32_t> pair(s, n1), q1(s, n1), q2(s, n1), ctl(s, 2);
  B200_CUDA(cudaMemsetAsync(pair.p, 0xFF, (size_t)n1 * 4, s));
  const uint32_t init[2] = {1u, 0u};
  B200_CUDA(cudaMemcpyAsync(ctl.p, init, 8, cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemcpyAsync(q1.p, &a.start, 4, cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemcpyAsync(q2.p, &b.start, 4, cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemcpyAsync(pair.p + a.start, &b.start, 4, cudaMemcpyHostToDevice, s));
  IsoParams P{a.offsets.p, rows1.p, a.finals.p, b.offsets.p, rows2.p, b.finals.p, pair.p, q1.p, q2.p, ctl.p, delta};
  uint32_t lo = 0, hi = 1, h[2] = {1, 0};
  while (lo < hi) {
    k_iso_level<<<blocks_for((size_t)(hi - lo) * 32), kThreads, 0, s>>>(P, lo, hi);
    B200_CUDA(cudaMemcpyAsync(h, ctl.p, 8, cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaStreamSynchronize(s));
    if (h[1] & kIsoFail) return (h[1] & kIsoNonDet) ? -1 : 0;
    lo = hi; hi = h[0];
  }
  return 1;
}

}  // namespace b200
