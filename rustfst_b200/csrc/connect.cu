// connect.cu — trim (accessible AND coaccessible) with order-preserving renumbering, on device.
//
// Replaces (paths relative to /root/reference):
//   rustfst/src/algorithms/connect.rs:51-189        connect + ConnectVisitor (Tarjan DFS -> access/coaccess flags)
//   rustfst/src/algorithms/dfs_visit.rs:97-187      the sequential DFS driver
//   rustfst/src/fst_impls/vector_fst/mutable_fst.rs:132-189  del_states (stable compaction of states and arcs)
//
// The DFS only serves to compute two set-valued facts that do not depend on the visiting order — which states
// are reachable from the start and which can reach a final state — so they are recomputed as two frontier BFS
// sweeps (forward over the CSR, backward over a reverse CSR built with one histogram + one scatter), and
// del_states becomes two prefix sums (state keep-flags, surviving out-degrees) plus one gather.
#include "algos.h"
#include "coop_utils.cuh"

namespace cg = cooperative_groups;

namespace b200 {
namespace {
using namespace coop;

__global__ void k_rev_degree(const uint32_t* __restrict__ off, const Tr* __restrict__ arcs, uint32_t n,
                             uint32_t* __restrict__ rdeg) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  for (uint32_t i = off[s]; i < off[s + 1]; i++) atomicAdd(&rdeg[__ldg(&arcs[i].nextstate)], 1u);
}
__global__ void k_rev_fill(const uint32_t* __restrict__ off, const Tr* __restrict__ arcs, uint32_t n,
                           const uint32_t* __restrict__ roff, uint32_t* __restrict__ cursor,
                           uint32_t* __restrict__ rsrc) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  for (uint32_t i = off[s]; i < off[s + 1]; i++) {
    uint32_t t = __ldg(&arcs[i].nextstate);
    rsrc[roff[t] + atomicAdd(&cursor[t], 1u)] = s;
  }
}
__global__ void k_seed_finals(const float* __restrict__ fin, uint32_t n, uint32_t* __restrict__ mark,
                              uint32_t* __restrict__ frontier, uint32_t* __restrict__ count) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  bool f = s < n && fin[s] != w_zero();
  if (s < n) mark[s] = f ? 1u : 0u;
  // warp-aggregated append
  uint32_t m = __ballot_sync(0xFFFFFFFFu, f);
  if (!m) return;
  uint32_t lane = threadIdx.x & 31, leader = __ffs(m) - 1, basep = 0;
  if (lane == leader) basep = atomicAdd(count, __popc(m));
  basep = __shfl_sync(0xFFFFFFFFu, basep, leader);
  if (f) frontier[basep + __popc(m & ((1u << lane) - 1u))] = s;
}
// One BFS level, one WARP per frontier vertex (lanes stride over its adjacency list).  kForward: neighbours are
// arcs[i].nextstate; otherwise adj[i] (reverse CSR sources).  Vertices with more than kHeavyDegree neighbours (hub
// states: superfinal states, the root of an n-best tree) are only listed; k_bfs_heavy then spreads each of them over
// 64 CTAs.  count[0] = size of the next frontier, count[1] = number of heavy vertices of this level.
constexpr uint32_t kHeavyDegree = 4096;
constexpr uint32_t kHeavyCtas = 64;
template <bool kForward>
__device__ __forceinline__ void bfs_visit(const Tr* __restrict__ arcs, const uint32_t* __restrict__ adj, uint32_t k,
                                          uint32_t* __restrict__ mark, uint32_t* __restrict__ fout,
                                          uint32_t* __restrict__ count) {
  const uint32_t u = kForward ? __ldg(&arcs[k].nextstate) : __ldg(&adj[k]);
  if (mark[u] == 0 && atomicExch(&mark[u], 1u) == 0) fout[atomicAdd(count, 1u)] = u;
}
template <bool kForward>
__global__ void __launch_bounds__(kThreads)
k_bfs_level(const uint32_t* __restrict__ off, const Tr* __restrict__ arcs, const uint32_t* __restrict__ adj,
            const uint32_t* __restrict__ fin_in, uint32_t n_in, uint32_t* __restrict__ mark,
            uint32_t* __restrict__ fout, uint32_t* __restrict__ count, uint32_t* __restrict__ heavy) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_in) return;
  const uint32_t v = fin_in[w];
  const uint32_t b = off[v], e = off[v + 1];
  if (e - b > kHeavyDegree) {
    if (lane == 0) heavy[atomicAdd(&count[1], 1u)] = v;
    return;
  }
  for (uint32_t k = b + lane; k < e; k += 32) bfs_visit<kForward>(arcs, adj, k, mark, fout, count);
}
template <bool kForward>
__global__ void __launch_bounds__(kThreads)
k_bfs_heavy(const uint32_t* __restrict__ off, const Tr* __restrict__ arcs, const uint32_t* __restrict__ adj,
            const uint32_t* __restrict__ heavy, uint32_t* __restrict__ mark, uint32_t* __restrict__ fout,
            uint32_t* __restrict__ count) {
  const uint32_t v = heavy[blockIdx.y];
  const uint32_t b = off[v], e = off[v + 1];
  for (uint32_t k = b + blockIdx.x * blockDim.x + threadIdx.x; k < e; k += gridDim.x * blockDim.x)
    bfs_visit<kForward>(arcs, adj, k, mark, fout, count);
}
__global__ void k_keep_flags(const uint32_t* __restrict__ access, const uint32_t* __restrict__ coaccess, uint32_t n,
                             uint32_t* __restrict__ keep) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > n) return;
  keep[s] = (s < n && (access ? access[s] != 0 : true) && coaccess[s] != 0) ? 1u : 0u;
}
__global__ void k_new_degree(const uint32_t* __restrict__ off, const Tr* __restrict__ arcs, uint32_t n,
                             const uint32_t* __restrict__ keep, const uint32_t* __restrict__ new_id, uint32_t n_keep,
                             uint32_t* __restrict__ ndeg) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s == n) { ndeg[n_keep] = 0; return; }
  if (s > n || !keep[s]) return;
  uint32_t c = 0;
  for (uint32_t i = off[s]; i < off[s + 1]; i++) c += keep[__ldg(&arcs[i].nextstate)];
  ndeg[new_id[s]] = c;
}
__global__ void k_compact(const uint32_t* __restrict__ off, const Tr* __restrict__ arcs,
                          const float* __restrict__ fin, uint32_t n, const uint32_t* __restrict__ keep,
                          const uint32_t* __restrict__ new_id, const uint32_t* __restrict__ noff,
                          Tr* __restrict__ narcs, float* __restrict__ nfin) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n || !keep[s]) return;
  uint32_t ns = new_id[s];
  nfin[ns] = fin[s];
  uint32_t o = noff[ns];
  for (uint32_t i = off[s]; i < off[s + 1]; i++) {
    int4 v = __ldg(reinterpret_cast<const int4*>(&arcs[i]));
    uint32_t t = (uint32_t)v.w;
    if (keep[t]) {
      v.w = (int)new_id[t];
      *reinterpret_cast<int4*>(&narcs[o++]) = v;
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Persistent trim for a freshly composed FST.  The BFS that built it numbered the states wave by wave
// (wave k = ids [wave_lo[k], wave_lo[k+1])) and every state is accessible, so coaccessibility can be PULLED in one
// reverse sweep over the waves: a state is coaccessible iff it is final or one of its arcs reaches a coaccessible
// state; arcs into later waves see final values, arcs into the same or an earlier wave (cycles) may see stale ones,
// in which case the sweep is repeated until nothing changes (monotone fixpoint).  No reverse CSR, no atomics.
// Then del_states: slice-local scans of the keep flags and of the surviving out-degrees (global offsets are
// slice prefix + local), and one gather.  n_waves + 3 grid barriers per sweep, one launch.
struct TrimParams {
  const uint32_t* off; const Tr* arcs; const float* fin; uint32_t n;
  const uint32_t* wave_lo; uint32_t n_waves;
  uint8_t* coacc;
  uint32_t* id_loc;      // slice-local new id of every state
  uint32_t* deg_loc;     // slice-local new arc offset of every KEPT state (indexed by old id)
  uint32_t* part_keep; uint32_t* part_deg;
  uint32_t* noff; Tr* narcs; float* nfin;
  uint32_t* ctl;         // [0] changed flag per sweep, [1] back-arc seen at an undecided state, [2] n_keep, [3] a_keep, [4] sweeps,
                         // [5] barrier, [6] some arc stays inside its wave or goes back (first-sweep counts not final)
  const unsigned long long* tuples; uint32_t* ntag;   // optional: s1 component of the packed tuple, gathered
  uint32_t n_starts; uint32_t* start_map;             // optional: new ids of input states 0..n_starts-1
  uint32_t light_barrier;  // 1: arrival-counter barrier of coop_utils.cuh on ctl[5] instead of cg::grid_group::sync()
  // optional (compose_ws.cu): the arcs still sit in provisional (wave, warp) runs; `next` is the dense array of their
  // resolved next states in canonical order (all the sweeps need), st_first[s] = (run, index) of the first arc of state s
  const uint32_t* next; const Tr* prov; const uint2* st_first; const uint32_t* run_src; const uint32_t* run_cnt;
  const uint32_t* run_dst; uint32_t n_runs;  // canonical index of every run's first arc (n_runs + 1 entries)
};

// A state with thousands of arcs (a start state that fans out, a hub) would keep ONE thread busy for milliseconds in the
// sweep and in the gather; such states are set aside and walked by the whole CTA.
constexpr uint32_t kBigDegree = 512, kBigMax = 64;
#define TRIM_SYNC() do { if (P.light_barrier) grid_barrier(P.ctl + 5, bar_epoch); else grid.sync(); } while (0)
__global__ void __launch_bounds__(kCoopThreads, 2)
k_trim_coop(TrimParams P) {
  cg::grid_group grid = cg::this_grid();
  unsigned int bar_epoch = 0;
  __shared__ uint32_t s_warp[kCoopThreads / 32];
  __shared__ uint32_t s_big[kBigMax];  // states with more than kBigDegree arcs met by this CTA: handled by the whole CTA
  __shared__ uint32_t s_nbig, s_acc[2];
  extern __shared__ uint32_t s_dyn[];
  uint32_t* s_pref_id = s_dyn;
  uint32_t* s_pref_deg = s_dyn + gridDim.x + 1;
  const uint32_t G = gridDim.x, c = blockIdx.x, tid = threadIdx.x;
  const uint32_t gsize = G * kCoopThreads, gtid = c * kCoopThreads + tid;

  // ---- coaccessibility: reverse sweeps over the BFS waves
  const unsigned long long t_start = globaltimer_ns();
  uint32_t sweeps = 0;
  while (true) {
    for (uint32_t k = P.n_waves; k-- > 0;) {
      const uint32_t wlo = P.wave_lo[k], whi = P.wave_lo[k + 1];
      if (tid == 0) s_nbig = 0;
      __syncthreads();
      for (uint32_t s = wlo + gtid; s < whi; s += gsize) {
        uint8_t co = __ldcg(&P.coacc[s]);
        if (!co) {
          bool back = false;
          const uint32_t a_lo = P.off[s], a_hi = P.off[s + 1];
          if (sweeps == 0 && a_hi - a_lo > kBigDegree) {
            const uint32_t q = atomicAdd(&s_nbig, 1u);
            if (q < kBigMax) { s_big[q] = s; continue; }  // the CTA walks it below; beyond kBigMax: this thread does
          }
          if (sweeps == 0) {
            // first sweep: look at every arc (four at a time, look-ups in flight together) and count the coaccessible
            // targets — when no arc stays inside its wave or goes back (ctl[6]), that count is final and the degree
            // pass below does not have to read the arcs again
            uint32_t cnt = 0;
            for (uint32_t i = a_lo; i < a_hi; i += 4) {
              uint32_t t[4];
              uint8_t x[4];
#pragma unroll
              for (int q = 0; q < 4; q++) {
                const uint32_t k = min(i + q, a_hi - 1);
                t[q] = P.next ? __ldg(&P.next[k]) : __ldg(&P.arcs[k].nextstate);
              }
#pragma unroll
              for (int q = 0; q < 4; q++) x[q] = __ldcg(&P.coacc[t[q]]);
#pragma unroll
              for (int q = 0; q < 4; q++)
                if (i + q < a_hi) { if (t[q] < whi) back = true; cnt += x[q]; }
            }
            P.deg_loc[s] = cnt;
            co = (P.fin[s] != w_zero() || cnt) ? 1 : 0;
            if (back) P.ctl[6] = 1;
          } else {
            for (uint32_t i = a_lo; i < a_hi && !co; i++) {
              const uint32_t t = P.next ? __ldg(&P.next[i]) : __ldg(&P.arcs[i].nextstate);
              if (t < whi) back = true;  // same or earlier wave: value may still change in this sweep
              co = __ldcg(&P.coacc[t]);
            }
          }
          if (co) { P.coacc[s] = 1; if (sweeps > 0) P.ctl[0] = 1; }
          else if (back) P.ctl[1] = 1;
        }
      }
      __syncthreads();
      for (uint32_t q = 0; q < min(s_nbig, kBigMax); q++) {  // big states of this CTA and wave (first sweep only)
        const uint32_t s = s_big[q];
        const uint32_t a_lo = P.off[s], a_hi = P.off[s + 1];
        if (tid == 0) { s_acc[0] = 0; s_acc[1] = 0; }
        __syncthreads();
        uint32_t cnt = 0, back = 0;
        for (uint32_t i = a_lo + tid; i < a_hi; i += kCoopThreads) {
          const uint32_t t = P.next ? __ldg(&P.next[i]) : __ldg(&P.arcs[i].nextstate);
          if (t < whi) back = 1;
          cnt += __ldcg(&P.coacc[t]);
        }
        if (cnt) atomicAdd(&s_acc[0], cnt);
        if (back) s_acc[1] = 1;
        __syncthreads();
        if (tid == 0) {
          P.deg_loc[s] = s_acc[0];
          const bool co = P.fin[s] != w_zero() || s_acc[0];
          if (s_acc[1]) P.ctl[6] = 1;
          if (co) P.coacc[s] = 1;
          else if (s_acc[1]) P.ctl[1] = 1;
        }
        __syncthreads();
      }
      TRIM_SYNC();
    }
    sweeps++;
    // another sweep is needed only if some undecided state has an arc that was looked at too early
    const uint32_t back_seen = __ldcg(&P.ctl[1]), changed = __ldcg(&P.ctl[0]);
    TRIM_SYNC();
    if (!back_seen || (sweeps > 1 && !changed)) break;
    if (c == 0 && tid == 0) { P.ctl[0] = 0; P.ctl[1] = 0; }
    TRIM_SYNC();
  }

  const unsigned long long t_sweep = globaltimer_ns();
  // ---- new state ids: slice-local exclusive scan of the keep flags
  const uint32_t sc = ((P.n + G - 1) / G + kCoopThreads - 1) / kCoopThreads * kCoopThreads;
  const uint32_t s_begin = min(P.n, c * sc), s_end = min(P.n, s_begin + sc);
  {
    uint32_t run = 0;
    for (uint32_t s0 = s_begin; s0 < s_end; s0 += kCoopThreads) {
      const uint32_t s = s0 + tid;
      const uint32_t keep = (s < s_end) ? (uint32_t)__ldcg(&P.coacc[s]) : 0u;
      uint32_t tot;
      const uint32_t ex = cta_exclusive_scan(keep, s_warp, tot);
      if (s < s_end) P.id_loc[s] = keep ? run + ex : 0xFFFFFFFFu;  // all ones = state is dropped
      run += tot;
    }
    if (tid == 0) P.part_keep[c] = run;
  }
  TRIM_SYNC();
  cta_prefix_to_smem(P.part_keep, G, s_pref_id, s_warp);
  const uint32_t n_keep = s_pref_id[G];

  const unsigned long long t_ids = globaltimer_ns();
  // ---- surviving out-degrees: slice-local scan, indexed by old state id
  {
    const bool counted = __ldcg(&P.ctl[6]) == 0u;  // the first sweep's counts are final (no arc inside a wave or back)
    uint32_t run = 0;
    for (uint32_t s0 = s_begin; s0 < s_end; s0 += kCoopThreads) {
      const uint32_t s = s0 + tid;
      uint32_t d = 0;
      if (s < s_end && __ldcg(&P.coacc[s]) && counted) d = __ldcg(&P.deg_loc[s]);
      else if (s < s_end && __ldcg(&P.coacc[s])) {
        const uint32_t a_lo = P.off[s], a_hi = P.off[s + 1];
        for (uint32_t i = a_lo; i < a_hi; i += 4) {  // four arcs per round: the target look-ups fly together
          uint32_t t[4], x[4];
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const uint32_t k = min(i + q, a_hi - 1);
            t[q] = P.next ? __ldg(&P.next[k]) : __ldg(&P.arcs[k].nextstate);
          }
#pragma unroll
          for (int q = 0; q < 4; q++) x[q] = __ldcg(&P.id_loc[t[q]]);
#pragma unroll
          for (int q = 0; q < 4; q++) d += (i + q < a_hi && x[q] != 0xFFFFFFFFu) ? 1u : 0u;
        }
      }
      uint32_t tot;
      const uint32_t ex = cta_exclusive_scan(d, s_warp, tot);
      if (s < s_end) P.deg_loc[s] = run + ex;
      run += tot;
    }
    if (tid == 0) P.part_deg[c] = run;
  }
  TRIM_SYNC();
  cta_prefix_to_smem(P.part_deg, G, s_pref_deg, s_warp);
  const uint32_t a_keep = s_pref_deg[G];

  const unsigned long long t_deg = globaltimer_ns();
  // ---- gather (mutable_fst.rs:132-189: survivors keep their relative order, arcs into deleted states are dropped)
  if (tid == 0) s_nbig = 0;
  __syncthreads();
  for (uint32_t s = s_begin + tid; s < s_end; s += kCoopThreads) {
    if (!__ldcg(&P.coacc[s])) continue;
    const uint32_t ns = s_pref_id[c] + P.id_loc[s];
    uint32_t o = s_pref_deg[c] + P.deg_loc[s];
    P.nfin[ns] = P.fin[s];
    P.noff[ns] = o;
    if (P.ntag) P.ntag[ns] = (uint32_t)(P.tuples[s] & 0x7FFFFFFFull);
    const uint32_t a_lo = P.off[s], a_hi = P.off[s + 1];
    if (a_hi - a_lo > kBigDegree) {
      const uint32_t q = atomicAdd(&s_nbig, 1u);
      if (q < kBigMax) { s_big[q] = s; continue; }  // the CTA gathers it below
    }
    if (P.prov) {
      // the arcs of the state start at st_first[s] in a provisional run and continue at the start of the following
      // run(s); their resolved next states are next[a_lo .. a_hi) in canonical order
      const uint2 f = __ldg(&P.st_first[s]);
      uint32_t r = f.x, el = f.y, i = a_lo;
      while (i < a_hi) {
        const uint32_t rc = __ldg(&P.run_cnt[r]);
        if (el >= rc) { r++; el = 0; continue; }
        const uint32_t take = min(a_hi - i, rc - el);
        const Tr* __restrict__ src = P.prov + __ldg(&P.run_src[r]) + el;
        for (uint32_t k = 0; k < take; k += 4) {
          int4 v[4];
          uint32_t t[4], x[4];
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const uint32_t kk = min(k + q, take - 1);
            v[q] = __ldg(reinterpret_cast<const int4*>(&src[kk]));
            t[q] = __ldg(&P.next[i + kk]);
          }
#pragma unroll
          for (int q = 0; q < 4; q++) x[q] = __ldcg(&P.id_loc[t[q]]);
#pragma unroll
          for (int q = 0; q < 4; q++) {
            if (k + q < take && x[q] != 0xFFFFFFFFu) {
              v[q].w = (int)(s_pref_id[t[q] / sc] + x[q]);
              *reinterpret_cast<int4*>(&P.narcs[o++]) = v[q];
            }
          }
        }
        i += take; el += take;
      }
      continue;
    }
    for (uint32_t i = a_lo; i < a_hi; i += 4) {
      int4 v[4];
      uint32_t x[4];
#pragma unroll
      for (int q = 0; q < 4; q++) v[q] = __ldg(reinterpret_cast<const int4*>(&P.arcs[min(i + q, a_hi - 1)]));
#pragma unroll
      for (int q = 0; q < 4; q++) x[q] = __ldcg(&P.id_loc[(uint32_t)v[q].w]);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        if (i + q < a_hi && x[q] != 0xFFFFFFFFu) {
          v[q].w = (int)(s_pref_id[(uint32_t)v[q].w / sc] + x[q]);
          *reinterpret_cast<int4*>(&P.narcs[o++]) = v[q];
        }
      }
    }
  }
  __syncthreads();
  for (uint32_t q = 0; q < min(s_nbig, kBigMax); q++) {  // big states of this slice: order-preserving compaction by the CTA
    const uint32_t s = s_big[q];
    const uint32_t a_lo = P.off[s], a_hi = P.off[s + 1];
    uint32_t o = s_pref_deg[c] + P.deg_loc[s];
    uint32_t r_lo = 0, r_hi = 0;
    if (P.prov) { r_lo = __ldg(&P.st_first[s]).x; r_hi = P.n_runs; }
    for (uint32_t i0 = a_lo; i0 < a_hi; i0 += kCoopThreads) {
      const uint32_t i = i0 + tid;
      int4 v = make_int4(0, 0, 0, 0);
      uint32_t t = 0, x = 0xFFFFFFFFu;
      if (i < a_hi) {
        if (P.prov) {
          // canonical arc i lives in the run r with run_dst[r] <= i < run_dst[r + 1]
          uint32_t lo = r_lo, hi = r_hi;
          while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(&P.run_dst[mid]) <= i) lo = mid; else hi = mid; }
          v = __ldg(reinterpret_cast<const int4*>(&P.prov[__ldg(&P.run_src[lo]) + (i - __ldg(&P.run_dst[lo]))]));
          t = __ldg(&P.next[i]);
        } else {
          v = __ldg(reinterpret_cast<const int4*>(&P.arcs[i]));
          t = (uint32_t)v.w;
        }
        x = __ldcg(&P.id_loc[t]);
      }
      const bool keep = x != 0xFFFFFFFFu;
      uint32_t tot;
      const uint32_t ex = cta_exclusive_scan(keep ? 1u : 0u, s_warp, tot);
      if (keep) {
        v.w = (int)(s_pref_id[t / sc] + x);
        *reinterpret_cast<int4*>(&P.narcs[o + ex]) = v;
      }
      o += tot;
    }
  }
  if (P.start_map)
    for (uint32_t i = gtid; i < P.n_starts; i += gsize)
      P.start_map[i] = __ldcg(&P.coacc[i]) ? s_pref_id[i / sc] + __ldcg(&P.id_loc[i]) : 0xFFFFFFFFu;
  if (c == 0 && tid == 0) {
    P.noff[n_keep] = a_keep; P.ctl[2] = n_keep; P.ctl[3] = a_keep; P.ctl[4] = sweeps;
    // phase times of CTA 0 in microseconds: sweep, ids, degrees, gather (its own share only)
    P.ctl[8] = (uint32_t)((t_sweep - t_start) / 1000); P.ctl[9] = (uint32_t)((t_ids - t_sweep) / 1000);
    P.ctl[10] = (uint32_t)((t_deg - t_ids) / 1000); P.ctl[11] = (uint32_t)((globaltimer_ns() - t_deg) / 1000);
  }
}

#undef TRIM_SYNC
// One BFS level from the host: launches the level kernel (+ the hub kernel when the level contains hub vertices) and
// returns the size of the next frontier.
template <bool kForward>
uint32_t bfs_step(const uint32_t* off, const Tr* arcs, const uint32_t* adj, const uint32_t* fin_in, uint32_t nf,
                  uint32_t* mark, uint32_t* fout, uint32_t* count, uint32_t* heavy, uint64_t* nl, cudaStream_t s) {
  B200_CUDA(cudaMemsetAsync(count, 0, 8, s));
  k_bfs_level<kForward><<<blocks_for((size_t)nf * 32), kThreads, 0, s>>>(off, arcs, adj, fin_in, nf, mark, fout, count, heavy);
  (*nl)++;
  uint32_t c[2];
  B200_CUDA(cudaMemcpyAsync(c, count, 8, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  for (uint32_t h0 = 0; h0 < c[1]; h0 += 32768) {
    const uint32_t nh = std::min<uint32_t>(32768, c[1] - h0);
    k_bfs_heavy<kForward><<<dim3(kHeavyCtas, nh), kThreads, 0, s>>>(off, arcs, adj, heavy + h0, mark, fout, count);
    (*nl)++;
  }
  if (c[1]) c[0] = read_u32(count, s);
  return c[0];
}

}  // namespace

DevFst connect_device(const DevFst& in, bool assume_accessible, uint64_t* launches, cudaStream_t s) {
  uint64_t nl = 0;
  DevFst out(s);
  out.props = props::after_connect(in.props);
  const uint32_t n = in.num_states;
  // dfs_visit returns at once without a start state (dfs_visit.rs:103-109): nothing is accessible
  if (n == 0 || !in.has_start) {
    out.offsets.reserve_discard(1);
    B200_CUDA(cudaMemsetAsync(out.offsets.p, 0, 4, s));
    return out;
  }
  DevBuf<uint8_t> scan_tmp(s);
  DevBuf<uint32_t> fa(s, n), fb(s, n), count(s, 2), heavy(s, n);
  uint32_t* fin_p = fa.p;
  uint32_t* fout_p = fb.p;

  // ---- forward reachability from the start state
  DevBuf<uint32_t> access(s);
  if (!assume_accessible) {
    access.reserve_discard(n);
    B200_CUDA(cudaMemsetAsync(access.p, 0, (size_t)n * 4, s));
    uint32_t one = 1, st0 = in.start;
    B200_CUDA(cudaMemcpyAsync(access.p + in.start, &one, 4, cudaMemcpyHostToDevice, s));
    B200_CUDA(cudaMemcpyAsync(fin_p, &st0, 4, cudaMemcpyHostToDevice, s));
    B200_CUDA(cudaStreamSynchronize(s));
    uint32_t nf = 1;
    while (nf) {
      nf = bfs_step<true>(in.offsets.p, in.arcs.p, nullptr, fin_p, nf, access.p, fout_p, count.p, heavy.p, &nl, s);
      std::swap(fin_p, fout_p);
    }
  }

  // ---- reverse CSR (histogram -> scan -> scatter) and backward reachability from the final states
  DevBuf<uint32_t> roff(s, (size_t)n + 1), cursor(s, n), rsrc(s, in.num_arcs ? in.num_arcs : 1), coaccess(s, n);
  B200_CUDA(cudaMemsetAsync(roff.p, 0, ((size_t)n + 1) * 4, s));
  B200_CUDA(cudaMemsetAsync(cursor.p, 0, (size_t)n * 4, s));
  k_rev_degree<<<blocks_for(n), kThreads, 0, s>>>(in.offsets.p, in.arcs.p, n, roff.p);
  exclusive_sum_u32(roff.p, roff.p, (size_t)n + 1, scan_tmp, s);
  k_rev_fill<<<blocks_for(n), kThreads, 0, s>>>(in.offsets.p, in.arcs.p, n, roff.p, cursor.p, rsrc.p);
  B200_CUDA(cudaMemsetAsync(count.p, 0, 4, s));
  k_seed_finals<<<blocks_for(n), kThreads, 0, s>>>(in.finals.p, n, coaccess.p, fin_p, count.p);
  nl += 4;
  uint32_t nf = read_u32(count.p, s);
  while (nf) {
    nf = bfs_step<false>(roff.p, nullptr, rsrc.p, fin_p, nf, coaccess.p, fout_p, count.p, heavy.p, &nl, s);
    std::swap(fin_p, fout_p);
  }

  // ---- del_states: stable compaction of states, then of the arcs that stay inside the kept set
  DevBuf<uint32_t> keep(s, (size_t)n + 1), new_id(s, (size_t)n + 1);
  k_keep_flags<<<blocks_for((size_t)n + 1), kThreads, 0, s>>>(assume_accessible ? nullptr : access.p, coaccess.p, n,
                                                              keep.p);
  exclusive_sum_u32(keep.p, new_id.p, (size_t)n + 1, scan_tmp, s);
  nl += 2;
  uint32_t n_keep = read_u32(new_id.p + n, s);
  out.offsets.reserve_discard((size_t)n_keep + 1);
  out.finals.reserve_discard(n_keep);
  k_new_degree<<<blocks_for((size_t)n + 1), kThreads, 0, s>>>(in.offsets.p, in.arcs.p, n, keep.p, new_id.p, n_keep,
                                                              out.offsets.p);
  exclusive_sum_u32(out.offsets.p, out.offsets.p, (size_t)n_keep + 1, scan_tmp, s);
  nl += 2;
  uint32_t a_keep = read_u32(out.offsets.p + n_keep, s);
  out.arcs.reserve_discard(a_keep);
  k_compact<<<blocks_for(n), kThreads, 0, s>>>(in.offsets.p, in.arcs.p, in.finals.p, n, keep.p, new_id.p,
                                               out.offsets.p, out.arcs.p, out.finals.p);
  nl++;
  out.num_states = n_keep;
  out.num_arcs = a_keep;
  // start remap (mutable_fst.rs:176-183)
  uint32_t kv[2];
  B200_CUDA(cudaMemcpyAsync(&kv[0], keep.p + in.start, 4, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaMemcpyAsync(&kv[1], new_id.p + in.start, 4, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  out.has_start = kv[0] != 0;
  out.start = kv[0] ? kv[1] : 0;
  if (launches) *launches += nl;
  return out;
}

DevFst connect_waves_device(const DevFst& in, const uint32_t* d_wave_lo, uint32_t n_waves, uint64_t* launches,
                            cudaStream_t s, const TrimExtras* extras, const ProvArcs* pa) {
  DevFst out(s);
  out.props = props::after_connect(in.props);
  const uint32_t n = in.num_states;
  if (launches) *launches = 0;
  if (in.has_start && in.start != 0) throw FstError("connect_waves_device expects the start state to be state 0");
  if (n == 0 || !in.has_start || n_waves == 0) {
    if (extras && extras->out_tag) {
      extras->out_tag->reserve_discard(1);
      extras->out_start_map->reserve_discard(extras->n_starts ? extras->n_starts : 1);
      B200_CUDA(cudaMemsetAsync(extras->out_start_map->p, 0xFF, (size_t)(extras->n_starts ? extras->n_starts : 1) * 4, s));
    }
    out.offsets.reserve_discard(1);
    B200_CUDA(cudaMemsetAsync(out.offsets.p, 0, 4, s));
    return out;
  }
  DevBuf<uint8_t> coacc(s, n);
  DevBuf<uint32_t> id_loc(s, n), deg_loc(s, n), parts(s, 2 * 2049), ctl(s, 12);
  B200_CUDA(cudaMemsetAsync(coacc.p, 0, n, s));
  B200_CUDA(cudaMemsetAsync(ctl.p, 0, 48, s));
  out.offsets.reserve_discard((size_t)n + 1);
  out.finals.reserve_discard(n);
  out.arcs.reserve_discard(in.num_arcs ? in.num_arcs : 1);
  TrimParams P{};
  P.off = in.offsets.p; P.arcs = in.arcs.p; P.fin = in.finals.p; P.n = n;
  P.wave_lo = d_wave_lo; P.n_waves = n_waves;
  P.coacc = coacc.p; P.id_loc = id_loc.p; P.deg_loc = deg_loc.p;
  P.part_keep = parts.p; P.part_deg = parts.p + 2049;
  P.noff = out.offsets.p; P.narcs = out.arcs.p; P.nfin = out.finals.p; P.ctl = ctl.p;
  P.light_barrier = std::getenv("B200_TRIM_CG") ? 0u : 1u;
  if (pa) {
    P.next = pa->next; P.prov = pa->prov; P.st_first = pa->st_first; P.run_src = pa->run_src; P.run_cnt = pa->run_cnt;
    P.run_dst = pa->run_dst; P.n_runs = pa->n_runs;
  }
  if (extras && extras->out_tag) {
    extras->out_tag->reserve_discard(n);
    extras->out_start_map->reserve_discard(extras->n_starts);
    P.tuples = extras->tuples; P.ntag = extras->out_tag->p;
    P.n_starts = extras->n_starts; P.start_map = extras->out_start_map->p;
  }
  int per_sm = 0;
  size_t dyn = 2 * 2049 * sizeof(uint32_t);
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trim_coop, kCoopThreads, dyn));
  if (per_sm < 1) throw FstError("cooperative trim kernel does not fit on the device");
  int grid = std::min(sm_count() * per_sm, 2047);
  dyn = 2 * ((size_t)grid + 1) * sizeof(uint32_t);
  void* args[] = {(void*)&P};
  B200_CUDA(cudaLaunchCooperativeKernel((void*)k_trim_coop, dim3(grid), dim3(kCoopThreads), args, dyn, s));
  if (launches) *launches = 1;
  uint32_t h[12];
  B200_CUDA(cudaMemcpyAsync(h, ctl.p, 48, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  if (std::getenv("B200_COOP_TRACE"))
    std::fprintf(stderr, "[trim] %d CTAs, %u sweeps: sweep %u us, ids %u us, degrees %u us, gather %u us\n", grid, h[4], h[8],
                 h[9], h[10], h[11]);
  out.num_states = h[2]; out.num_arcs = h[3];
  // start remap (mutable_fst.rs:176-183): the start state is id 0 of the composed FST
  uint8_t keep0 = 0;
  B200_CUDA(cudaMemcpyAsync(&keep0, coacc.p + in.start, 1, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  out.has_start = keep0 != 0;  // state 0 keeps id 0 when it survives
  out.start = 0;
  return out;
}

// ---------------------------------------------------------------------------------------------------------------
// tr_sort on the device (rustfst/src/algorithms/tr_sort.rs:51-62: stable per-state sort by ilabel or olabel).
// One stable LSD radix sort of (state << 32 | label) over all arcs + one gather; state order and, inside equal
// labels, the original arc order are preserved exactly as Vec::sort_by does.
namespace {
__global__ void k_sort_keys(const uint32_t* __restrict__ off, const Tr* __restrict__ arcs, uint32_t n, int by_ilabel,
                            unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  for (uint32_t e = off[s]; e < off[s + 1]; e++) {
    const Label l = by_ilabel ? __ldg(&arcs[e].ilabel) : __ldg(&arcs[e].olabel);
    keys[e] = ((unsigned long long)s << 32) | l;
    vals[e] = e;
  }
}
__global__ void k_gather_arcs(const Tr* __restrict__ in, const uint32_t* __restrict__ perm, uint32_t a,
                              Tr* __restrict__ out) {
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a) return;
  *reinterpret_cast<int4*>(&out[e]) = __ldg(reinterpret_cast<const int4*>(&in[perm[e]]));
}
}  // namespace

void tr_sort_device(DevFst& f, bool ilabel, cudaStream_t s) {
  const uint32_t n = f.num_states, a = f.num_arcs;
  if (a == 0) return;
  DevBuf<unsigned long long> k_in(s, a), k_out(s, a);
  DevBuf<uint32_t> v_in(s, a), perm(s, a);
  DevBuf<Tr> sorted(s, a);
  DevBuf<uint8_t> tmp(s);
  k_sort_keys<<<blocks_for(n), kThreads, 0, s>>>(f.offsets.p, f.arcs.p, n, ilabel ? 1 : 0, k_in.p, v_in.p);
  sort_pairs_u64_u32(k_in.p, k_out.p, v_in.p, perm.p, a, 64, tmp, s);
  k_gather_arcs<<<blocks_for(a), kThreads, 0, s>>>(f.arcs.p, perm.p, a, sorted.p);
  f.arcs = std::move(sorted);
  f.columns.reset();  // derived label columns (compose matcher) describe the old arc order
  f.props = props::after_tr_sort(f.props, ilabel) & props::kTrinary;
  B200_CUDA(cudaStreamSynchronize(s));
}

}  // namespace b200
