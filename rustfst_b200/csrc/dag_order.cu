// dag_order.cu — the processing order of the reference's TopOrderQueue for an ACYCLIC machine, computed on the device.
//
// Replaces, for acyclic inputs (paths relative to /root/reference):
//   rustfst/src/algorithms/top_sort.rs:12-61        TopOrderVisitor: order[s] = position of s in REVERSE DFS FINISH order
//   rustfst/src/algorithms/dfs_visit.rs:97-187      the sequential DFS: root = start, then every still-white state
//                                                   0, 1, 2, ...; arcs in stored order
//   rustfst/src/algorithms/queues/top_order_queue.rs:20-42, auto_queue.rs:39-44
//
// The order decides which of several tied parents a shortest path takes, so it has to be THE order of that DFS, not
// just any topological order.  The DFS is sequential, but on a DAG its result has a closed form:
//
//   * Give the machine a virtual root R with arcs to [start, 0, 1, ..., n-1] in that order (the DFS roots in the order
//     dfs_visit tries them) and write a path from R as the string of arc positions it takes.  On a DAG the DFS tree path
//     of every state is the LEXICOGRAPHICALLY SMALLEST path from R to it: if the smallest path P and the tree path Q
//     part at a state x (P through arc i, Q through arc j > i), the DFS explores arc i of x first and does not return
//     to arc j before everything reachable through arc i — the state in question included — has been visited (no arc
//     into a grey state exists on a DAG), so Q cannot be the tree path.
//   * Hence parent(v) = the in-arc (u, pos) that minimises path(u) . pos, and two candidates are compared without
//     materialising the strings: lift the deeper source to the depth of the other (binary lifting), and either one
//     source is an ancestor of the other — compare the arc taken below it with the candidate's own position — or the
//     sources part at their lowest common ancestor — compare the positions of its two children.  (Tried and
//     dropped: carrying the first five positions of every path as a packed key so that two loads decide most
//     comparisons — 7.4 vs 7.1 ms on C4: the comparisons are not what the kernel waits for.)
//   * States are processed in Kahn levels (a state after all its predecessors), one grid barrier per level: every
//     processed state PUSHES its candidacy to its successors with a compare-and-swap loop on a packed
//     (source, position) word, then decrements their in-degree; no reverse CSR is needed.  A level starts by recording
//     depth and ancestor table of its own states (their parents are final), then a second barrier, then the candidacies:
//     everything a comparison reads was written before a grid barrier.
//   * Reverse finish order of a tree = visit a node, then its children from RIGHT to LEFT:
//       order(child_i) = order(parent) + 1 + sum of the subtree sizes of the children to the right of child_i.
//     Subtree sizes are accumulated in one reverse sweep over the levels, the orders in one forward sweep.
//
// Cost: O((n + a) log depth) work in 4 x (#Kahn levels) grid barriers.  Very deep machines (more levels than
// kMaxLevels, e.g. a long chain) are left to the host DFS of queue_plan.cpp, which is fast exactly there.
#include <cooperative_groups.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "algos.h"
#include "coop_utils.cuh"

namespace b200 {
namespace {
using namespace coop;

constexpr uint32_t kDagThreads = 512;
constexpr uint32_t kSub = 8;             // lanes that share one state (its arcs are strided over them)
constexpr uint32_t kMaxLevels = 1u << 16;
enum DagStatus : uint32_t { kDagCyclic = 1, kDagTooDeep = 2 };

struct DagParams {
  const uint32_t* off; const Tr* arcs; uint32_t n; uint32_t start;
  uint32_t* indeg;            // in-arcs of every state that are still to be processed
  unsigned long long* best;   // n + 1: (tree parent << 32 | position of the tree arc in the parent's list); node n = R
  uint32_t* depth;            // n + 1, depth[R] = 0
  uint32_t* up;               // up[j * (n + 1) + v] = 2^j-th ancestor of v, written for 2^j <= depth[v]
  uint32_t* lev_nodes;        // states in Kahn level order; level k = lev_nodes[lev_off[k] .. lev_off[k + 1])
  uint32_t* lev_off;
  uint32_t* sizes;            // subtree sizes
  uint32_t* order;            // result
  uint32_t* ctl;              // [0..2] rotating level counters, [3] #levels, [4] status, [5] barrier of k_dag_tree, [6] #states processed,
                              // [7] barrier of k_dag_orders
};

__device__ __forceinline__ unsigned long long pack_cand(uint32_t node, uint32_t pos) {
  return ((unsigned long long)node << 32) | pos;
}
__device__ __forceinline__ uint32_t up_at(const DagParams& P, uint32_t j, uint32_t v) {
  return __ldcg(&P.up[(size_t)j * (P.n + 1) + v]);
}
__device__ __forceinline__ uint32_t pos_in_parent(const DagParams& P, uint32_t v) {
  return (uint32_t)__ldcg(&P.best[v]);
}
__device__ __forceinline__ uint32_t lift(const DagParams& P, uint32_t v, uint32_t k) {
  while (k) { const uint32_t j = 31u - __clz(k); v = up_at(P, j, v); k -= 1u << j; }
  return v;
}
// path(a) . pa  <  path(b) . pb   (lexicographic; a and b are processed states or R, a at depth da)
__device__ __forceinline__ bool cand_less(const DagParams& P, uint32_t a, uint32_t pa, uint32_t da, uint32_t b, uint32_t pb) {
  if (a == b) return pa < pb;
  const uint32_t db = __ldcg(&P.depth[b]);
  uint32_t x = a, y = b, d = da;
  if (da > db) {  // is b an ancestor of a?  then a's path leaves b through the arc towards a
    x = lift(P, a, da - db - 1);
    const uint32_t px = up_at(P, 0, x);
    if (px == b) return pos_in_parent(P, x) < pb;
    x = px; d = db;
  } else if (db > da) {
    y = lift(P, b, db - da - 1);
    const uint32_t py = up_at(P, 0, y);
    if (py == a) return pa < pos_in_parent(P, y);
    y = py;
  }
  if (x == y) return false;  // cannot happen on a DAG (one candidate's path would run through the target)
  // x != y at the same depth d >= 1: climb to the children of their lowest common ancestor
  for (int j = 31 - __clz(d); j >= 0; j--) {
    if ((1u << j) > d) continue;
    const uint32_t ux = up_at(P, j, x), uy = up_at(P, j, y);
    if (ux != uy) { x = ux; y = uy; d -= 1u << j; }
  }
  return pos_in_parent(P, x) < pos_in_parent(P, y);
}

// in-degrees, initial candidates (R, position of the virtual arc: start first, then the states by id), sizes
__global__ void k_dag_init(DagParams P) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < P.n) {
    P.best[v] = pack_cand(P.n, v == P.start ? 0u : v + 1u);
    P.sizes[v] = 1u;
  } else if (v == P.n) {
    P.best[v] = pack_cand(P.n, 0u);
    P.depth[v] = 0u;
    for (int k = 0; k < 8; k++) P.ctl[k] = 0u;
  }
}
__global__ void k_dag_indeg(const Tr* __restrict__ arcs, uint32_t a, uint32_t* __restrict__ indeg) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < a) atomicAdd(&indeg[__ldg(&arcs[e].nextstate)], 1u);
}
__global__ void k_dag_seed(DagParams P) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  const bool ready = v < P.n && P.indeg[v] == 0u;
  const uint32_t m = __ballot_sync(0xFFFFFFFFu, ready);
  if (!m) return;
  const uint32_t lane = threadIdx.x & 31u, leader = __ffs(m) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(&P.ctl[0], __popc(m));
  base = __shfl_sync(0xFFFFFFFFu, base, leader);
  if (ready) {
    P.lev_nodes[base + __popc(m & ((1u << lane) - 1u))] = v;
    P.depth[v] = 1u;   // no in-arc: a child of R
    P.up[v] = P.n;
  }
}

// Kahn levels + tree parents + subtree sizes.  Level counters rotate over three words so that nobody reads a counter
// while it is reset or appended to: during level L appends go to ctl[(L + 1) % 3], ctl[(L + 2) % 3] is cleared, and
// ctl[L % 3] (the size of level L) was read by everybody before any append of level L + 1 can happen.
__global__ void __launch_bounds__(kDagThreads)
k_dag_tree(DagParams P) {
  const uint32_t G = gridDim.x, c = blockIdx.x, tid = threadIdx.x, lane = tid & 31u;
  const uint32_t sub = lane & (kSub - 1), grp = (c * kDagThreads + tid) / kSub, n_grp = G * kDagThreads / kSub;
  unsigned int bar_epoch = 0;
  uint32_t lo = 0, hi = __ldcg(&P.ctl[0]), level = 0, status = 0;
  while (lo < hi) {
    if (level + 2 >= kMaxLevels) { status = kDagTooDeep; break; }  // uniform
    if (c == 0 && tid == 0) { P.lev_off[level] = lo; P.ctl[(level + 2) % 3] = 0u; }
    uint32_t* const next_cnt = &P.ctl[(level + 1) % 3];
    const uint32_t count = hi - lo;
    // ---- phase 1: every in-arc of the states of this level was processed in an earlier level, so their tree parents
    // are final: record depth and the binary-lifting ancestor table (one thread per state; a chain of log2(depth)
    // dependent loads that would otherwise serialise inside the diverged lane that takes a state's last in-arc away)
    for (uint32_t i = c * kDagThreads + tid; i < count; i += G * kDagThreads) {
      const uint32_t v = __ldcg(&P.lev_nodes[lo + i]);
      const uint32_t parent = (uint32_t)(__ldcg(&P.best[v]) >> 32);
      const uint32_t dv = __ldcg(&P.depth[parent]) + 1u;
      P.depth[v] = dv;
      uint32_t* upv = P.up + v;
      upv[0] = parent;
      uint32_t anc = parent;
      for (uint32_t j = 1; (1u << j) <= dv; j++) { anc = up_at(P, j - 1, anc); upv[(size_t)j * (P.n + 1)] = anc; }
    }
    grid_barrier(P.ctl + 5, bar_epoch);  // a comparison below may meet any state of this level as the other candidate
    // ---- phase 2: candidacies.  Uniform trip count per warp: the lanes of a warp vote inside the loop.
    for (uint32_t g0 = grp - (lane / kSub); g0 < count; g0 += n_grp) {
      const uint32_t g = g0 + lane / kSub;
      const bool live = g < count;
      uint32_t v = 0, d = 0, a_lo = 0, a_hi = 0;
      if (live) {
        v = __ldcg(&P.lev_nodes[lo + g]);
        d = __ldcg(&P.depth[v]);
        a_lo = __ldg(&P.off[v]); a_hi = __ldg(&P.off[v + 1]);
      }
      uint32_t rounds = live ? (a_hi - a_lo + kSub - 1) / kSub : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rounds = max(rounds, __shfl_xor_sync(0xFFFFFFFFu, rounds, o));
      for (uint32_t r = 0; r < rounds; r++) {
        const uint32_t e = a_lo + r * kSub + sub;
        bool ready = false;
        uint32_t t = 0;
        if (live && e < a_hi) {
          t = __ldg(&P.arcs[e].nextstate);
          const uint32_t pos = e - a_lo;
          const unsigned long long mine = pack_cand(v, pos);
          unsigned long long cur = __ldcg(&P.best[t]);
          while (cand_less(P, v, pos, d, (uint32_t)(cur >> 32), (uint32_t)cur)) {
            const unsigned long long old = atomicCAS(&P.best[t], cur, mine);
            if (old == cur) break;
            cur = old;
          }
          // No fence between the candidacy and the count-down: the loop above only ends once a compare-and-swap has
          // RETURNED (or none was needed), i.e. after it was performed at L2, the point of coherence of both atomics, and
          // best[t] is next read after a grid barrier.
          ready = atomicSub(&P.indeg[t], 1u) == 1u;  // last in-arc: t joins the next level
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, ready);
        if (m) {
          const uint32_t leader = __ffs(m) - 1;
          uint32_t base = 0;
          if (lane == leader) base = atomicAdd(next_cnt, __popc(m));
          base = __shfl_sync(0xFFFFFFFFu, base, leader);
          if (ready) P.lev_nodes[hi + base + __popc(m & ((1u << lane) - 1u))] = t;
        }
      }
    }
    grid_barrier(P.ctl + 5, bar_epoch);
    lo = hi;
    hi += __ldcg(next_cnt);
    level++;
  }
  const uint32_t n_levels = level;
  if (c == 0 && tid == 0) {
    P.lev_off[n_levels] = lo;
    P.ctl[3] = n_levels;
    P.ctl[6] = lo;
    if (status == 0 && lo < P.n) status = kDagCyclic;  // states on or behind a cycle never become ready
    P.ctl[4] = status;
  }
  if (status || lo < P.n) return;  // uniform
  // ---- subtree sizes: reverse sweep over the levels (the children of a state sit in later levels)
  grid_barrier(P.ctl + 5, bar_epoch);  // lev_off is complete
  const uint32_t gtid = c * kDagThreads + tid, gsize = G * kDagThreads;
  for (uint32_t L = n_levels; L-- > 0;) {
    const uint32_t l0 = __ldcg(&P.lev_off[L]), l1 = __ldcg(&P.lev_off[L + 1]);
    for (uint32_t i = l0 + gtid; i < l1; i += gsize) {
      const uint32_t v = __ldcg(&P.lev_nodes[i]);
      const uint32_t parent = (uint32_t)(__ldcg(&P.best[v]) >> 32);
      if (parent != P.n) atomicAdd(&P.sizes[parent], __ldcg(&P.sizes[v]));
    }
    grid_barrier(P.ctl + 5, bar_epoch);
  }
}

// Roots (children of R): start comes first, then the other roots by id, so in reverse finish order start follows all the
// other trees and a root follows the trees of the roots with larger ids.
__global__ void k_dag_root_sizes(DagParams P, uint32_t* __restrict__ g) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v > P.n) return;
  g[v] = (v < P.n && v != P.start && (uint32_t)(P.best[v] >> 32) == P.n) ? P.sizes[v] : 0u;
}
__global__ void k_dag_root_orders(DagParams P, const uint32_t* __restrict__ g, const uint32_t* __restrict__ pref) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= P.n || (uint32_t)(P.best[v] >> 32) != P.n) return;
  const uint32_t total = pref[P.n];
  P.order[v] = v == P.start ? total : total - pref[v] - g[v];
}

// order(child) = order(parent) + 1 + sizes of the tree children to its right: forward sweep over the levels, kSub lanes
// per state walking its arcs from the last to the first.
__global__ void __launch_bounds__(kDagThreads)
k_dag_orders(DagParams P, uint32_t n_levels) {
  const uint32_t G = gridDim.x, c = blockIdx.x, tid = threadIdx.x, lane = tid & 31u;
  const uint32_t sub = lane & (kSub - 1), grp = (c * kDagThreads + tid) / kSub, n_grp = G * kDagThreads / kSub;
  const uint32_t sub_mask = ((1u << kSub) - 1u) << (lane & ~(kSub - 1));  // the lanes of my group
  unsigned int bar_epoch = 0;
  for (uint32_t L = 0; L < n_levels; L++) {
    const uint32_t l0 = __ldcg(&P.lev_off[L]), count = __ldcg(&P.lev_off[L + 1]) - l0;
    for (uint32_t g0 = grp - (lane / kSub); g0 < count; g0 += n_grp) {
      const uint32_t g = g0 + lane / kSub;
      const bool live = g < count;
      uint32_t u = 0, a_lo = 0, a_hi = 0, running = 0;
      if (live) {
        u = __ldcg(&P.lev_nodes[l0 + g]);
        a_lo = __ldg(&P.off[u]); a_hi = __ldg(&P.off[u + 1]);
        running = __ldcg(&P.order[u]) + 1u;
      }
      uint32_t rounds = live ? (a_hi - a_lo + kSub - 1) / kSub : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rounds = max(rounds, __shfl_xor_sync(0xFFFFFFFFu, rounds, o));
      for (uint32_t r = 0; r < rounds; r++) {
        const uint32_t k = r * kSub + sub;  // k-th arc from the end
        uint32_t t = 0, sz = 0;
        if (live && k < a_hi - a_lo) {
          const uint32_t e = a_hi - 1u - k;
          t = __ldg(&P.arcs[e].nextstate);
          if (__ldcg(&P.best[t]) == pack_cand(u, e - a_lo)) sz = __ldcg(&P.sizes[t]);  // tree arc
        }
        // exclusive scan of sz over the group's lanes (lane order = right to left in the arc list)
        uint32_t inc = sz;
#pragma unroll
        for (uint32_t o = 1; o < kSub; o <<= 1) {
          const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, inc, o);
          if (sub >= o) inc += x;
        }
        if (sz) P.order[t] = running + inc - sz;
        running += __shfl_sync(0xFFFFFFFFu, inc, (lane & ~(kSub - 1)) + kSub - 1);
      }
      (void)sub_mask;
    }
    grid_barrier(P.ctl + 7, bar_epoch);  // its own arrival counter: ctl[5] still holds the count of k_dag_tree
  }
}

int coop_grid(const void* kern, int threads) {
  int per_sm = 0;
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0));
  if (per_sm < 1) throw FstError("cooperative kernel does not fit on the device");
  return sm_count() * per_sm;
}

}  // namespace

// order[s] of the reference's TopOrderQueue for an acyclic machine (see the header of this file).  Returns false when
// the machine is cyclic or has more Kahn levels than the device path handles; the caller then takes the host DFS.
bool dag_top_order_device(const DevFst& f, DevBuf<uint32_t>& order, float* ms, uint64_t* launches, cudaStream_t s) {
  const uint32_t n = f.num_states, a = f.num_arcs;
  if (n == 0 || !f.has_start) return false;
  DeviceExclusive excl(device_exclusive());  // the persistent kernels want every SM (device_common.cu)
  cudaEvent_t e0, e1;
  B200_CUDA(cudaEventCreate(&e0)); B200_CUDA(cudaEventCreate(&e1));
  B200_CUDA(cudaEventRecord(e0, s));
  uint32_t log_n = 1;
  while ((1ull << log_n) <= (unsigned long long)n + 1) log_n++;
  order.reserve_discard(n);
  DevBuf<uint32_t> indeg(s, n), depth(s, (size_t)n + 1), lev_nodes(s, n), lev_off(s, kMaxLevels + 2), sizes(s, n), ctl(s, 8);
  DevBuf<uint32_t> up(s, (size_t)log_n * ((size_t)n + 1)), g(s, (size_t)n + 1), pref(s, (size_t)n + 1);
  DevBuf<unsigned long long> best(s, (size_t)n + 1);
  DevBuf<uint8_t> scan_tmp(s);
  DagParams P{};
  P.off = f.offsets.p; P.arcs = f.arcs.p; P.n = n; P.start = f.start;
  P.indeg = indeg.p; P.best = best.p; P.depth = depth.p; P.up = up.p; P.lev_nodes = lev_nodes.p; P.lev_off = lev_off.p;
  P.sizes = sizes.p; P.order = order.p; P.ctl = ctl.p;
  B200_CUDA(cudaMemsetAsync(indeg.p, 0, (size_t)n * 4, s));
  k_dag_init<<<blocks_for((size_t)n + 1), kThreads, 0, s>>>(P);
  if (a) k_dag_indeg<<<blocks_for(a), kThreads, 0, s>>>(f.arcs.p, a, indeg.p);
  k_dag_seed<<<blocks_for(n), kThreads, 0, s>>>(P);
  void* args[] = {(void*)&P};
  const int grid_a = coop_grid((void*)k_dag_tree, kDagThreads);
  B200_CUDA(cudaLaunchCooperativeKernel((void*)k_dag_tree, dim3(grid_a), dim3(kDagThreads), args, 0, s));
  uint32_t h[8];
  B200_CUDA(cudaMemcpyAsync(h, ctl.p, 32, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  if (launches) *launches += 4;
  bool ok = h[4] == 0 && h[6] == n;
  if (ok) {
    uint32_t n_levels = h[3];
    k_dag_root_sizes<<<blocks_for((size_t)n + 1), kThreads, 0, s>>>(P, g.p);
    exclusive_sum_u32(g.p, pref.p, (size_t)n + 1, scan_tmp, s);
    k_dag_root_orders<<<blocks_for(n), kThreads, 0, s>>>(P, g.p, pref.p);
    void* args_b[] = {(void*)&P, (void*)&n_levels};
    const int grid_b = coop_grid((void*)k_dag_orders, kDagThreads);
    B200_CUDA(cudaLaunchCooperativeKernel((void*)k_dag_orders, dim3(grid_b), dim3(kDagThreads), args_b, 0, s));
    if (launches) *launches += 4;
    if (std::getenv("B200_COOP_TRACE")) std::fprintf(stderr, "[dag-order] %u states, %u arcs, %u levels\n", n, a, n_levels);
  }
  B200_CUDA(cudaEventRecord(e1, s));
  B200_CUDA(cudaStreamSynchronize(s));
  if (ms) B200_CUDA(cudaEventElapsedTime(ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return ok;
}

}  // namespace b200
