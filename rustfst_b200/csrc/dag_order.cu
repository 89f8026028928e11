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
#include <vector>

#include "algos.h"
#include "coop_utils.cuh"

namespace b200 {
namespace {
using namespace coop;

constexpr uint32_t kDagThreads = 512;
constexpr uint32_t kSub = 8;             // lanes that share one state (its arcs are strided over them)
constexpr uint32_t kMaxLevels = 1u << 16;
constexpr uint32_t kReadyBuf = 96;         // per-warp buffer of newly ready states (shared memory)
enum DagStatus : uint32_t { kDagCyclic = 1, kDagTooDeep = 2, kDagNeedRows = 4 };

constexpr uint32_t kLow = 4;  // ancestors 2^0 .. 2^3 of a state live in its record; higher ones in the table up_hi
constexpr uint32_t kWinOver = 15;  // the record's window: the last 8 arc positions of the path, 4 bits each, 15 = "15 or more"

// Everything a comparison needs to know about a processed state, in ONE 32-byte sector (one 256-bit load):
// depth, position of the tree arc in the parent's list, position of the path's FIRST arc (the one that leaves R), the
// last eight positions of the path packed 4 bits each (oldest in the top nibble; kWinOver = "15 or more"), and the
// 2^j-th ancestors for j < kLow (R where the path is shorter).
struct alignas(32) NodeRec { uint32_t depth, pos, rootpos, win, up[kLow]; };
__device__ __forceinline__ bool win_overflows(uint32_t w) {  // some nibble equals 15
  return ((w & (w >> 1) & (w >> 2) & (w >> 3)) & 0x11111111u) != 0u;
}

struct DagParams {
  const uint32_t* off; const Tr* arcs; uint32_t n; uint32_t start;
  uint32_t* indeg;            // in-arcs of every state that are still to be processed
  unsigned long long* best;   // n + 1: (tree parent << 32 | position of the tree arc in the parent's list); node n = R
  NodeRec* rec;               // n + 1 records, rec[R] = {0, 0, R ...}; written when the state's level starts
  uint32_t* up_hi;            // up_hi[(j - kLow) * (n + 1) + v] = 2^j-th ancestor of v, written for kLow <= j, 2^j <= depth[v]
  uint32_t* lev_nodes;        // states in Kahn level order; level k = lev_nodes[lev_off[k] .. lev_off[k + 1])
  uint32_t* lev_off;
  uint32_t* sizes;            // subtree sizes
  uint32_t* order;            // result
  uint32_t* ctl;              // [0..2] rotating level counters, [3] #levels, [4] status, [5] barrier of k_dag_tree, [6] #states processed,
                              // [7] barrier of k_dag_orders
  uint32_t hi_rows;           // rows of up_hi that exist: ancestors up to 2^(kLow + hi_rows - 1)
  unsigned long long* trace;  // optional (B200_COOP_TRACE): ns of CTA 0 in [0] records, [1] barrier, [2] candidacies, [3] barrier, [4] sizes
};

__device__ __forceinline__ unsigned long long pack_cand(uint32_t node, uint32_t pos) {
  return ((unsigned long long)node << 32) | pos;
}
// L2-coherent 256-bit accesses (LDG/STG.E.ENL2.256): records are written by other SMs one grid barrier earlier
__device__ __forceinline__ NodeRec ld_rec(const DagParams& P, uint32_t v) {
  NodeRec r;
  asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.depth), "=r"(r.pos), "=r"(r.rootpos), "=r"(r.win), "=r"(r.up[0]), "=r"(r.up[1]), "=r"(r.up[2]), "=r"(r.up[3])
               : "l"(P.rec + v));
  return r;
}
__device__ __forceinline__ void st_rec(const DagParams& P, uint32_t v, const NodeRec& r) {
  asm volatile("st.global.cg.v8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"
               :: "r"(r.depth), "r"(r.pos), "r"(r.rootpos), "r"(r.win), "r"(r.up[0]), "r"(r.up[1]), "r"(r.up[2]), "r"(r.up[3]), "l"(P.rec + v)
               : "memory");
}
__device__ __forceinline__ uint32_t ld_next_state(const Tr* arc) {  // pinned in program order: issued a round ahead of its use
  uint32_t t;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(t) : "l"(&arc->nextstate));
  return t;
}
__device__ __forceinline__ uint32_t up_low(const NodeRec& r, uint32_t j) {  // j < kLow, not a compile-time constant
  uint32_t u = r.up[0];
#pragma unroll
  for (uint32_t k = 1; k < kLow; k++) u = j == k ? r.up[k] : u;
  return u;
}
__device__ __forceinline__ uint32_t up_hi_at(const DagParams& P, uint32_t j, uint32_t v) {
  return __ldcg(&P.up_hi[(size_t)(j - kLow) * (P.n + 1) + v]);
}
// climbs k levels from x, whose record is in rx; leaves the record of the state it arrives at in rx
__device__ __forceinline__ void lift(const DagParams& P, uint32_t& x, NodeRec& rx, uint32_t k) {
  while (k) {
    const uint32_t j = 31u - __clz(k);
    x = j < kLow ? up_low(rx, j) : up_hi_at(P, j, x);
    rx = ld_rec(P, x);
    k -= 1u << j;
  }
}
// path(a) . pa  <  path(b) . pb   (lexicographic; a and b are processed states or R; rx = record of a, consumed).
// Typical cost on a lattice (candidates of similar depth that part a few levels up): the record of b plus one pair of
// records per set bit of the distance to the parting point.
__device__ __forceinline__ bool cand_less(const DagParams& P, uint32_t a, NodeRec& rx, uint32_t pa, uint32_t b, uint32_t pb) {
  if (a == b) return pa < pb;
  if (b == P.n) return rx.rootpos < pb;  // b's candidate is an arc of R itself: the first arcs decide (they cannot be equal)
  NodeRec ry = ld_rec(P, b);
  // Same depth and the same ancestor eight levels up (or both paths shorter than that): the paths are that ancestor's
  // path followed by the two windows, so the windows decide — no further loads.  A window with a position >= 15 in it
  // takes the general route.
  if (rx.depth == ry.depth && rx.up[3] == ry.up[3] && rx.win != ry.win && !win_overflows(rx.win) && !win_overflows(ry.win))
    return rx.win < ry.win;
  uint32_t x = a, y = b, d = rx.depth;
  if (rx.depth > ry.depth) {  // is b an ancestor of a?  then a's path leaves b through the arc towards a
    lift(P, x, rx, rx.depth - ry.depth - 1);
    if (rx.up[0] == b) return rx.pos < pb;
    x = rx.up[0]; rx = ld_rec(P, x); d = ry.depth;
  } else if (ry.depth > rx.depth) {
    lift(P, y, ry, ry.depth - rx.depth - 1);
    if (ry.up[0] == a) return pa < ry.pos;
    y = ry.up[0]; ry = ld_rec(P, y);
  }
  if (x == y) return false;  // cannot happen on a DAG (one candidate's path would run through the target)
  // x != y at the same depth d >= 1: climb to the children of their lowest common ancestor
  if (d >> kLow) {
    bool moved = false;
    for (int j = 31 - __clz(d); j >= (int)kLow; j--) {
      if ((1u << j) > d) continue;
      const uint32_t ux = up_hi_at(P, j, x), uy = up_hi_at(P, j, y);
      if (ux != uy) { x = ux; y = uy; d -= 1u << j; moved = true; }
    }
    if (moved) { rx = ld_rec(P, x); ry = ld_rec(P, y); }
  }
#pragma unroll
  for (int j = (int)kLow - 1; j >= 0; j--) {  // ancestors beyond the root read R on both sides: no depth test needed
    const uint32_t ux = rx.up[j], uy = ry.up[j];
    if (ux != uy) { x = ux; y = uy; rx = ld_rec(P, x); ry = ld_rec(P, y); }
  }
  return rx.pos < ry.pos;
}

// in-degrees, initial candidates (R, position of the virtual arc: start first, then the states by id), sizes
__global__ void k_dag_init(DagParams P) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < P.n) {
    P.best[v] = pack_cand(P.n, v == P.start ? 0u : v + 1u);
    P.sizes[v] = 1u;
  } else if (v == P.n) {
    P.best[v] = pack_cand(P.n, 0u);
    NodeRec r; r.depth = 0u; r.pos = 0u; r.rootpos = 0u; r.win = 0u;
    for (uint32_t j = 0; j < kLow; j++) r.up[j] = P.n;
    st_rec(P, v, r);
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
  if (ready) P.lev_nodes[base + __popc(m & ((1u << lane) - 1u))] = v;  // no in-arc: a child of R
}

// Kahn levels + tree parents + subtree sizes.  Level counters rotate over three words so that nobody reads a counter
// while it is reset or appended to: during level L appends go to ctl[(L + 1) % 3], ctl[(L + 2) % 3] is cleared, and
// ctl[L % 3] (the size of level L) was read by everybody before any append of level L + 1 can happen.
__global__ void __launch_bounds__(kDagThreads, 2)
k_dag_tree(DagParams P) {
  const uint32_t G = gridDim.x, c = blockIdx.x, tid = threadIdx.x, lane = tid & 31u;
  __shared__ uint32_t s_ready[kDagThreads / 32u][kReadyBuf];
  unsigned int bar_epoch = 0;
  uint32_t lo = 0, hi = __ldcg(&P.ctl[0]), level = 0, status = 0;
  const bool tracing = P.trace != nullptr && tid == 0;
  unsigned long long t_prev = tracing ? globaltimer_ns() : 0ull, t_acc[5] = {0, 0, 0, 0, 0};
  auto lap = [&](int k) { if (tracing) { const unsigned long long t = globaltimer_ns(); t_acc[k] += t - t_prev; t_prev = t; } };
  while (lo < hi) {
    if (level + 2 >= kMaxLevels) { status = kDagTooDeep; break; }  // uniform
    // a state of Kahn level L sits at depth <= L + 1: the ancestor table must reach that far (the host retries with all rows)
    if (P.hi_rows < 32u - kLow && level + 1 >= (1u << (kLow + P.hi_rows))) { status = kDagNeedRows; break; }
    if (c == 0 && tid == 0) { P.lev_off[level] = lo; P.ctl[(level + 2) % 3] = 0u; }
    uint32_t* const next_cnt = &P.ctl[(level + 1) % 3];
    const uint32_t count = hi - lo;
    // ---- phase 1: every in-arc of the states of this level was processed in an earlier level, so their tree parents
    // are final: record depth and the binary-lifting ancestor table (one thread per state; a chain of log2(depth)
    // dependent loads that would otherwise serialise inside the diverged lane that takes a state's last in-arc away)
    for (uint32_t i = c * kDagThreads + tid; i < count; i += G * kDagThreads) {
      const uint32_t v = __ldcg(&P.lev_nodes[lo + i]);
      const unsigned long long bv = __ldcg(&P.best[v]);
      const uint32_t parent = (uint32_t)(bv >> 32);
      const NodeRec rp = ld_rec(P, parent);
      NodeRec r;
      r.depth = rp.depth + 1u; r.pos = (uint32_t)bv; r.up[0] = parent; r.up[1] = rp.up[0];
      r.rootpos = parent == P.n ? r.pos : rp.rootpos;
      r.win = (rp.win << 4) | min(r.pos, kWinOver);  // the window of R is empty
#pragma unroll
      for (uint32_t j = 2; j < kLow; j++) r.up[j] = __ldcg(&P.rec[r.up[j - 1]].up[j - 1]);  // rec[R].up[*] = R
      st_rec(P, v, r);
      uint32_t anc = r.up[kLow - 1];
      for (uint32_t j = kLow; (1u << j) <= r.depth; j++) {
        anc = j == kLow ? __ldcg(&P.rec[anc].up[kLow - 1]) : up_hi_at(P, j - 1, anc);
        P.up_hi[(size_t)(j - kLow) * (P.n + 1) + v] = anc;
      }
    }
    __syncthreads(); lap(0);
    grid_barrier(P.ctl + 5, bar_epoch);  // a comparison below may meet any state of this level as the other candidate
    lap(1);
    // ---- phase 2: candidacies, one lane per ARC.  A warp takes a tile of consecutive states of the level (sized so that
    // every warp of the grid gets one), scans their degrees, and its lanes walk the tile's arcs 32 at a time; the owner
    // state of an arc is found by a binary search over the scanned degrees held in the lanes.
    const uint32_t n_warps = G * (kDagThreads / 32u), gw = c * (kDagThreads / 32u) + (tid >> 5);
    uint32_t* const my_ready = s_ready[tid >> 5];
    uint32_t n_buf = 0;  // warp-uniform: states that became ready and wait in shared memory for one common append
    auto flush = [&]() {
      __syncwarp();
      uint32_t nb = 0;
      if (lane == 0) nb = atomicAdd(next_cnt, n_buf);
      nb = __shfl_sync(0xFFFFFFFFu, nb, 0);
      for (uint32_t i = lane; i < n_buf; i += 32u) P.lev_nodes[hi + nb + i] = my_ready[i];
      __syncwarp();
      n_buf = 0;
    };
    const uint32_t tile = min(32u, max(1u, (count + n_warps - 1u) / n_warps));
    for (uint32_t base = gw * tile; base < count; base += n_warps * tile) {
      const bool own = lane < tile && base + lane < count;
      uint32_t v_own = 0, lo_own = 0, deg = 0;
      if (own) {
        v_own = __ldcg(&P.lev_nodes[lo + base + lane]);
        lo_own = __ldg(&P.off[v_own]); deg = __ldg(&P.off[v_own + 1]) - lo_own;
      }
      uint32_t incl = deg;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, incl, o); if ((int)lane >= o) incl += x; }
      const uint32_t excl = incl - deg, total = __shfl_sync(0xFFFFFFFFu, incl, 31);
      // Software pipeline over the rounds of 32 arcs: each stage's loads are issued one iteration before their results
      // are needed, so that a round's critical path is the comparison and the compare-and-swap only.  Iteration i runs
      //   [B] best[t] and the source's record for round i-1 (t arrived during the previous iteration),
      //   [A] the arc of round i,
      //   [C] comparison, compare-and-swap and the in-degree count-down for round i-2,
      //   [D] the answer of the count-down of round i-3.
      // A stale best[t] costs at most a failed compare-and-swap, which returns the current value.  (The plain loop — one
      // round at a time, 40 registers, three CTAs per SM — runs the kernel in the same 3.0 ms: a level lasts as long as
      // its slowest warp, ~48 us, while the median CTA needs ~30.)
      const uint32_t rounds = (total + 31u) >> 5;
      uint32_t tA = 0, vA = 0, posA = 0, tB = 0, vB = 0, posB = 0, tD = 0, oldD = 0;
      bool actA = false, actB = false, actD = false;
      unsigned long long curB = 0;
      NodeRec rvB{};
      for (uint32_t i = 0; i < rounds + 3u; i++) {
        const uint32_t tC = tB, vC = vB, posC = posB;
        const bool actC = actB;
        unsigned long long cur = curB;
        NodeRec rvC = rvB;
        // [B]
        tB = tA; vB = vA; posB = posA; actB = actA;
        if (actB) { curB = __ldcg(&P.best[tB]); rvB = ld_rec(P, vB); }
        // [A]
        const uint32_t k = i * 32u + lane;
        actA = i < rounds && k < total;
        {
          uint32_t o = 0;  // first lane whose inclusive degree sum exceeds k
#pragma unroll
          for (uint32_t step = 16; step > 0; step >>= 1) {
            const uint32_t probe = __shfl_sync(0xFFFFFFFFu, incl, (o + step - 1u) & 31u);
            if (probe <= k) o += step;
          }
          o = actA ? o : 0u;
          vA = __shfl_sync(0xFFFFFFFFu, v_own, o);
          posA = k - __shfl_sync(0xFFFFFFFFu, excl, o);
          const uint32_t e = __shfl_sync(0xFFFFFFFFu, lo_own, o) + posA;
          if (actA) tA = ld_next_state(P.arcs + e);
        }
        // [C]
        uint32_t old_c = 0;
        if (actC) {
          const unsigned long long mine = pack_cand(vC, posC);
          while (cand_less(P, vC, rvC, posC, (uint32_t)(cur >> 32), (uint32_t)cur)) {
            const unsigned long long old = atomicCAS(&P.best[tC], cur, mine);
            if (old == cur) break;
            cur = old;
            rvC = ld_rec(P, vC);  // the comparison consumed its copy
          }
          // No fence between the candidacy and the count-down: the loop above only ends once a compare-and-swap has
          // RETURNED (or none was needed), i.e. after it was performed at L2, the point of coherence of both atomics, and
          // best[t] is next read after a grid barrier.
          old_c = atomicSub(&P.indeg[tC], 1u);
        }
        // [D] whoever takes the last in-arc away appends the state to the next level
        const bool ready = actD && oldD == 1u;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, ready);
        if (m) {  // one append to the next level per kReadyBuf states, not per round: the counter is ONE address
          if (n_buf + __popc(m) > kReadyBuf) flush();
          if (ready) my_ready[n_buf + __popc(m & ((1u << lane) - 1u))] = tD;
          n_buf += __popc(m);
        }
        tD = tC; oldD = old_c; actD = actC;
      }
    }
    if (n_buf) flush();
    __syncthreads(); lap(2);
    grid_barrier(P.ctl + 5, bar_epoch);
    lap(3);
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
  t_prev = tracing ? globaltimer_ns() : 0ull;
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
  lap(4);
  if (tracing) for (int k = 0; k < 5; k++) P.trace[(size_t)c * 8 + k] = t_acc[k];
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
  if (const char* e = std::getenv("B200_DAG_CTAS_PER_SM")) per_sm = std::max(1, std::min(per_sm, std::atoi(e)));
  return sm_count() * per_sm;
}

}  // namespace

// One attempt with the short (all_rows = false) or the full ancestor table: 1 = done, 0 = cyclic / too many levels, -1 = the
// machine is deeper than the short table reaches.
static int dag_top_order_try(const DevFst& f, DevBuf<uint32_t>& order, float* ms, uint64_t* launches, cudaStream_t s, bool all_rows) {
  const uint32_t n = f.num_states, a = f.num_arcs;
  if (n == 0 || !f.has_start) return 0;
  DeviceExclusive excl(device_exclusive());  // the persistent kernels want every SM (device_common.cu)
  cudaEvent_t e0, e1;
  B200_CUDA(cudaEventCreate(&e0)); B200_CUDA(cudaEventCreate(&e1));
  B200_CUDA(cudaEventRecord(e0, s));
  uint32_t log_n = 1;
  while ((1ull << log_n) <= (unsigned long long)n + 1) log_n++;
  order.reserve_discard(n);
  DevBuf<uint32_t> indeg(s, n), lev_nodes(s, n), lev_off(s, kMaxLevels + 2), sizes(s, n), ctl(s, 8);
  // Ancestors 2^kLow and up live in a side table of (n + 1)-word rows.  A lattice is a few dozen to a few hundred levels
  // deep: six rows (depth < 1024) are allocated first, all log2(n) rows only when the machine turns out to be deeper.
  const uint32_t full_rows = log_n > kLow ? log_n - kLow : 1;
  const uint32_t hi_rows = all_rows ? full_rows : std::min<uint32_t>(full_rows, 6u);
  DevBuf<uint32_t> up_hi(s, (size_t)hi_rows * ((size_t)n + 1)), g(s, (size_t)n + 1), pref(s, (size_t)n + 1);
  DevBuf<unsigned long long> best(s, (size_t)n + 1);
  DevBuf<NodeRec> rec(s, (size_t)n + 1);
  DevBuf<uint8_t> scan_tmp(s);
  DagParams P{};
  P.off = f.offsets.p; P.arcs = f.arcs.p; P.n = n; P.start = f.start;
  P.indeg = indeg.p; P.best = best.p; P.rec = rec.p; P.up_hi = up_hi.p; P.hi_rows = hi_rows; P.lev_nodes = lev_nodes.p; P.lev_off = lev_off.p;
  P.sizes = sizes.p; P.order = order.p; P.ctl = ctl.p;
  const bool tracing = std::getenv("B200_COOP_TRACE") != nullptr;
  DevBuf<unsigned long long> trace(s, 8 * 4096);
  if (tracing) { B200_CUDA(cudaMemsetAsync(trace.p, 0, 8 * 8 * 4096, s)); P.trace = trace.p; }
  B200_CUDA(cudaMemsetAsync(indeg.p, 0, (size_t)n * 4, s));
  k_dag_init<<<blocks_for((size_t)n + 1), kThreads, 0, s>>>(P);
  if (a) k_dag_indeg<<<blocks_for(a), kThreads, 0, s>>>(f.arcs.p, a, indeg.p);
  k_dag_seed<<<blocks_for(n), kThreads, 0, s>>>(P);
  cudaEvent_t e_a = nullptr, e_b = nullptr;
  if (tracing) { B200_CUDA(cudaEventCreate(&e_a)); B200_CUDA(cudaEventCreate(&e_b)); B200_CUDA(cudaEventRecord(e_a, s)); }
  void* args[] = {(void*)&P};
  const int grid_a = coop_grid((void*)k_dag_tree, kDagThreads);
  B200_CUDA(cudaLaunchCooperativeKernel((void*)k_dag_tree, dim3(grid_a), dim3(kDagThreads), args, 0, s));
  if (tracing) B200_CUDA(cudaEventRecord(e_b, s));
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
    if (tracing) {
      std::vector<unsigned long long> all((size_t)8 * grid_a);
      B200_CUDA(cudaMemcpyAsync(all.data(), trace.p, all.size() * 8, cudaMemcpyDeviceToHost, s));
      B200_CUDA(cudaStreamSynchronize(s));
      const unsigned long long* t = all.data();
      std::vector<double> cand;
      for (int k = 0; k < grid_a; k++) cand.push_back(all[(size_t)k * 8 + 2] * 1e-6);
      std::sort(cand.begin(), cand.end());
      std::fprintf(stderr, "[dag-order] candidacies per CTA ms: min %.3f median %.3f p90 %.3f max %.3f\n", cand.front(), cand[cand.size() / 2],
                   cand[cand.size() * 9 / 10], cand.back());
      std::fprintf(stderr, "[dag-order] %u states, %u arcs, %u levels; CTA 0 ms: records %.3f barrier %.3f candidacies %.3f barrier %.3f sizes %.3f\n",
                   n, a, n_levels, t[0] * 1e-6, t[1] * 1e-6, t[2] * 1e-6, t[3] * 1e-6, t[4] * 1e-6);
    }
  }
  B200_CUDA(cudaEventRecord(e1, s));
  B200_CUDA(cudaStreamSynchronize(s));
  if (ms) B200_CUDA(cudaEventElapsedTime(ms, e0, e1));
  if (tracing) {
    float t_pre = 0, t_tree = 0, t_post = 0;
    cudaEventElapsedTime(&t_pre, e0, e_a); cudaEventElapsedTime(&t_tree, e_a, e_b); cudaEventElapsedTime(&t_post, e_b, e1);
    std::fprintf(stderr, "[dag-order] ms: in-degrees + seed %.3f, tree kernel %.3f, host read-back + orders %.3f\n", t_pre, t_tree, t_post);
    cudaEventDestroy(e_a); cudaEventDestroy(e_b);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return ok ? 1 : (h[4] == kDagNeedRows ? -1 : 0);
}

// order[s] of the reference's TopOrderQueue for an acyclic machine (see the header of this file).  Returns false when
// the machine is cyclic or has more Kahn levels than the device path handles; the caller then takes the host DFS.
bool dag_top_order_device(const DevFst& f, DevBuf<uint32_t>& order, float* ms, uint64_t* launches, cudaStream_t s) {
  float ms1 = 0, ms2 = 0;
  int r = dag_top_order_try(f, order, &ms1, launches, s, false);
  if (r < 0) r = dag_top_order_try(f, order, &ms2, launches, s, true);  // deeper than the short ancestor table reaches
  if (ms) *ms = ms1 + ms2;
  return r > 0;
}

}  // namespace b200
