// compose_coop.cu — the whole composition BFS as ONE persistent cooperative kernel (B200: 148 SMs x resident CTAs).
//
// Same algorithm and the same canonical numbering as compose.cu (see its header for the reference citations); what
// changes is the execution model: instead of ~9 launches and 3 host read-backs per BFS wave, a grid that exactly
// fills the machine stays resident and separates the phases of a wave with grid-wide barriers.  A wave
// (frontier = product ids [lo, hi)) is four phases:
//
//   A  match   a group of G lanes per frontier state: lane j takes item j (item 0 = implicit epsilon loop, item j =
//              j-th arc of the iterated side), binary-searches the sorted side, applies the filter and records
//              (pos, count, filter states); per-state / per-group / per-CTA emission counts
//   B  emit    CTA prefix over the per-CTA counts gives every state its canonical first emission index; the same
//              group re-walks its states and writes the 16-byte output arcs (128-bit stores), looks the destination
//              tuple up in the open-addressed table (CAS insert) and atomicMin's its first emission index
//   C  rank    every arc of the wave, in emission order: "am I the first emission of a new tuple?"; CTA-local
//              exclusive ranks (ballot + popc), per-CTA counts
//   D  resolve CTA prefix over those counts turns (CTA, local rank) into the state id of the next wave; pending
//              arcs get their nextstate, first emitters publish id + tuple
//
// Frontier bounds, arc totals and overflow decisions are recomputed identically by every CTA from the per-CTA
// partial arrays, so control flow is uniform without any broadcast.  All buffers are pre-sized; if one would
// overflow the kernel stops and compose_device() falls back to the growing multi-kernel back end.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <string>

#include "compose_common.cuh"

namespace cg = cooperative_groups;

namespace b200 {
namespace {
using namespace composeimpl;

constexpr int kCoopThreads = 256;
constexpr uint32_t kTempFlag = 0x80000000u;   // slot.id holds (cta << 20 | local rank) between phases C and D
constexpr uint32_t kLocalRankBits = 20;

enum Overflow : uint32_t { kOvArcs = 1, kOvStates = 2, kOvTable = 4, kOvScratch = 8, kOvChunk = 16 };

struct CoopParams {
  FstView a, b;
  int kind, side;
  unsigned long long* tuples; uint32_t states_cap;
  uint32_t* out_offsets; float* out_finals;
  Tr* out_arcs; uint32_t arcs_cap;
  Slot* slots; uint32_t mask; uint32_t table_cap;
  uint2* scratch; uint32_t scratch_cap;   // per-item records of the current wave
  uint32_t* st_cnt; uint32_t st_cnt_cap;  // arcs emitted per frontier state of the current wave
  uint32_t* part_arcs;    // gridDim entries: arcs emitted by each CTA's state chunk
  uint32_t* part_items;   // gridDim entries: items of each CTA's state chunk (statistics)
  uint32_t* part_new;     // gridDim entries: first emissions found in each CTA's arc chunk
  uint32_t* ctl;          // [0] scratch cursor, [1] overflow flags, [2..]: results
  unsigned long long* stats;  // states_expanded, arcs_iterated, arcs_emitted, waves, ns phase A, B, C, D
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Sum of v[0..n) and of v[0..upto) computed by the whole CTA (n <= a few thousand).
__device__ __forceinline__ void cta_sum_prefix(const uint32_t* __restrict__ v, uint32_t n, uint32_t upto,
                                               uint32_t* smem2, uint32_t& total, uint32_t& prefix) {
  uint32_t t = 0, p = 0;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    uint32_t x = __ldcg(&v[i]);
    t += x;
    if (i < upto) p += x;
  }
  for (int o = 16; o > 0; o >>= 1) { t += __shfl_down_sync(0xFFFFFFFFu, t, o); p += __shfl_down_sync(0xFFFFFFFFu, p, o); }
  __syncthreads();
  if (threadIdx.x == 0) { smem2[0] = 0; smem2[1] = 0; }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { atomicAdd(&smem2[0], t); atomicAdd(&smem2[1], p); }
  __syncthreads();
  total = smem2[0]; prefix = smem2[1];
  __syncthreads();
}

template <int G>
__global__ void __launch_bounds__(kCoopThreads)
k_compose_coop(CoopParams P) {
  cg::grid_group grid = cg::this_grid();
  constexpr int kGroupsPerCta = kCoopThreads / G;
  __shared__ uint32_t s_group_arcs[kGroupsPerCta];
  __shared__ uint32_t s_group_off[kGroupsPerCta];
  __shared__ uint32_t s_tmp[2];
  __shared__ uint32_t s_warp[kCoopThreads / 32];
  __shared__ uint32_t s_tile_base;
  extern __shared__ uint32_t s_prefix[];  // gridDim entries (phase D)

  const uint32_t lane = threadIdx.x % G;
  const uint32_t group_in_cta = threadIdx.x / G;
  const uint32_t n_groups = gridDim.x * kGroupsPerCta;
  const uint32_t gid = blockIdx.x * kGroupsPerCta + group_in_cta;
  // lanes of my group inside the warp
  const uint32_t group_mask = (G == 32) ? 0xFFFFFFFFu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));

  uint32_t lo = 0, hi = 1, base = 0;  // uniform across the grid by construction
  unsigned long long n_states_exp = 0, n_items = 0, n_arcs = 0, n_waves = 0;
  unsigned long long t_a = 0, t_b = 0, t_c = 0, t_d = 0;
  uint32_t overflow = 0;

  while (lo < hi) {
    const uint32_t F = hi - lo;
    unsigned long long tp0 = globaltimer_ns();
    const uint32_t chunk = (F + n_groups - 1) / n_groups;
    const uint32_t c_begin = min(F, gid * chunk), c_end = min(F, c_begin + chunk);

    // ------------------------------------------------------------------ phase A: match
    // scratch region for my chunk: one atomic per group
    uint32_t my_items = 0;
    for (uint32_t i = c_begin + lane; i < c_end; i += G) {
      uint32_t fs, s1, s2;
      unpack_key(P.tuples[lo + i], fs, s1, s2);
      uint32_t d1 = P.a.off[s1 + 1] - P.a.off[s1], d2 = P.b.off[s2 + 1] - P.b.off[s2];
      bool mi = P.side == kMatchInput || (P.side == kMatchBoth && d1 <= d2);
      my_items += 1u + (mi ? d1 : d2);
    }
    for (int o = G / 2; o > 0; o >>= 1) my_items += __shfl_xor_sync(group_mask, my_items, o, G);
    uint32_t region = 0;
    if (lane == 0 && my_items) region = atomicAdd(&P.ctl[0], my_items);
    region = __shfl_sync(group_mask, region, 0, G);
    const bool scratch_ok = (unsigned long long)region + my_items <= P.scratch_cap;
    if (!scratch_ok && lane == 0) atomicOr(&P.ctl[1], (uint32_t)kOvScratch);

    uint32_t group_arcs = 0, cursor = region;
    for (uint32_t i = c_begin; i < c_end; i++) {
      uint32_t fs, s1, s2;
      unpack_key(P.tuples[lo + i], fs, s1, s2);
      const uint32_t alo = P.a.off[s1], ahi = P.a.off[s1 + 1], blo = P.b.off[s2], bhi = P.b.off[s2 + 1];
      const bool match_input = P.side == kMatchInput || (P.side == kMatchBoth && (ahi - alo) <= (bhi - blo));
      const uint32_t nitems = 1u + (match_input ? (ahi - alo) : (bhi - blo));
      const FsFlags ff = state_flags(P.a, P.b, s1, s2);
      uint32_t state_arcs = 0;
      for (uint32_t j0 = 0; j0 < nitems; j0 += G) {
        const uint32_t j = j0 + lane;
        uint32_t cnt_out = 0;
        if (j < nitems) {
          Label label;
          if (j == 0) label = kNoLabel;
          else label = match_input ? __ldg(&P.a.arcs[alo + j - 1].olabel) : __ldg(&P.b.arcs[blo + j - 1].ilabel);
          const bool has_loop = (label == kEps);
          const Label key = (label == kNoLabel) ? kEps : label;
          uint32_t pos, end, fs_loop, fs_real;
          if (match_input) {
            pos = has_loop ? blo : lower_bound_label<false>(P.b.arcs, blo, bhi, key);
            end = run_end<false>(P.b.arcs, pos, bhi, key);
            fs_loop = has_loop ? filter_eval(P.kind, fs, ff, label, kNoLabel) : kNoFs;
            fs_real = filter_eval(P.kind, fs, ff, label, key);
          } else {
            pos = has_loop ? alo : lower_bound_label<true>(P.a.arcs, alo, ahi, key);
            end = run_end<true>(P.a.arcs, pos, ahi, key);
            fs_loop = has_loop ? filter_eval(P.kind, fs, ff, kNoLabel, label) : kNoFs;
            fs_real = filter_eval(P.kind, fs, ff, key, label);
          }
          const uint32_t cnt = end - pos;
          const bool loop_ok = has_loop && fs_loop != kNoFs;
          const bool real_ok = fs_real != kNoFs && cnt > 0;
          cnt_out = (loop_ok ? 1u : 0u) + (real_ok ? cnt : 0u);
          // record: x = pos, y = count of REAL matches that are emitted (27 bits) | loop_ok | fs_loop | fs_real
          if (scratch_ok)
            P.scratch[cursor + j] = make_uint2(pos, (real_ok ? cnt : 0u) | (loop_ok ? 1u << 27 : 0u) |
                                                        ((fs_loop & 3u) << 28) | ((fs_real & 3u) << 30));
        }
        uint32_t s = cnt_out;
        for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(group_mask, s, o, G);
        state_arcs += s;
      }
      cursor += nitems;
      if (lane == 0) {
        P.st_cnt[i] = state_arcs;
        float fw = w_times(P.a.fin[s1], P.b.fin[s2]);  // compose_fst_op.rs:420-449
        P.out_finals[lo + i] = w_is_zero(fw) ? w_zero() : fw;
      }
      group_arcs += state_arcs;
    }
    if (lane == 0) s_group_arcs[group_in_cta] = group_arcs;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t ta = 0;
      for (int g = 0; g < kGroupsPerCta; g++) { s_group_off[g] = ta; ta += s_group_arcs[g]; }
      P.part_arcs[blockIdx.x] = ta;
    }
    {
      // items of the CTA (statistics only)
      uint32_t it = (lane == 0) ? my_items : 0;
      for (int o = 16; o > 0; o >>= 1) it += __shfl_down_sync(0xFFFFFFFFu, it, o);
      if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x / 32] = it;
      __syncthreads();
      if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kCoopThreads / 32; w++) t += s_warp[w];
        P.part_items[blockIdx.x] = t;
      }
    }
    grid.sync();
    unsigned long long tp1 = globaltimer_ns();

    // ------------------------------------------------------------------ phase B: emit
    uint32_t E, cta_off, T_items, dummy;
    cta_sum_prefix(P.part_arcs, gridDim.x, blockIdx.x, s_tmp, E, cta_off);
    cta_sum_prefix(P.part_items, gridDim.x, 0, s_tmp, T_items, dummy);
    overflow = __ldcg(&P.ctl[1]);
    if ((unsigned long long)base + E > P.arcs_cap) overflow |= kOvArcs;
    if (((unsigned long long)hi + E) * 2ull > P.table_cap) overflow |= kOvTable;
    const uint32_t arc_chunk = ((E + gridDim.x - 1) / gridDim.x + 31u) & ~31u;
    if (arc_chunk > (1u << kLocalRankBits)) overflow |= kOvChunk;
    if (overflow) break;  // uniform: every CTA derives the same flags from the same global data

    {
      uint32_t state_base = cta_off + s_group_off[group_in_cta];  // wave-local emission index of my next state
      uint32_t cur = region;
      Tr* __restrict__ wave_arcs = P.out_arcs + base;
      for (uint32_t i = c_begin; i < c_end; i++) {
        uint32_t fs, s1, s2;
        unpack_key(P.tuples[lo + i], fs, s1, s2);
        const uint32_t alo = P.a.off[s1], ahi = P.a.off[s1 + 1], blo = P.b.off[s2], bhi = P.b.off[s2 + 1];
        const bool match_input = P.side == kMatchInput || (P.side == kMatchBoth && (ahi - alo) <= (bhi - blo));
        const uint32_t nitems = 1u + (match_input ? (ahi - alo) : (bhi - blo));
        if (lane == 0) P.out_offsets[lo + i] = base + state_base;
        uint32_t run = state_base;
        for (uint32_t j0 = 0; j0 < nitems; j0 += G) {
          const uint32_t j = j0 + lane;
          uint2 rec = make_uint2(0, 0);
          if (j < nitems) rec = P.scratch[cur + j];
          const bool loop_ok = (rec.y >> 27) & 1u;
          const uint32_t n_real = rec.y & 0x07FFFFFFu;
          const uint32_t mine = n_real + (loop_ok ? 1u : 0u);
          // exclusive scan of `mine` across the group's lanes
          uint32_t incl = mine;
          for (int o = 1; o < G; o <<= 1) {
            uint32_t v = __shfl_up_sync(group_mask, incl, o, G);
            if ((int)lane >= o) incl += v;
          }
          const uint32_t total = __shfl_sync(group_mask, incl, G - 1, G);
          uint32_t e = run + incl - mine;
          if (mine) {
            Tr it;
            if (j == 0) it = match_input ? Tr{kEps, kNoLabel, 0.0f, s1} : Tr{kNoLabel, kEps, 0.0f, s2};
            else it = match_input ? load_tr(&P.a.arcs[alo + j - 1]) : load_tr(&P.b.arcs[blo + j - 1]);
            for (uint32_t k = 0; k < mine; k++, e++) {
              Tr cand;
              uint32_t fsn;
              if (loop_ok && k == 0) {
                cand = match_input ? Tr{kNoLabel, kEps, 0.0f, s2} : Tr{kEps, kNoLabel, 0.0f, s1};
                fsn = (rec.y >> 28) & 3u;
              } else {
                const uint32_t idx = rec.x + k - (loop_ok ? 1u : 0u);
                cand = match_input ? load_tr(&P.b.arcs[idx]) : load_tr(&P.a.arcs[idx]);
                fsn = (rec.y >> 30) & 3u;
              }
              const Tr& arc1 = match_input ? it : cand;
              const Tr& arc2 = match_input ? cand : it;
              Tr out;
              out.ilabel = arc1.ilabel;
              out.olabel = arc2.olabel;
              out.weight = w_times(arc1.weight, arc2.weight);
              const unsigned long long key = pack_key(fsn, arc1.nextstate, arc2.nextstate);
              uint32_t h = hash_key(key) & P.mask;
              while (true) {
                unsigned long long curk = *reinterpret_cast<volatile unsigned long long*>(&P.slots[h].key);
                if (curk == key) break;
                if (curk == kEmptyKey) {
                  unsigned long long prev = atomicCAS(&P.slots[h].key, kEmptyKey, key);
                  if (prev == kEmptyKey || prev == key) break;
                }
                h = (h + 1) & P.mask;
              }
              const uint32_t id = *reinterpret_cast<volatile uint32_t*>(&P.slots[h].id);
              if (id != kUnassigned) out.nextstate = id;
              else { atomicMin(&P.slots[h].emin, e); out.nextstate = kPendingBit | h; }
              store_tr(&wave_arcs[e], out);
            }
          }
          run += total;
        }
        cur += nitems;
        state_base = run;
      }
    }
    grid.sync();
    unsigned long long tp2 = globaltimer_ns();

    // ------------------------------------------------------------------ phase C: rank first emissions
    const uint32_t e_begin = min(E, blockIdx.x * arc_chunk), e_end = min(E, e_begin + arc_chunk);
    {
      const Tr* __restrict__ wave_arcs = P.out_arcs + base;
      uint32_t cta_new = 0;  // running count (uniform inside the CTA)
      for (uint32_t t0 = e_begin; t0 < e_end; t0 += kCoopThreads) {
        const uint32_t e = t0 + threadIdx.x;
        bool owner = false;
        uint32_t h = 0;
        if (e < e_end) {
          uint32_t ns = __ldcg(&wave_arcs[e].nextstate);
          if (ns & kPendingBit) { h = ns & ~kPendingBit; owner = (__ldcg(&P.slots[h].emin) == e); }
        }
        const uint32_t word = __ballot_sync(0xFFFFFFFFu, owner);
        const uint32_t wl = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (wl == 0) s_warp[wid] = __popc(word);
        __syncthreads();
        uint32_t before = 0, tile_total = 0;
        for (int w = 0; w < kCoopThreads / 32; w++) { uint32_t c = s_warp[w]; if (w < (int)wid) before += c; tile_total += c; }
        if (owner) {
          uint32_t local = cta_new + before + __popc(word & ((1u << wl) - 1u));
          P.slots[h].id = kTempFlag | (blockIdx.x << kLocalRankBits) | local;
        }
        cta_new += tile_total;
        __syncthreads();
      }
      if (threadIdx.x == 0) P.part_new[blockIdx.x] = cta_new;
    }
    grid.sync();
    unsigned long long tp3 = globaltimer_ns();

    // ------------------------------------------------------------------ phase D: resolve
    // exclusive prefix of part_new over all CTAs, kept in shared memory
    {
      for (uint32_t i = threadIdx.x; i < gridDim.x; i += blockDim.x) s_prefix[i] = __ldcg(&P.part_new[i]);
      __syncthreads();
      if (threadIdx.x < 32) {
        uint32_t carry = 0;
        for (uint32_t i0 = 0; i0 < gridDim.x; i0 += 32) {
          uint32_t i = i0 + threadIdx.x;
          uint32_t v = i < gridDim.x ? s_prefix[i] : 0u, incl = v;
          for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o); if ((int)threadIdx.x >= o) incl += u; }
          if (i < gridDim.x) s_prefix[i] = carry + incl - v;
          carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        if (threadIdx.x == 0) s_tile_base = carry;
      }
      __syncthreads();
    }
    const uint32_t n_new = s_tile_base;
    if ((unsigned long long)hi + n_new > P.states_cap || (unsigned long long)hi + n_new >= 0x7FFFFFFFull) {
      overflow |= kOvStates;
      break;  // uniform
    }
    {
      Tr* __restrict__ wave_arcs = P.out_arcs + base;
      for (uint32_t e = e_begin + threadIdx.x; e < e_end; e += kCoopThreads) {
        uint32_t ns = wave_arcs[e].nextstate;
        if (!(ns & kPendingBit)) continue;
        const uint32_t h = ns & ~kPendingBit;
        const uint32_t v = *reinterpret_cast<volatile uint32_t*>(&P.slots[h].id);
        uint32_t id = v;
        if (v & kTempFlag) id = hi + s_prefix[(v & ~kTempFlag) >> kLocalRankBits] + (v & ((1u << kLocalRankBits) - 1u));
        wave_arcs[e].nextstate = id;
        if (__ldcg(&P.slots[h].emin) == e) {  // first emitter: publish
          P.slots[h].id = id;
          P.tuples[id] = P.slots[h].key;
        }
      }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) P.ctl[0] = 0;  // scratch cursor for the next wave
    n_states_exp += F; n_items += T_items; n_arcs += E; n_waves++;
    base += E;
    lo = hi;
    hi += n_new;
    grid.sync();
    unsigned long long tp4 = globaltimer_ns();
    t_a += tp1 - tp0; t_b += tp2 - tp1; t_c += tp3 - tp2; t_d += tp4 - tp3;
  }

  if (blockIdx.x == 0 && threadIdx.x == 0) {
    P.ctl[1] = overflow;
    P.ctl[2] = hi;    // number of product states
    P.ctl[3] = base;  // number of arcs
    P.out_offsets[hi] = base;
    P.stats[0] = n_states_exp; P.stats[1] = n_items - n_states_exp; P.stats[2] = n_arcs; P.stats[3] = n_waves;
    P.stats[4] = t_a; P.stats[5] = t_b; P.stats[6] = t_c; P.stats[7] = t_d;
  }
}

__global__ void k_coop_init(Slot* slots, uint32_t mask, unsigned long long* tuples, unsigned long long key0,
                            uint32_t* ctl) {
  uint32_t h = hash_key(key0) & mask;
  slots[h].key = key0; slots[h].id = 0; slots[h].emin = 0;
  tuples[0] = key0;
  ctl[0] = ctl[1] = ctl[2] = ctl[3] = 0;
}

template <int G>
bool run_coop(const CoopParams& P0, int sms, cudaStream_t s, float* ms_kernel) {
  CoopParams P = P0;
  int per_sm = 0;
  // dynamic smem = gridDim entries; grid <= 148 * 8 -> < 8 KB; query occupancy with a safe upper bound
  size_t dyn = (size_t)sms * 16 * sizeof(uint32_t);
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_compose_coop<G>, kCoopThreads, dyn));
  if (per_sm < 1) throw FstError("cooperative compose kernel does not fit on the device");
  if (per_sm > 8) per_sm = 8;
  int grid = sms * per_sm;
  if (grid >= 2048) grid = 2047;  // CTA index must fit 11 bits next to the 20-bit local rank
  dyn = (size_t)grid * sizeof(uint32_t);
  void* args[] = {(void*)&P};
  cudaEvent_t e0, e1;
  B200_CUDA(cudaEventCreate(&e0)); B200_CUDA(cudaEventCreate(&e1));
  B200_CUDA(cudaEventRecord(e0, s));
  B200_CUDA(cudaLaunchCooperativeKernel((void*)k_compose_coop<G>, dim3(grid), dim3(kCoopThreads), args, dyn, s));
  B200_CUDA(cudaEventRecord(e1, s));
  B200_CUDA(cudaStreamSynchronize(s));
  B200_CUDA(cudaEventElapsedTime(ms_kernel, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return true;
}

}  // namespace

bool compose_device_coop(const DevFst& fa, const DevFst& fb, const ComposeOptions& opt, ComposeStats* stats,
                         cudaStream_t s, DevFst* result) {
  int kind = opt.filter == kAutoFilter ? kSequenceFilter : opt.filter;
  if (kind < kNullFilter || kind > kNoMatchFilter) throw FstError("EnumConversionError");
  int side = resolve_match_side(fa.props, fb.props);
  if (fa.num_states >= 0x7FFFFFFFu || fb.num_states >= 0x7FFFFFFFu)
    throw FstError("compose: operands with >= 2^31 states are not supported");
  if (!fa.has_start || !fb.has_start) return false;  // trivial case handled by the multi-kernel back end

  ComposeStats local;
  ComposeStats& st = stats ? *stats : local;
  st = ComposeStats();
  cudaEvent_t ev0, ev1, ev2;
  B200_CUDA(cudaEventCreate(&ev0)); B200_CUDA(cudaEventCreate(&ev1)); B200_CUDA(cudaEventCreate(&ev2));
  B200_CUDA(cudaEventRecord(ev0, s));

  DevBuf<uint32_t> neps1(s), neps2(s);
  bool need_eps = (kind == kSequenceFilter || kind == kAltSequenceFilter || kind == kMatchFilter);
  CoopParams P{};
  P.a = FstView{fa.offsets.p, fa.arcs.p, fa.finals.p, nullptr, fa.num_states};
  P.b = FstView{fb.offsets.p, fb.arcs.p, fb.finals.p, nullptr, fb.num_states};
  if (need_eps && !(fa.props & props::kNoOEpsilons)) {
    neps1.reserve_discard(fa.num_states);
    launch_count_eps(P.a.off, P.a.arcs, P.a.n, 1, neps1.p, s);
    P.a.neps = neps1.p; st.kernel_launches++;
  }
  if (need_eps && !(fb.props & props::kNoIEpsilons)) {
    neps2.reserve_discard(fb.num_states);
    launch_count_eps(P.b.off, P.b.arcs, P.b.n, 0, neps2.p, s);
    P.b.neps = neps2.p; st.kernel_launches++;
  }
  P.kind = kind; P.side = side;

  // ---- pre-sized buffers (HBM is plentiful: 180 GB); an overflow falls back to the growing back end
  const size_t sum_states = (size_t)fa.num_states + fb.num_states, sum_arcs = (size_t)fa.num_arcs + fb.num_arcs;
  size_t states_cap = std::max<size_t>(1 << 16, 8 * sum_states);
  size_t arcs_cap = std::max<size_t>(1 << 18, 4 * sum_arcs);
  states_cap = std::min<size_t>(states_cap, 0x7FFFFFF0ull);
  arcs_cap = std::min<size_t>(arcs_cap, 0xFFFFFFF0ull);
  size_t table_cap = 1 << 17;
  while (table_cap < 2 * states_cap && table_cap < (1ull << 30)) table_cap <<= 1;
  size_t scratch_cap = std::max<size_t>(1 << 18, arcs_cap / 2);

  DevFst out(s);
  out.offsets.reserve_discard(states_cap + 1);
  out.finals.reserve_discard(states_cap);
  out.arcs.reserve_discard(arcs_cap);
  DevBuf<unsigned long long> tuples(s, states_cap), dstats(s, 8);
  DevBuf<Slot> slots(s, table_cap);
  DevBuf<uint2> scratch(s, scratch_cap);
  DevBuf<uint32_t> st_cnt(s, states_cap), parts(s, 3 * 2048), ctl(s, 8);
  B200_CUDA(cudaMemsetAsync(slots.p, 0xFF, table_cap * sizeof(Slot), s));
  B200_CUDA(cudaMemsetAsync(dstats.p, 0, 8 * sizeof(unsigned long long), s));
  P.tuples = tuples.p; P.states_cap = (uint32_t)states_cap;
  P.out_offsets = out.offsets.p; P.out_finals = out.finals.p; P.out_arcs = out.arcs.p; P.arcs_cap = (uint32_t)arcs_cap;
  P.slots = slots.p; P.mask = (uint32_t)table_cap - 1; P.table_cap = (uint32_t)table_cap;
  P.scratch = scratch.p; P.scratch_cap = (uint32_t)scratch_cap;
  P.st_cnt = st_cnt.p; P.st_cnt_cap = (uint32_t)states_cap;
  P.part_arcs = parts.p; P.part_items = parts.p + 2048; P.part_new = parts.p + 4096;
  P.ctl = ctl.p; P.stats = dstats.p;
  uint32_t start_fs = (kind == kNullFilter || kind == kTrivialFilter || kind == kNoMatchFilter) ? 1u : 0u;
  k_coop_init<<<1, 1, 0, s>>>(slots.p, P.mask, tuples.p, pack_key(start_fs, fa.start, fb.start), ctl.p);
  st.kernel_launches++;

  // lanes per frontier state: enough to cover the average (1 + degree) of the smaller-degree side
  double d1 = fa.num_states ? (double)fa.num_arcs / fa.num_states : 0, d2 = fb.num_states ? (double)fb.num_arcs / fb.num_states : 0;
  double items = 1.0 + (side == kMatchInput ? d1 : side == kMatchOutput ? d2 : std::min(d1, d2));
  float ms_kernel = 0;
  int sms = sm_count();
  if (items > 16.0) run_coop<32>(P, sms, s, &ms_kernel);
  else if (items > 8.0) run_coop<16>(P, sms, s, &ms_kernel);
  else run_coop<8>(P, sms, s, &ms_kernel);
  st.kernel_launches++; st.emit_launches = 1;

  uint32_t hctl[4];
  unsigned long long hstats[8];
  B200_CUDA(cudaMemcpyAsync(hctl, ctl.p, 16, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaMemcpyAsync(hstats, dstats.p, 64, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  if (hctl[1] != 0) {  // a pre-sized buffer was too small
    cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev2);
    return false;
  }
  st.states_expanded = hstats[0]; st.arcs_iterated = hstats[1]; st.arcs_emitted = hstats[2]; st.waves = hstats[3];
  st.ms_emit_kernel = ms_kernel;
  st.ms_phase[0] = hstats[4] * 1e-6f; st.ms_phase[1] = hstats[5] * 1e-6f;
  st.ms_phase[2] = hstats[6] * 1e-6f; st.ms_phase[3] = hstats[7] * 1e-6f;
  out.num_states = hctl[2]; out.num_arcs = hctl[3];
  out.has_start = true; out.start = 0;
  out.props = props::of_compose(fa.props, fb.props);
  B200_CUDA(cudaEventRecord(ev1, s));
  if (opt.connect) {
    uint64_t launches = 0;
    DevFst trimmed = connect_device(out, true, &launches, s);
    st.kernel_launches += launches;
    out = std::move(trimmed);
  }
  st.states_out = out.num_states; st.arcs_out = out.num_arcs;
  B200_CUDA(cudaEventRecord(ev2, s));
  B200_CUDA(cudaStreamSynchronize(s));
  B200_CUDA(cudaEventElapsedTime(&st.ms_expand, ev0, ev1));
  B200_CUDA(cudaEventElapsedTime(&st.ms_connect, ev1, ev2));
  cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev2);
  *result = std::move(out);
  return true;
}

DevFst compose_device(const DevFst& a, const DevFst& b, const ComposeOptions& opt, ComposeStats* stats,
                      cudaStream_t s) {
  const char* impl = std::getenv("B200_COMPOSE_IMPL");
  bool want_waves = impl && std::string(impl) == "waves";
  if (!want_waves) {
    DevFst out(s);
    if (compose_device_coop(a, b, opt, stats, s, &out)) return out;
  }
  return compose_device_waves(a, b, opt, stats, s);
}

}  // namespace b200
