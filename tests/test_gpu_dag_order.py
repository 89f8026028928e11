"""The device-side TopOrderQueue order (csrc/dag_order.cu) against the oracle's sequential DFS
(oracle.hpp restating top_sort.rs:12-61 + dfs_visit.rs:97-187): the two must give the SAME order array, not just
some topological order, because the order breaks ties between equally good parents in the shortest path."""
import ctypes as C

import numpy as np
import pytest

from tests import oracle_lib as O
from tests.parity_utils import assert_same, both_from_dict

pytestmark = pytest.mark.gpu


def _device_order(p):
    from rustfst_b200.ffi import check_ffi_error, lib
    n = p.num_states()
    order = np.zeros(max(1, n), dtype=np.uint32)
    ok, ms = C.c_int32(), C.c_float()
    check_ffi_error(lib.b200_dag_top_order_device(p.ptr, order.ctypes.data, C.byref(ok), C.byref(ms)), "dag order")
    return bool(ok.value), order[:n], ms.value


def _dag(rng, n, max_deg, window=None, permute=True, start=None, parallel_arcs=False):
    """Random DAG as a CSR dict whose property word says ACYCLIC but NOT top-sorted: arcs go from a lower to a higher
    position of a hidden topological numbering, then the state ids are shuffled."""
    from rustfst_b200 import props as P
    from rustfst_b200.fst import TR_DTYPE
    perm = rng.permutation(n) if permute else np.arange(n)
    rows = [[] for _ in range(n)]
    for i in range(n - 1):
        k = int(rng.integers(0, max_deg + 1))
        for _ in range(k):
            hi = n if window is None else min(n, i + 1 + window)
            j = int(rng.integers(i + 1, hi))
            rows[perm[i]].append(perm[j])
            if parallel_arcs and rng.random() < 0.2:
                rows[perm[i]].append(perm[j])
    offsets = np.zeros(n + 1, dtype=np.uint32)
    offsets[1:] = np.cumsum([len(r) for r in rows])
    arcs = np.zeros(int(offsets[-1]), dtype=TR_DTYPE)
    arcs["nextstate"] = np.concatenate([np.array(r, dtype=np.uint32) for r in rows]) if len(arcs) else []
    arcs["ilabel"] = rng.integers(1, 50, size=len(arcs)); arcs["olabel"] = arcs["ilabel"]
    arcs["weight"] = rng.integers(0, 16, size=len(arcs)) / 4.0  # coarse grid: many exact ties
    finals = np.where(rng.random(n) < 0.2, rng.integers(0, 16, size=n) / 4.0, np.inf).astype(np.float32)
    pr = P.ACYCLIC | P.INITIAL_ACYCLIC | P.WEIGHTED | P.ACCEPTOR
    return {"offsets": offsets, "arcs": arcs, "finals": finals, "start": int(perm[0] if start is None else start),
            "props": int(pr), "num_states": n}


def _check(d, what):
    p, o = both_from_dict(d)
    kind, expected, _ = O.queue_plan(o)
    assert kind == 1, f"{what}: the oracle should pick the TopOrderQueue"
    ok, got, _ = _device_order(p)
    assert ok, f"{what}: the device path declined"
    assert np.array_equal(got, expected), f"{what}: order differs at {np.flatnonzero(got != expected)[:10]}"
    return p, o


@pytest.mark.parametrize("seed", range(6))
def test_random_dags_order_identical_to_the_sequential_dfs(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 400))
    _check(_dag(rng, n, int(rng.integers(1, 6)), parallel_arcs=seed % 2 == 0), f"random dag {seed}")


def test_small_shapes():
    from rustfst_b200 import props as P
    from rustfst_b200.fst import TR_DTYPE

    def mk(n, edges, start):
        rows = [[] for _ in range(n)]
        for a, b in edges:
            rows[a].append(b)
        off = np.zeros(n + 1, dtype=np.uint32); off[1:] = np.cumsum([len(r) for r in rows])
        arcs = np.zeros(int(off[-1]), dtype=TR_DTYPE)
        if len(arcs):
            arcs["nextstate"] = np.concatenate([np.array(r, dtype=np.uint32) for r in rows if r])
        arcs["ilabel"] = 1; arcs["olabel"] = 1
        fin = np.full(n, np.inf, dtype=np.float32); fin[n - 1] = 0.0
        return {"offsets": off, "arcs": arcs, "finals": fin, "start": start,
                "props": int(P.ACYCLIC | P.INITIAL_ACYCLIC | P.ACCEPTOR), "num_states": n}

    _check(mk(1, [], 0), "single state")
    _check(mk(4, [(0, 1), (0, 2), (1, 3), (2, 3)], 0), "diamond")
    _check(mk(4, [(0, 2), (0, 1), (1, 3), (2, 3)], 0), "diamond, arcs swapped")
    # a later DFS root (5) has an arc into an earlier root's tree (3): cross arc, 3 stays a root
    _check(mk(6, [(0, 1), (5, 3), (3, 4), (2, 4)], 0), "several roots, cross arc")
    # the start state is not state 0 and state 0 is reachable from a later root only
    _check(mk(5, [(2, 3), (3, 4), (1, 0), (0, 4)], 2), "start in the middle")
    # skip-level arcs: the lexicographically smallest path is the longer one
    _check(mk(5, [(0, 1), (0, 3), (1, 2), (2, 3), (3, 4), (0, 4)], 0), "long path first")
    _check(mk(5, [(0, 3), (0, 1), (1, 2), (2, 3), (3, 4), (0, 4)], 0), "short path first")


def test_wide_skip_level_dag_and_unreachable_states():
    rng = np.random.default_rng(11)
    d = _dag(rng, 20_000, 6, window=500)
    _check(d, "window dag")
    d = _dag(rng, 5_000, 3, start=int(rng.integers(0, 5_000)))  # most states unreachable from the start: many roots
    _check(d, "many roots")


def test_deep_dags_take_the_full_ancestor_table():
    """More than 1024 Kahn levels: the first attempt (six rows of the 2^j-ancestor table) gives up and the call is
    repeated with all log2(n) rows; ancestors 16 and more levels up are then read from the side table."""
    rng = np.random.default_rng(12)
    d = _dag(rng, 6_000, 3, window=3)       # arcs reach at most three positions ahead: thousands of levels
    _check(d, "narrow deep dag")
    d = _dag(rng, 3_000, 2, window=40)      # a few hundred levels, comparisons across many depths
    _check(d, "medium deep dag")


def test_layered_lattice_and_shortest_path_through_the_device_order():
    import rustfst_b200 as R
    from rustfst_b200 import props as PR
    from rustfst_b200 import synth
    g = synth.layered_acceptor(200_000, 2_000_000, 1000, 6, 50)
    g = dict(g, props=g["props"] & ~(PR.TOP_SORTED | PR.NOT_TOP_SORTED))
    p, o = _check(g, "layered lattice")
    sp, st = R.shortestpath_with_stats(p)
    assert st["queue_kind"] == 1 and st["order_on_device"] == 1 and st["ms_queue_plan_host"] < 5.0
    assert_same(sp, O.shortest_path(o), "shortest path with the device order")
    # continuous weights: the certificate fails and the order-faithful fold consumes the same device order
    gc = synth.layered_acceptor(50_000, 500_000, 1000, 7, 20, continuous=True)
    gc = dict(gc, props=gc["props"] & ~(PR.TOP_SORTED | PR.NOT_TOP_SORTED))
    pc, oc = both_from_dict(gc)
    spc, stc = R.shortestpath_with_stats(pc)
    assert stc["order_on_device"] == 1
    assert_same(spc, O.shortest_path(oc), "near-tie weights with the device order")


def test_cyclic_and_very_deep_machines_fall_back_to_the_host_dfs():
    import rustfst_b200 as R
    from rustfst_b200 import props as P
    from rustfst_b200.fst import TR_DTYPE
    # a chain deeper than the device path's level limit
    n = 70_000
    arcs = np.zeros(n - 1, dtype=TR_DTYPE)
    arcs["ilabel"] = 1; arcs["olabel"] = 1; arcs["weight"] = 0.5; arcs["nextstate"] = np.arange(1, n)
    rev = np.arange(n)[::-1].copy()  # state ids run against the chain, so the machine is not top-sorted
    arcs["nextstate"] = rev[np.arange(1, n)]
    order_rows = np.argsort(rev[:-1])
    off = np.zeros(n + 1, dtype=np.uint32)
    cnt = np.zeros(n, dtype=np.int64); cnt[rev[:-1]] = 1
    off[1:] = np.cumsum(cnt)
    d = {"offsets": off, "arcs": arcs[order_rows], "finals": np.full(n, np.inf, dtype=np.float32), "start": int(rev[0]),
         "props": int(P.ACYCLIC | P.INITIAL_ACYCLIC | P.ACCEPTOR | P.WEIGHTED), "num_states": n}
    d["finals"][rev[-1]] = 1.0
    p, o = both_from_dict(d)
    ok, _, _ = _device_order(p)
    assert not ok
    sp, st = R.shortestpath_with_stats(p)
    assert st["order_on_device"] == 0 and st["queue_kind"] == 1
    assert_same(sp, O.shortest_path(o), "deep chain through the host DFS")
    # a cycle (with a property word that wrongly claims ACYCLIC the reference panics; the device path just declines)
    c = {"offsets": np.array([0, 1, 2], dtype=np.uint32), "arcs": np.zeros(2, dtype=TR_DTYPE),
         "finals": np.array([np.inf, 0.0], dtype=np.float32), "start": 0, "props": 0, "num_states": 2}
    c["arcs"]["nextstate"] = [1, 0]; c["arcs"]["ilabel"] = 1; c["arcs"]["olabel"] = 1
    pc, _ = both_from_dict(c)
    ok, _, _ = _device_order(pc)
    assert not ok
