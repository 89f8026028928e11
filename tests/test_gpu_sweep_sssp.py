"""The in-order sweep (sssp.cu: k_relax_sweep) that takes over on TOP_SORTED machines when the label-correcting waves go
over their visit budget (deep DAGs whose arcs skip levels).  Results must be the oracle's bit for bit, for the single
shortest path and for the forward distances behind nshortest > 1."""
import os

import numpy as np
import pytest

from tests import oracle_lib as O
from tests.parity_utils import assert_same, both_from_dict, random_fst

pytestmark = pytest.mark.gpu


@pytest.fixture
def zero_budget():
    old = os.environ.get("B200_RELAX_VISIT_BUDGET")
    os.environ["B200_RELAX_VISIT_BUDGET"] = "0"
    yield
    if old is None:
        del os.environ["B200_RELAX_VISIT_BUDGET"]
    else:
        os.environ["B200_RELAX_VISIT_BUDGET"] = old


@pytest.mark.parametrize("seed", range(10))
def test_sweep_forced_on_random_dags(seed, zero_budget):
    """Budget 0 sends every top-sorted DAG through the sweep: unreachable states, states without arcs, parallel arcs,
    epsilon arcs, arcs that reach beyond the shared-memory ring (odd seeds: thousands of states).  Machines that are
    acyclic but not top-sorted (TopOrderQueue) keep the waves."""
    import rustfst_b200 as R
    rng = np.random.default_rng(31000 + seed)
    n_states = int(rng.integers(2, 300)) if seed % 2 == 0 else int(rng.integers(9000, 30000))
    d = random_fst(rng, n_states, 5, 6, eps_prob=0.1, cyclic=False, weight_grid=(seed % 3 < 2))
    scrambled = seed % 5 == 4
    if scrambled:  # scramble the numbering: not TOP_SORTED any more, the order comes from the DFS
        n = d["num_states"]
        perm = rng.permutation(n)
        off, arcs, fin = d["offsets"], d["arcs"], d["finals"]
        inv = np.argsort(perm)
        rows = [arcs[off[inv[s]]:off[inv[s] + 1]].copy() for s in range(n)]
        for r in rows:
            r["nextstate"] = perm[r["nextstate"]]
        d = dict(d, offsets=np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.uint32),
                 arcs=np.concatenate(rows) if rows else arcs, finals=fin[inv], start=int(perm[0]))
        o = O.OFst.from_csr(d["offsets"].astype(np.uint64), d["arcs"], d["finals"], d["start"], 0)
        o.compute_props()
        d["props"] = o.props
    p, o = both_from_dict(d)
    got, st = R.shortestpath_with_stats(p)
    assert st["path"] != 1, st
    if st["path"] == 0:
        assert st["sweep"] == (0 if scrambled else 1), st
    assert_same(got, O.shortest_path(o), f"sweep sssp seed={seed}")
    got, st = R.shortestpath_with_stats(p, R.ShortestPathConfig(nshortest=3))
    assert_same(got, O.shortest_path(o, nshortest=3), f"sweep n-best seed={seed}")


def test_window_dag_goes_over_the_visit_budget_on_its_own():
    """SURVEY.md 8d's acyclic acceptor with targets up to 1000 ids ahead, shrunk: label-correcting waves revisit states
    hundreds of times, the budget (4 visits per state) trips and the in-order sweep finishes the job."""
    import rustfst_b200 as R
    from rustfst_b200 import synth
    g = synth.window_dag(200_000, 2_000_000, 1000, 6, window=1000)
    p, o = both_from_dict(g)
    got, st = R.shortestpath_with_stats(p)
    assert st["sweep"] == 1 and st["path"] == 0, st
    assert_same(got, O.shortest_path(o), "window DAG")
    # a layered lattice stays on the wave kernel
    g = synth.layered_acceptor(200_000, 2_000_000, 1000, 6, 40)
    p, o = both_from_dict(g)
    got, st = R.shortestpath_with_stats(p)
    assert st["sweep"] == 0 and st["path"] == 0, st
    assert_same(got, O.shortest_path(o), "layered lattice")
