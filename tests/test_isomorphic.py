"""fst_isomorphic (rustfst-ffi/src/algorithms/isomorphic.rs:11-30), host side: the reference's Python KATs
(rustfst-python/tests/algorithms/test_isomorphic.py) and renumbering / reordering properties."""
import numpy as np
import pytest

import rustfst_b200 as R
from tests.parity_utils import random_fst


def test_isomorphic_kats_from_the_reference_python_tests():
    f1 = R.VectorFst()
    s1, s2 = f1.add_state(), f1.add_state()
    f1.set_start(s1)
    f1.set_final(s2)
    f1.add_tr(s1, R.Tr(12, 25, None, s2))
    f2 = f1.copy()
    assert f1.isomorphic(f2)
    f2.add_tr(s1, R.Tr(1, 2, None, s2))
    assert not f1.isomorphic(f2)
    # test_isomorphic_2: the same machine with the two states swapped
    g = R.VectorFst()
    t1, t2 = g.add_state(), g.add_state()
    g.set_start(t2)
    g.set_final(t1)
    g.add_tr(t2, R.Tr(12, 25, None, t1))
    assert f1.isomorphic(g)


def _permuted(d, rng, shuffle_arcs=True):
    n = len(d["finals"])
    perm = rng.permutation(n)
    inv = np.argsort(perm)
    off = d["offsets"].astype(np.int64)
    rows = []
    for s_new in range(n):
        s_old = inv[s_new]
        row = d["arcs"][off[s_old]:off[s_old + 1]].copy()
        row["nextstate"] = perm[row["nextstate"]]
        if shuffle_arcs:
            rng.shuffle(row)
        rows.append(row)
    noff = np.zeros(n + 1, dtype=np.uint32)
    noff[1:] = np.cumsum([len(r) for r in rows])
    arcs = np.concatenate(rows) if rows else d["arcs"]
    return {"offsets": noff, "arcs": arcs, "finals": d["finals"][inv], "start": int(perm[d["start"]]), "props": 0}


@pytest.mark.parametrize("seed", range(20))
def test_renumbered_and_reordered_copies_are_isomorphic_and_perturbed_ones_are_not(seed):
    rng = np.random.default_rng(6000 + seed)
    # distinct (ilabel, olabel) per state keeps the machines deterministic "as unweighted automata", the domain on
    # which the reference's single-permutation pairing is exact (isomorphic.rs:118-126)
    d = random_fst(rng, int(rng.integers(2, 25)), 3, 40, eps_prob=0.0, cyclic=True)
    off = d["offsets"].astype(np.int64)
    for s in range(len(d["finals"])):
        row = d["arcs"][off[s]:off[s + 1]]
        row["ilabel"] = 1 + np.arange(len(row))
    a = R.VectorFst.from_csr(d["offsets"], d["arcs"], d["finals"], d["start"], 0)
    p = _permuted(d, rng)
    b = R.VectorFst.from_csr(p["offsets"], p["arcs"], p["finals"], p["start"], 0)
    assert a.isomorphic(b) and b.isomorphic(a)
    if b.num_trs(p["start"]):  # only the part reachable from the start state is compared
        q = dict(p, arcs=p["arcs"].copy())
        q["arcs"]["weight"] += 1.0  # far more than KDELTA
        c = R.VectorFst.from_csr(q["offsets"], q["arcs"], q["finals"], q["start"], 0)
        assert not a.isomorphic(c)
        q2 = dict(p, arcs=p["arcs"].copy())
        q2["arcs"]["weight"] += 1e-4  # within KDELTA
        assert a.isomorphic(R.VectorFst.from_csr(q2["offsets"], q2["arcs"], q2["finals"], q2["start"], 0))


def test_start_state_cases_and_nondeterminism_error():
    e1, e2 = R.VectorFst(), R.VectorFst()
    assert e1.isomorphic(e2)
    e2.add_state()
    e2.set_start(0)
    assert not e1.isomorphic(e2)
    # two arcs with the same labels and weight to different targets, paired in the "wrong" order on one side
    a, b = R.VectorFst(), R.VectorFst()
    for f in (a, b):
        for _ in range(3):
            f.add_state()
        f.set_start(0)
    a.set_final(1, 1.0); b.set_final(2, 1.0)
    a.add_tr(0, R.Tr(1, 1, 0.5, 1)); a.add_tr(0, R.Tr(1, 1, 0.5, 2))
    b.add_tr(0, R.Tr(1, 1, 0.5, 1)); b.add_tr(0, R.Tr(1, 1, 0.5, 2))
    with pytest.raises(ValueError, match="Non-determinism as an unweighted automaton"):
        a.isomorphic(b)
