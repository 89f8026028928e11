"""isomorphic() on the device (csrc/iso.cu) against the sequential host restatement of isomorphic.rs:49-160: the same
answers on renumbered / reordered / perturbed machines, `undecided` exactly where the reference's answer depends on its
visiting order, and as a download-free verifier of composed lattices."""
import os

import numpy as np
import pytest

from tests.parity_utils import both_from_dict, random_fst
from tests.test_isomorphic import _permuted

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _force_device_path():
    os.environ["B200_ISO_DEVICE"] = "1"   # fst_isomorphic pairs small machines on the host unless told otherwise
    yield
    del os.environ["B200_ISO_DEVICE"]


@pytest.mark.parametrize("seed", range(12))
def test_device_pairing_agrees_with_the_host_restatement(seed):
    import rustfst_b200 as R
    rng = np.random.default_rng(7000 + seed)
    d = random_fst(rng, int(rng.integers(2, 60)), 4, 40, eps_prob=0.0, cyclic=True)
    off = d["offsets"].astype(np.int64)
    for s in range(len(d["finals"])):
        row = d["arcs"][off[s]:off[s + 1]]
        row["ilabel"] = 1 + np.arange(len(row))   # deterministic as an unweighted automaton
    a = R.VectorFst.from_csr(d["offsets"], d["arcs"], d["finals"], d["start"], 0)
    p = _permuted(d, rng)
    b = R.VectorFst.from_csr(p["offsets"], p["arcs"], p["finals"], p["start"], 0)
    da, db = R.DeviceFst.upload(a), R.DeviceFst.upload(b)
    assert da.isomorphic(db) is True and db.isomorphic(da) is True
    assert a.isomorphic(b)
    if b.num_trs(p["start"]):
        q = dict(p, arcs=p["arcs"].copy())
        q["arcs"]["weight"] += 1.0
        c = R.VectorFst.from_csr(q["offsets"], q["arcs"], q["finals"], q["start"], 0)
        assert da.isomorphic(R.DeviceFst.upload(c)) is False and not a.isomorphic(c)
        q2 = dict(p, arcs=p["arcs"].copy())
        q2["arcs"]["weight"] += 1e-4   # within KDELTA
        c2 = R.VectorFst.from_csr(q2["offsets"], q2["arcs"], q2["finals"], q2["start"], 0)
        assert da.isomorphic(R.DeviceFst.upload(c2)) is True
        q3 = dict(p, finals=p["finals"].copy())
        q3["finals"][p["start"]] = 3.5 if not np.isfinite(q3["finals"][p["start"]]) else np.inf
        c3 = R.VectorFst.from_csr(q3["offsets"], q3["arcs"], q3["finals"], q3["start"], 0)
        assert da.isomorphic(R.DeviceFst.upload(c3)) is False


def test_order_dependent_cases_are_left_to_the_host():
    import rustfst_b200 as R
    a, b = R.VectorFst(), R.VectorFst()
    for f in (a, b):
        for _ in range(3):
            f.add_state()
        f.set_start(0)
    a.set_final(1, 1.0); b.set_final(2, 1.0)
    a.add_tr(0, R.Tr(1, 1, 0.5, 1)); a.add_tr(0, R.Tr(1, 1, 0.5, 2))
    b.add_tr(0, R.Tr(1, 1, 0.5, 1)); b.add_tr(0, R.Tr(1, 1, 0.5, 2))
    assert R.DeviceFst.upload(a).isomorphic(R.DeviceFst.upload(b)) is None
    with pytest.raises(ValueError, match="Non-determinism as an unweighted automaton"):
        a.isomorphic(b)   # device undecided -> host restatement -> the reference's error
    e1, e2 = R.VectorFst(), R.VectorFst()
    assert e1.isomorphic(e2)
    e2.add_state(); e2.set_start(0)
    assert not e1.isomorphic(e2)


def test_composed_lattice_is_verified_without_leaving_the_device():
    import rustfst_b200 as R
    from rustfst_b200 import synth
    a1, a2 = synth.workload("C2", scale=0.25)
    pa, _ = both_from_dict(a1)
    pb, _ = both_from_dict(a2)
    d1, d2 = R.DeviceFst.upload(pa), R.DeviceFst.upload(pb)
    r1, _ = R.device_compose(d1, d2)
    r2, _ = R.device_compose(d1, d2)
    assert r1.isomorphic(r2) is True
    # a renumbered copy: the lattice has parallel arcs that are equal as an unweighted automaton, so the reference's
    # single-permutation pairing (isomorphic.rs:118-126) may give up with its non-determinism error; the device then
    # says "undecided" and fst_isomorphic hands the reference's answer — true or that error — through
    h = r1.download()
    off, arcs, fin, start = h.to_csr()
    rng = np.random.default_rng(5)
    p = _permuted({"offsets": off, "arcs": arcs, "finals": fin, "start": start}, rng, shuffle_arcs=True)
    hp = R.VectorFst.from_csr(p["offsets"], p["arcs"], p["finals"], p["start"], 0)
    on_device = r1.isomorphic(R.DeviceFst.upload(hp))
    assert on_device in (True, None)
    try:
        assert h.isomorphic(hp) is True
    except ValueError as e:
        assert on_device is None and "Non-determinism as an unweighted automaton" in str(e)
    # the same machine with one final weight changed is not isomorphic to itself
    fin2 = fin.copy()
    k = int(np.flatnonzero(np.isfinite(fin2))[0])
    fin2[k] += 2.0
    hq = R.VectorFst.from_csr(off, arcs, fin2, start, 0)
    assert r1.isomorphic(R.DeviceFst.upload(hq)) in (False, None)
    try:
        assert not h.isomorphic(hq)
    except ValueError as e:
        assert "Non-determinism as an unweighted automaton" in str(e)
