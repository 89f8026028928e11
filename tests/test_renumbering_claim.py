"""DESIGN.md §8 proposes to expand a composition asynchronously (provisional state ids in discovery order) and to
recover the reference's numbering afterwards.  That only works if the reference's numbering — first-emission order
under a FIFO BFS (lazy_fst.rs:226-269, state_table.rs:49-59) — is a function of the finished graph alone.  This test
checks the claim on oracle results: scramble the state ids of an untrimmed composition, renumber by BFS first-visit
order over (source id, arc position), and get the original back bit for bit."""
import numpy as np
import pytest

from tests import oracle_lib as O
from tests.parity_utils import FIXTURES, golden_path, random_fst


def scramble(off, arcs, fin, start, rng):
    n = len(fin)
    perm = rng.permutation(n)          # old id -> scrambled id
    inv = np.argsort(perm)             # scrambled id -> old id
    deg = np.diff(off.astype(np.int64))
    noff = np.zeros(n + 1, dtype=np.int64)
    noff[1:] = np.cumsum(deg[inv])
    narcs = np.zeros(len(arcs), dtype=arcs.dtype)
    for s_new in range(n):
        s_old = inv[s_new]
        row = arcs[int(off[s_old]):int(off[s_old + 1])].copy()
        row["nextstate"] = perm[row["nextstate"]]
        narcs[int(noff[s_new]):int(noff[s_new + 1])] = row
    return noff, narcs, fin[inv], int(perm[start])


def canonical_renumber(off, arcs, fin, start):
    """BFS from the start state; a state's id is the rank of its first visit in (wave, source id, arc position) order."""
    n = len(fin)
    new_id = np.full(n, -1, dtype=np.int64)
    order = [start]
    new_id[start] = 0
    head = 0
    while head < len(order):
        s = order[head]; head += 1
        for k in range(int(off[s]), int(off[s + 1])):
            t = int(arcs[k]["nextstate"])
            if new_id[t] < 0:
                new_id[t] = len(order)
                order.append(t)
    assert len(order) == n, "an untrimmed composition only contains accessible states"
    order = np.array(order)
    deg = np.diff(off.astype(np.int64))
    coff = np.zeros(n + 1, dtype=np.int64)
    coff[1:] = np.cumsum(deg[order])
    carcs = np.zeros(len(arcs), dtype=arcs.dtype)
    for i, s in enumerate(order):
        row = arcs[int(off[s]):int(off[s + 1])].copy()
        row["nextstate"] = new_id[row["nextstate"]]
        carcs[int(coff[i]):int(coff[i + 1])] = row
    return coff, carcs, fin[order]


def check(oa, ob, seed):
    try:
        c = O.compose(oa, ob, connect=False)
    except O.OracleError:
        return
    if c.start is None or c.num_states < 2:
        return
    off, arcs, fin = c.to_csr()
    rng = np.random.default_rng(seed)
    soff, sarcs, sfin, sstart = scramble(off, arcs, fin, c.start, rng)
    coff, carcs, cfin = canonical_renumber(soff, sarcs, sfin, sstart)
    assert np.array_equal(coff, off.astype(np.int64))
    assert carcs.tobytes() == arcs.tobytes()
    assert cfin.tobytes() == fin.tobytes()


@pytest.mark.parametrize("name", FIXTURES)
def test_fixture_compositions_are_recovered_from_scrambled_ids(name):
    check(O.OFst.from_path(golden_path(name, "raw")), O.OFst.from_path(golden_path(name, "compose")), 1)


@pytest.mark.parametrize("seed", range(6))
def test_random_compositions_are_recovered_from_scrambled_ids(seed):
    rng = np.random.default_rng(500 + seed)
    a = random_fst(rng, 25, 4, 5, eps_prob=0.15, sort="olabel")
    b = random_fst(rng, 25, 4, 5, eps_prob=0.15, sort="ilabel")
    oa = O.OFst.from_csr(a["offsets"].astype(np.uint64), a["arcs"], a["finals"], a["start"], a["props"])
    ob = O.OFst.from_csr(b["offsets"].astype(np.uint64), b["arcs"], b["finals"], b["start"], b["props"])
    check(oa, ob, seed)
