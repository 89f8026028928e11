"""Pins the oracle's n > 1 shortest paths (oracle.hpp: shortest_distance, reverse, n_shortest_path) with the criterion
the reference's own test uses (rustfst/src/tests_openfst/algorithms/shortest_path.rs:62-92; the OpenFst goldens it
compares against are not in the tree): (1) same number of paths as the true n best, (2) paths at the same position
have the same weight, (3) every path exists, with that weight, in the input.  The "true n best" come from a
brute-force enumeration of every successful path of small machines."""
import numpy as np
import pytest

from tests import oracle_lib as O
from tests.parity_utils import FIXTURES, golden_path, random_fst


def csr_rows(o):
    off, arcs, fin = o.to_csr()
    return off, arcs, fin, o.start


def all_paths(o, limit=200000, max_len=64):
    """Every successful path (ilabels, olabels, weight) of an acyclic machine, by DFS from the start state."""
    off, arcs, fin, start = csr_rows(o)
    out = []
    if start is None:
        return out
    stack = [(start, (), (), np.float32(0.0))]
    while stack:
        s, il, ol, w = stack.pop()
        if np.isfinite(fin[s]):
            out.append((il, ol, np.float32(w + fin[s])))
            if len(out) > limit:
                raise RuntimeError("too many paths")
        if len(il) >= max_len:
            raise RuntimeError("path too long (cycle?)")
        for k in range(int(off[s]), int(off[s + 1])):
            a = arcs[k]
            stack.append((int(a["nextstate"]), il + (int(a["ilabel"]),), ol + (int(a["olabel"]),),
                          np.float32(w + a["weight"])))
    return out


def result_paths(r):
    """Paths of an n-shortest result in the order of the start state's arcs (the reference's paths_iter order)."""
    off, arcs, fin, start = csr_rows(r)
    paths = []
    if start is None:
        return paths
    for k in range(int(off[start]), int(off[start + 1])):
        a = arcs[k]
        il, ol, w = [], [], np.float32(a["weight"])
        if a["ilabel"] or a["olabel"]:
            il.append(int(a["ilabel"])); ol.append(int(a["olabel"]))
        s = int(a["nextstate"])
        while True:
            n = int(off[s + 1] - off[s])
            if n == 0:
                assert np.isfinite(fin[s])
                w = np.float32(w + fin[s])
                break
            assert n == 1 and not np.isfinite(fin[s]), "interior states of an n-best tree carry exactly one arc"
            b = arcs[int(off[s])]
            il.append(int(b["ilabel"])); ol.append(int(b["olabel"]))
            w = np.float32(w + b["weight"])
            s = int(b["nextstate"])
        paths.append((tuple(il), tuple(ol), w))
    return paths


def strip_eps(p):
    return tuple(x for x in p if x != 0)


def check_nbest(o, n):
    r = O.shortest_path(o, nshortest=n)
    truth = sorted(all_paths(o), key=lambda p: p[2])
    got = result_paths(r)
    assert len(got) == min(n, len(truth))
    for g, t in zip(got, truth):
        assert g[2] == t[2], f"weight at the same position differs: {g[2]} vs {t[2]}"
    # every returned path exists in the input with that weight (labels compared with epsilons removed: the result
    # spells the superfinal transitions as 0:0 arcs)
    pool = {}
    for il, ol, w in truth:
        pool.setdefault((strip_eps(il), strip_eps(ol)), []).append(w)
    for il, ol, w in got:
        assert w in pool.get((strip_eps(il), strip_eps(ol)), []), "returned path is not a path of the input"
    # distinct derivations: n-best without `unique` returns distinct PATHS (state sequences), so at most as many
    # copies of a label sequence as the input has
    return r


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("n", [2, 3, 5])
def test_nbest_matches_bruteforce_on_random_dags(seed, n):
    rng = np.random.default_rng(1000 + seed)
    d = random_fst(rng, n_states=int(rng.integers(4, 12)), max_arcs=3, n_labels=4, eps_prob=0.1, cyclic=False)
    o = O.OFst.from_csr(d["offsets"].astype(np.uint64), d["arcs"], d["finals"], d["start"], d["props"])
    check_nbest(o, n)


def test_nbest_of_empty_and_startless_machines():
    o = O.OFst()
    assert O.shortest_path(o, nshortest=3).num_states == 0
    o.add_state()
    assert O.shortest_path(o, nshortest=3).num_states == 0  # no start state: shortest_path.rs:427-434
    o.set_start(0)
    assert O.shortest_path(o, nshortest=3).num_states == 0  # no final state: distance of the superinitial is zero()


def test_nbest_kat_two_paths():
    """The doc example of shortest_path.rs (n = 2 picture): two parallel paths, both must come back, best first."""
    o = O.OFst()
    for _ in range(4):
        o.add_state()
    o.set_start(0)
    o.set_final(3, 0.5)
    o.add_tr(0, 1, 1, 1.0, 1)
    o.add_tr(0, 2, 2, 3.0, 2)
    o.add_tr(1, 3, 3, 1.0, 3)
    o.add_tr(2, 4, 4, 0.25, 3)
    r = check_nbest(o, 2)
    p = result_paths(r)
    assert [strip_eps(x[0]) for x in p] == [(1, 3), (2, 4)]
    assert [float(x[2]) for x in p] == [2.5, 3.75]
    # n = 1 through the n-best route is not taken (nshortest == 1 uses single_shortest_path)
    assert O.shortest_path(o, nshortest=1).num_states == 3


@pytest.mark.parametrize("name", FIXTURES)
def test_nbest_runs_on_fixture_machines(name):
    """Cyclic, epsilon-rich fixtures exercise the LIFO / SCC queue branches of shortest_distance and the `enqueued`
    quirk (shortest_distance.rs:224): results must be trees of n paths with non-decreasing weights."""
    o = O.OFst.from_path(golden_path(name, "raw"))
    for n in (2, 4):
        r = O.shortest_path(o, nshortest=n)
        w = [float(p[2]) for p in result_paths(r)]
        assert len(w) <= n
        assert all(w[i] <= w[i + 1] + 1e-3 for i in range(len(w) - 1))
        if w:
            # the best of the n-best equals the single shortest path's weight
            r1 = O.shortest_path(o, nshortest=1)
            off, arcs, fin = r1.to_csr()
            total = float(np.float32(arcs["weight"].astype(np.float64).sum() + fin[np.isfinite(fin)].sum()))
            assert abs(total - w[0]) < 1e-3


# ---- unique = true: determinize_with_distance of the reversed machine, then the same n-best search ----------------

def check_unique_nbest(o, n, shortest_path=None):
    """The reference's criterion (tests_openfst/algorithms/shortest_path.rs:62-92) against brute force over DISTINCT
    label strings: determinization treats epsilon as an ordinary symbol (determinize_fsa_op.rs:62-82), so two paths are
    the same string iff their raw label sequences are equal, and a string weighs what its lightest path weighs."""
    r = (shortest_path or O.shortest_path)(o, nshortest=n, unique=True)
    best = {}
    for il, ol, w in all_paths(o):
        assert il == ol
        best[il] = min(best.get(il, np.float32(np.inf)), w)
    truth = sorted(best.items(), key=lambda kv: kv[1])
    got = result_paths(r)
    assert len(got) == min(n, len(truth))
    for g, t in zip(got, truth):  # residual weights are quantised to multiples of delta = 1e-6: compare approximately
        assert abs(float(g[2]) - float(t[1])) < 1e-3, f"weight at the same position differs: {g[2]} vs {t[1]}"
    seen = set()
    for il, ol, w in got:
        assert il == ol
        key = strip_eps(il)
        cands = [wt for s, wt in best.items() if strip_eps(s) == key]
        assert any(abs(float(w) - float(c)) < 1e-3 for c in cands), "returned path is not a path of the input"
        # the result spells strings with its own 0:0 bridge arcs, so distinctness is checked on (string, weight)
        assert (il, float(w)) not in seen
        seen.add((il, float(w)))
    return r


@pytest.mark.parametrize("seed", range(16))
@pytest.mark.parametrize("n", [2, 3, 6])
def test_unique_nbest_matches_bruteforce_on_random_acyclic_acceptors(seed, n):
    rng = np.random.default_rng(4000 + seed)
    # few labels and many arcs: plenty of paths that spell the same string
    d = random_fst(rng, n_states=int(rng.integers(4, 11)), max_arcs=4, n_labels=2, eps_prob=0.1, acceptor=True, cyclic=False)
    o = O.OFst.from_csr(d["offsets"].astype(np.uint64), d["arcs"], d["finals"], d["start"], d["props"])
    check_unique_nbest(o, n)


def test_unique_nbest_kat_duplicate_strings():
    """Three paths, two of which spell `1 2`: unique n-best returns `1 2` once (with its lighter weight) and `1 3`."""
    o = O.OFst()
    for _ in range(4):
        o.add_state()
    o.set_start(0)
    o.set_final(3, 0.0)
    o.add_tr(0, 1, 1, 1.0, 1)
    o.add_tr(0, 1, 1, 2.0, 2)
    o.add_tr(1, 2, 2, 1.0, 3)   # 1 2 / 2.0
    o.add_tr(2, 2, 2, 0.5, 3)   # 1 2 / 2.5  (same string, heavier)
    o.add_tr(2, 3, 3, 1.0, 3)   # 1 3 / 3.0
    plain = [(strip_eps(p[0]), float(p[2])) for p in result_paths(O.shortest_path(o, nshortest=3))]
    assert plain == [((1, 2), 2.0), ((1, 2), 2.5), ((1, 3), 3.0)]
    uniq = [(strip_eps(p[0]), round(float(p[2]), 4)) for p in result_paths(check_unique_nbest(o, 3))]
    assert uniq == [((1, 2), 2.0), ((1, 3), 3.0)]


def test_unique_nbest_refuses_transducers():
    """determinize_fsa_op.rs:137-139: the reversed machine must be an acceptor."""
    o = O.OFst()
    o.add_state(); o.add_state()
    o.set_start(0); o.set_final(1, 0.0)
    o.add_tr(0, 1, 2, 1.0, 1)
    with pytest.raises(Exception, match="expected acceptor"):
        O.shortest_path(o, nshortest=2, unique=True)
