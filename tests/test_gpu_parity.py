"""GPU parity tests proper: every result of the CUDA path (called through the C-ABI) is compared bit-for-bit
with the CPU oracle on the same inputs: state ids, arc order, labels, next states, weight bit patterns,
final weights and property words."""
import os

import numpy as np
import pytest

from tests import oracle_lib as O
from tests.parity_utils import FIXTURES, assert_same, both_from_dict, both_from_path, golden_path, random_fst

pytestmark = pytest.mark.gpu

FILTERS = [0, 1, 2, 3, 4, 5, 6]


def _compose_both(pa, oa, pb, ob, filt, connect):
    import rustfst_b200 as R
    try:
        o = O.compose(oa, ob, filter=filt, connect=connect)
    except O.OracleError as e:
        with pytest.raises(ValueError):
            R.compose_with_config(pa, pb, R.ComposeConfig(R.ComposeFilter(filt), connect))
        return None, str(e)
    p = R.compose_with_config(pa, pb, R.ComposeConfig(R.ComposeFilter(filt), connect))
    return (p, o), None


@pytest.mark.parametrize("name", FIXTURES)
@pytest.mark.parametrize("connect", [False, True])
def test_fixture_compose_all_filters(name, connect):
    """fst_NNN.get_fst() o fst_NNN.get_fst_compose() (rustfst-tests-data/main.cpp:1191-1194) under the seven
    filter settings of tests_openfst/algorithms/compose.rs:256-302, with and without connect."""
    pa, oa = both_from_path(golden_path(name, "raw"))
    pb, ob = both_from_path(golden_path(name, "compose"))
    for filt in FILTERS:
        res, err = _compose_both(pa, oa, pb, ob, filt, connect)
        if res:
            assert_same(res[0], res[1], f"{name} filter={filt} connect={connect}")


def test_fixture_default_compose_and_cross_pair():
    import rustfst_b200 as R
    for name in FIXTURES:
        pa, oa = both_from_path(golden_path(name, "raw"))
        pb, ob = both_from_path(golden_path(name, "compose"))
        try:
            o = O.compose(oa, ob)
        except O.OracleError:
            with pytest.raises(ValueError):
                pa.compose(pb)
            continue
        assert_same(pa.compose(pb), o, f"{name} default")
    # the literal fst_003 o fst_004 pairing of BASELINE.json configs[0]: no common labels -> empty after connect
    pa, oa = both_from_path(golden_path("fst_003", "raw"))
    pb, ob = both_from_path(golden_path("fst_004", "raw"))
    r = pa.compose(pb)
    assert r.num_states() == 0 and r.start() is None
    assert_same(r, O.compose(oa, ob), "fst_003 o fst_004")


@pytest.mark.parametrize("name", FIXTURES)
def test_fixture_shortest_path(name):
    """shortest_path(n=1) on every fixture input (cyclic, epsilon-rich: SccQueue / LIFO / TopOrder branches)."""
    import rustfst_b200 as R
    for which in ("raw", "compose"):
        p, o = both_from_path(golden_path(name, which))
        expected = O.shortest_path(o)
        got, st = R.shortestpath_with_stats(p)
        assert_same(got, expected, f"{name}/{which} shortest_path (path={st['path']} queue={st['queue_kind']})")
        # the order-faithful serial kernel must agree as well
        got2, _ = R.shortestpath_with_stats(p, force_serial=True)
        assert_same(got2, expected, f"{name}/{which} shortest_path serial")


def test_fixture_compose_then_shortest_path():
    """BASELINE.json configs[0] plumbing: compose + shortest_path chained on the device results."""
    import rustfst_b200 as R
    for name in FIXTURES:
        pa, oa = both_from_path(golden_path(name, "raw"))
        pb, ob = both_from_path(golden_path(name, "compose"))
        try:
            oc = O.compose(oa, ob)
        except O.OracleError:
            continue
        pc = pa.compose(pb)
        assert_same(pc.shortest_path(), O.shortest_path(oc), f"{name} compose+shortest_path")


def test_python_kats_through_cabi():
    """The reference's own FFI-level known answers (rustfst-python/tests/algorithms/test_compose.py:13-154,
    test_shortest_path.py:5-51), written with the mirrored API."""
    from rustfst_b200 import ComposeConfig, ComposeFilter, ShortestPathConfig, Tr, VectorFst

    def mk():
        fst1 = VectorFst()
        s1, s2, s3 = fst1.add_state(), fst1.add_state(), fst1.add_state()
        fst1.set_start(s1); fst1.set_final(s2); fst1.set_final(s3)
        fst1.add_tr(s1, Tr(1, 2, 1.0, s2)); fst1.add_tr(s1, Tr(1, 4, 2.0, s3)); fst1.add_tr(s2, Tr(3, 5, 2.0, s2))
        fst2 = VectorFst()
        s1, s2, s3 = fst2.add_state(), fst2.add_state(), fst2.add_state()
        fst2.set_start(s1); fst2.set_final(s3)
        fst2.add_tr(s1, Tr(2, 6, 1.0, s2)); fst2.add_tr(s2, Tr(5, 7, 2.5, s3)); fst2.add_tr(s3, Tr(5, 8, 1.5, s3))
        fst2.add_tr(s1, Tr(4, 9, 3.0, s3))
        e = VectorFst()
        s1, s2, s3, s4 = e.add_state(), e.add_state(), e.add_state(), e.add_state()
        e.set_start(s1); e.set_final(s3); e.set_final(s4)
        e.add_tr(s1, Tr(1, 6, 2.0, s2)); e.add_tr(s1, Tr(1, 9, 5.0, s3)); e.add_tr(s2, Tr(3, 7, 4.5, s4))
        e.add_tr(s4, Tr(3, 8, 3.5, s4))
        return fst1, fst2, e

    fst1, fst2, expected = mk()
    assert fst1.compose(fst2) == expected
    assert fst1.compose(fst2, ComposeConfig(ComposeFilter.TRIVIALFILTER, True)) == expected

    f = VectorFst()
    s1, s2, s3, s4 = f.add_state(), f.add_state(), f.add_state(), f.add_state()
    f.set_start(s1); f.set_final(s4, 2.0)
    f.add_tr(s1, Tr(1, 1, 3.0, s2)); f.add_tr(s2, Tr(2, 2, 2.0, s2)); f.add_tr(s2, Tr(3, 3, 4.0, s4))
    f.add_tr(s1, Tr(4, 4, 5.0, s3)); f.add_tr(s3, Tr(5, 5, 4.0, s4))
    e = VectorFst()
    s1, s2, s3 = e.add_state(), e.add_state(), e.add_state()
    e.set_start(s3); e.set_final(s1, 2.0)
    e.add_tr(s3, Tr(1, 1, 3.0, s2)); e.add_tr(s2, Tr(3, 3, 4.0, s1))
    assert f.shortest_path(ShortestPathConfig(1, True)) == e


@pytest.mark.parametrize("seed", range(12))
def test_random_compose_fuzz(seed):
    """Random small FSTs with epsilons and cycles; one or both sides sorted; all filters."""
    rng = np.random.default_rng(1000 + seed)
    sort_a = [None, "olabel", "olabel"][seed % 3]
    sort_b = ["ilabel", None, "ilabel"][seed % 3]
    da = random_fst(rng, int(rng.integers(1, 30)), 6, 6, eps_prob=0.25, sort=sort_a)
    db = random_fst(rng, int(rng.integers(1, 30)), 6, 6, eps_prob=0.25, sort=sort_b)
    pa, oa = both_from_dict(da)
    pb, ob = both_from_dict(db)
    for filt in FILTERS:
        for connect in (False, True):
            res, err = _compose_both(pa, oa, pb, ob, filt, connect)
            if res:
                assert_same(res[0], res[1], f"fuzz seed={seed} filter={filt} connect={connect}")


@pytest.mark.parametrize("seed", range(10))
def test_random_shortest_path_fuzz(seed):
    import rustfst_b200 as R
    rng = np.random.default_rng(2000 + seed)
    cyclic = seed % 2 == 0
    d = random_fst(rng, int(rng.integers(2, 60)), 5, 8, eps_prob=0.1, cyclic=cyclic, weight_grid=(seed % 4 < 2))
    p, o = both_from_dict(d)
    expected = O.shortest_path(o)
    got, st = R.shortestpath_with_stats(p)
    assert_same(got, expected, f"sssp fuzz seed={seed} path={st['path']}")
    got2, _ = R.shortestpath_with_stats(p, force_serial=True)
    assert_same(got2, expected, f"sssp fuzz serial seed={seed}")


@pytest.mark.parametrize("scale", [0.02, 0.25])
def test_synthetic_compose_and_sssp(scale):
    """BASELINE.json configs[1] (C2) shrunk so the oracle finishes in seconds; same generator as bench.py."""
    import rustfst_b200 as R
    from rustfst_b200 import synth
    a, b = synth.workload("C2", scale=scale)
    pa, oa = both_from_dict(a)
    pb, ob = both_from_dict(b)
    for connect in (False, True):
        o, ost = O.compose(oa, ob, connect=connect, want_stats=True)
        p, st = R.compose_with_stats(pa, pb, R.ComposeConfig(R.ComposeFilter.AUTOFILTER, connect))
        assert st["states_expanded"] == ost["states_expanded"] and st["arcs_emitted"] == ost["arcs_emitted"]
        assert_same(p, o, f"C2 x{scale} connect={connect}")
    # shortest path on the composed lattice (ACYCLIC known, TOP_SORTED unknown -> TopOrderQueue via DFS order)
    expected = O.shortest_path(o)
    got, st = R.shortestpath_with_stats(p)
    assert st["path"] == 0, "dyadic weights must pass the certificate"
    assert_same(got, expected, "C2 lattice shortest path")
    # and directly on the TOP_SORTED acceptor (StateOrderQueue)
    got, st = R.shortestpath_with_stats(pa)
    assert st["queue_kind"] == 0 and st["path"] == 0
    assert_same(got, O.shortest_path(oa), "acceptor shortest path")


def test_device_resident_api_matches_host_api():
    import rustfst_b200 as R
    from rustfst_b200 import synth
    a, b = synth.workload("C2", scale=0.05)
    pa, _ = both_from_dict(a)
    pb, _ = both_from_dict(b)
    da, db = R.DeviceFst.upload(pa), R.DeviceFst.upload(pb)
    dr, st = R.device_compose(da, db)
    host = pa.compose(pb)
    r = dr.download()
    assert r == host and r.properties == host.properties
    sp, _ = R.device_shortest_path(dr, plan_from=r)
    assert sp == host.shortest_path()


def test_tr_sort_then_compose_chain():
    """Compose output has unknown sortedness (mutate_properties.rs:151-184): composing it on the LEFT of an unsorted
    machine must fail until fst_tr_sort restores the bit."""
    import rustfst_b200 as R
    rng = np.random.default_rng(7)
    da = random_fst(rng, 12, 4, 5, sort="olabel", cyclic=False)
    db = random_fst(rng, 12, 4, 5, sort="ilabel", cyclic=False)
    dc = random_fst(rng, 12, 4, 5, sort=None, cyclic=False)
    pa, oa = both_from_dict(da); pb, ob = both_from_dict(db); pc, oc = both_from_dict(dc)
    pab, oab = pa.compose(pb), O.compose(oa, ob)
    assert_same(pab, oab, "a o b")
    unsorted_c = not (oc.props & R.props.I_LABEL_SORTED)
    if unsorted_c:
        with pytest.raises(ValueError):
            pab.compose(pc)
        with pytest.raises(O.OracleError):
            O.compose(oab, oc)
    pab.tr_sort(ilabel_cmp=False); oab.tr_sort(ilabel=False)
    assert_same(pab, oab, "tr_sort(olabel)")
    assert_same(pab.compose(pc), O.compose(oab, oc), "(a o b) o c")


@pytest.mark.parametrize("connect", [True, False])
def test_batched_compose_equals_individual_composes(connect):
    """BASELINE.json configs[4] shape: many acceptors against one shared transducer in ONE device BFS
    (b200_compose_batch); every result must equal the stand-alone composition bit-for-bit."""
    import rustfst_b200 as R
    from rustfst_b200 import synth
    t = synth.random_graph_transducer(3000, 30000, 40, seed=5)
    rng = np.random.default_rng(9)
    t["finals"] = np.where(rng.random(3000) < 0.6, rng.integers(0, 640, size=3000) / 64.0, np.inf).astype(np.float32)
    pt, ot = both_from_dict(t)
    accs = []
    for i in range(64):
        if i % 4 == 3:   # a label string that T does not necessarily accept
            labels = np.random.default_rng(100 + i).integers(1, 41, size=12)
        else:            # sampled from T: accepted at least up to a dead end
            labels = synth.sample_path_labels(t, 12 + (i % 5), seed=100 + i)
        accs.append(synth.linear_acceptor(labels, seed=300 + i))
    pairs = [both_from_dict(a) for a in accs]
    cfg = R.ComposeConfig(R.ComposeFilter.AUTOFILTER, connect)
    results, st = R.compose_batch([p for p, _ in pairs], pt, cfg)
    assert st["waves"] <= 20, "the batch must run as one BFS, not 64"
    nonempty = 0
    for i, ((p, o), r) in enumerate(zip(pairs, results)):
        expected = O.compose(o, ot, connect=connect)
        assert_same(r, expected, f"batch item {i} connect={connect}")
        nonempty += expected.num_states > 0
    assert nonempty > 10
    # the same batch as ONE packed block, transducer resident in HBM; the block survives a trip through bytes
    dt = R.DeviceFst.upload(pt)
    pb, st2 = R.compose_batch_packed([p for p, _ in pairs], config=cfg, device_transducer=dt)
    assert len(pb) == 64 and st2["arcs_out"] == st["arcs_out"]
    back = R.PackedBatch.from_buffer(pb.to_bytes())
    assert back.info() == pb.info()
    for i, r in enumerate(results):
        assert pb.result(i).to_bytes() == r.to_bytes() == back.result(i).to_bytes(), f"packed batch item {i}"


def test_batched_compose_heterogeneous_falls_back():
    import rustfst_b200 as R
    rng = np.random.default_rng(11)
    db = random_fst(rng, 10, 4, 4, sort="ilabel")
    pb, ob = both_from_dict(db)
    das = [random_fst(rng, 8, 3, 4, eps_prob=0.2, sort=("olabel" if i % 2 else None)) for i in range(6)]
    pairs = [both_from_dict(d) for d in das]
    results, _ = R.compose_batch([p for p, _ in pairs], pb)
    for (p, o), r in zip(pairs, results):
        assert_same(r, O.compose(o, ob), "heterogeneous batch")
    packed, _ = R.compose_batch_packed([p for p, _ in pairs], pb)
    for i, r in enumerate(results):
        assert packed.result(i).to_bytes() == r.to_bytes()


def test_shortest_path_with_near_ties_uses_the_order_faithful_parallel_path():
    """Continuous weights produce candidates within KDELTA of each other, so the certificate of the atomicMin path
    fails; the order-faithful parallel fold must then reproduce the reference's approximate relaxation exactly
    (shortest_path.rs:222-236), for both topological queue kinds."""
    import rustfst_b200 as R
    from rustfst_b200 import synth
    # (1) TOP_SORTED lattice -> StateOrderQueue
    g = synth.layered_acceptor(200_000, 2_000_000, 1000, 6, 40, continuous=True)
    p, o = both_from_dict(g)
    got, st = R.shortestpath_with_stats(p)
    assert st["queue_kind"] == 0 and st["path"] == 2, st
    assert_same(got, O.shortest_path(o), "continuous weights, state order")
    # (2) composed lattice: ACYCLIC known, TOP_SORTED unknown -> TopOrderQueue over the DFS order
    a = synth.layered_acceptor(20_000, 200_000, 32, 1, 20, continuous=True)
    b = synth.bigram_transducer(20_000, 200_000, 32, 2, 20, out_vocab=500, continuous=True)
    pa, oa = both_from_dict(a)
    pb, ob = both_from_dict(b)
    pc, oc = pa.compose(pb), O.compose(oa, ob)
    assert_same(pc, oc, "continuous weights compose")
    got, st = R.shortestpath_with_stats(pc)
    assert st["queue_kind"] == 1 and st["path"] in (0, 2), st
    assert_same(got, O.shortest_path(oc), f"continuous weights, top order (path={st['path']})")
    # the serial replay agrees as well
    got2, st2 = R.shortestpath_with_stats(pc, force_serial=True)
    assert st2["path"] == 1
    assert_same(got2, O.shortest_path(oc), "continuous weights, serial replay")


def test_dense_product_overflows_presized_buffers_and_grows():
    """Single-label complete machines: the product has (n1*n2) states and (n1*n2)^2-ish arcs, far beyond the
    persistent kernel's first capacity guess (4 x input arcs) -> it must stop cleanly, the call must double what
    overflowed and run again until the result fits, and the growing multi-kernel back end must agree (also exercises
    table rehash and buffer growth)."""
    import rustfst_b200 as R
    from rustfst_b200.fst import TR_DTYPE
    from rustfst_b200 import props as P

    def complete(n, seed):
        rng = np.random.default_rng(seed)
        arcs = np.zeros(n * n, dtype=TR_DTYPE)
        arcs["ilabel"] = 1; arcs["olabel"] = 1
        arcs["weight"] = rng.integers(0, 64, size=n * n) / 8.0
        arcs["nextstate"] = np.tile(np.arange(n), n)
        finals = np.full(n, np.inf, dtype=np.float32); finals[n - 1] = 0.5
        d = {"offsets": (np.arange(n + 1) * n).astype(np.uint32), "arcs": arcs, "finals": finals, "start": 0,
             "props": 0, "num_states": n}
        o = O.OFst.from_csr(d["offsets"].astype(np.uint64), arcs, finals, 0, 0)
        o.compute_props()
        d["props"] = o.props
        return d

    pa, oa = both_from_dict(complete(40, 1))
    pb, ob = both_from_dict(complete(45, 2))
    got, st = R.compose_with_stats(pa, pb)
    assert st["emit_launches"] == 1, "expected the persistent kernel (after capacity growth)"
    assert st["arcs_emitted"] == (40 * 45) * (40 * 45)
    expected = O.compose(oa, ob)
    assert_same(got, expected, "dense product")
    os.environ["B200_COMPOSE_IMPL"] = "waves"
    try:
        got2, st2 = R.compose_with_stats(pa, pb)
    finally:
        del os.environ["B200_COMPOSE_IMPL"]
    assert st2["emit_launches"] > 1
    assert_same(got2, expected, "dense product, multi-kernel back end")


def test_hub_state_with_long_label_runs():
    """Degree / run-length skew: one state with 20 000 arcs over 5 labels on each side (runs of ~4000 equal labels),
    so single items expand to thousands of arcs and single states to 20 001 items."""
    import rustfst_b200 as R
    from rustfst_b200.fst import TR_DTYPE
    rng = np.random.default_rng(3)

    def hub(n_leaf, deg, labels, by, seed):
        r = np.random.default_rng(seed)
        lab = np.sort(r.integers(1, labels + 1, size=deg))
        arcs = np.zeros(deg + n_leaf, dtype=TR_DTYPE)
        other = r.integers(1, labels + 1, size=deg)
        arcs["ilabel"][:deg] = lab if by == "ilabel" else other
        arcs["olabel"][:deg] = lab if by == "olabel" else other
        arcs["weight"][:deg] = r.integers(0, 64, size=deg) / 8.0
        arcs["nextstate"][:deg] = r.integers(1, n_leaf + 1, size=deg)
        # leaves loop back to the hub with one arc each
        arcs["ilabel"][deg:] = 1; arcs["olabel"][deg:] = 1; arcs["weight"][deg:] = 0.25; arcs["nextstate"][deg:] = 0
        offsets = np.concatenate([[0, deg], deg + 1 + np.arange(n_leaf)]).astype(np.uint32)
        finals = np.full(n_leaf + 1, np.inf, dtype=np.float32); finals[1] = 1.0
        d = {"offsets": offsets, "arcs": arcs, "finals": finals, "start": 0, "props": 0, "num_states": n_leaf + 1}
        o = O.OFst.from_csr(offsets.astype(np.uint64), arcs, finals, 0, 0)
        o.compute_props()
        d["props"] = o.props
        return d

    pa, oa = both_from_dict(hub(30, 20000, 5, "olabel", 10))
    pb, ob = both_from_dict(hub(30, 600, 5, "ilabel", 11))
    for connect in (False, True):
        got = R.compose_with_config(pa, pb, R.ComposeConfig(R.ComposeFilter.AUTOFILTER, connect))
        assert_same(got, O.compose(oa, ob, connect=connect), f"hub connect={connect}")


def test_tr_sort_on_device_is_stable():
    """fst_tr_sort of a large machine runs on the device (radix sort + gather) and must order arcs exactly like
    the reference's stable per-state sort (algorithms/tr_sort.rs:51-62)."""
    rng = np.random.default_rng(21)
    d = random_fst(rng, 20000, 12, 6, eps_prob=0.1)   # ~120k arcs: above the device threshold; many equal labels
    for ilabel in (True, False):
        p, o = both_from_dict(d)
        p.tr_sort(ilabel); o.tr_sort(ilabel)
        assert_same(p, o, f"device tr_sort ilabel={ilabel}")


# ------------------------------------------------------------------------------------------------ sigma matcher
def _acceptor(labels):
    """rustfst-python `acceptor()` shape: linear chain, weight one, last state final."""
    from rustfst_b200 import Tr, VectorFst
    f = VectorFst()
    states = [f.add_state() for _ in range(len(labels) + 1)]
    f.set_start(states[0])
    f.set_final(states[-1])
    for i, l in enumerate(labels):
        f.add_tr(states[i], Tr(l, l, None, states[i + 1]))
    return f


def test_sigma_compose_python_kats():
    """rustfst-python/tests/algorithms/test_compose.py:157-214 with the mirrored API."""
    from rustfst_b200 import ComposeConfig, ComposeFilter, MatcherConfig, MatcherRewriteMode, compose_with_config
    # symt: <eps> play david queen please <sigma>  ->  1 play, 3 queen, 4 please, 5 <sigma>
    query_fst, sigma_fst = _acceptor([1, 3, 4]), _acceptor([1, 5, 4])
    cfg = ComposeConfig(compose_filter=ComposeFilter.SEQUENCEFILTER, connect=True,
                        matcher2_config=MatcherConfig(sigma_label=5, rewrite_mode=MatcherRewriteMode.AUTO))
    assert compose_with_config(query_fst, sigma_fst, cfg) == query_fst
    # allowlist: symt <eps> play bowie queen radiohead please <sigma>
    sigma_fst = _acceptor([1, 6, 5])
    cfg = ComposeConfig(compose_filter=ComposeFilter.SEQUENCEFILTER, connect=True,
                        matcher2_config=MatcherConfig(sigma_label=6, rewrite_mode=MatcherRewriteMode.AUTO,
                                                      sigma_allowed_matches=[3, 2]))
    for artist, ok in ((3, True), (2, True), (4, False)):
        q = _acceptor([1, artist, 5])
        assert (compose_with_config(q, sigma_fst, cfg) == q) is ok
    # AutoFilter + custom matcher config is an error (compose_static.rs:219-223)
    bad = ComposeConfig(compose_filter=ComposeFilter.AUTOFILTER, connect=True,
                        matcher2_config=MatcherConfig(sigma_label=6))
    with pytest.raises(ValueError, match="AutoFilter"):
        compose_with_config(_acceptor([1, 2, 5]), sigma_fst, bad)


@pytest.mark.parametrize("seed", range(8))
def test_sigma_compose_fuzz(seed):
    """Random machines where label 9 plays sigma on fst2 (seed even) or on fst1 (seed odd), all rewrite modes,
    optional allow-lists, every non-auto filter — against the oracle's restatement of sigma_matcher.rs."""
    import rustfst_b200 as R
    rng = np.random.default_rng(4000 + seed)
    SIG = 9
    on_right = seed % 2 == 0

    def with_sigma(d, field):
        arcs = d["arcs"].copy()
        pick = rng.random(len(arcs)) < 0.25
        arcs[field][pick] = SIG
        if rng.random() < 0.5:   # acceptor-like sigma arcs exercise rewrite_both
            other = "olabel" if field == "ilabel" else "ilabel"
            arcs[other][pick] = SIG
        d = dict(d); d["arcs"] = arcs
        o = O.OFst.from_csr(d["offsets"].astype(np.uint64), arcs, d["finals"], d["start"], 0)
        o.tr_sort(ilabel=(field == "ilabel"))
        o.compute_props()
        off, a, f = o.to_csr()
        return {"offsets": off.astype(np.uint32), "arcs": a, "finals": f, "start": d["start"], "props": o.props,
                "num_states": d["num_states"]}

    da = random_fst(rng, int(rng.integers(2, 25)), 5, 6, eps_prob=0.15, sort="olabel")
    db = random_fst(rng, int(rng.integers(2, 25)), 5, 6, eps_prob=0.15, sort="ilabel")
    if on_right:
        db = with_sigma(db, "ilabel")
    else:
        da = with_sigma(da, "olabel")
    pa, oa = both_from_dict(da)
    pb, ob = both_from_dict(db)
    for mode in (0, 1, 2):
        for allowed in (None, [1, 3, 5]):
            spec = (SIG, mode, allowed)
            mcfg = R.MatcherConfig(SIG, R.MatcherRewriteMode(mode), allowed)
            for filt in (1, 2, 3, 4, 5, 6):
                for connect in (False, True):
                    kw = {"matcher2_config": mcfg} if on_right else {"matcher1_config": mcfg}
                    cfg = R.ComposeConfig(R.ComposeFilter(filt), connect, **kw)
                    try:
                        expected = O.compose_sigma(oa, ob, filt, connect, **({"sigma2": spec} if on_right else {"sigma1": spec}))
                    except O.OracleError as e:
                        with pytest.raises(ValueError):
                            R.compose_with_config(pa, pb, cfg)
                        continue
                    got = R.compose_with_config(pa, pb, cfg)
                    assert_same(got, expected, f"sigma fuzz seed={seed} mode={mode} allowed={allowed} filt={filt} c={connect}")


def test_connect_with_hub_states_in_both_directions():
    """fst_connect on a star: the start state has 10 000 out-arcs and the final state 10 000 in-arcs (both above the
    4096-neighbour threshold that routes a vertex to the 64-CTA hub kernel of the BFS levels), plus dead branches
    that must be trimmed; compared with the oracle's DFS-based connect (connect.rs:51-66)."""
    from rustfst_b200.fst import TR_DTYPE
    m = 10_000
    n = m + 2 + 500  # start, m middle states, final, 500 dead states
    rows, offsets = [], [0]
    for k in range(m):
        rows.append((1 + k % 13, 1 + k % 7, (k % 64) / 8.0, 1 + k))
    for k in range(250):  # dead ends reachable from the start (not coaccessible)
        rows.append((99, 99, 1.0, m + 2 + k))
    offsets.append(len(rows))
    for k in range(m):
        if k % 10 != 9:  # every tenth middle state is a dead end
            rows.append((2, 3, (k % 32) / 4.0, m + 1))
        offsets.append(len(rows))
    offsets.append(len(rows))  # final state: no arcs
    for k in range(500):  # dead states; the last 250 are unreachable but point at the final state
        if k >= 250:
            rows.append((5, 5, 0.5, m + 1))
        offsets.append(len(rows))
    arr = np.zeros(len(rows), dtype=TR_DTYPE)
    for i, r in enumerate(rows):
        arr[i] = r
    finals = np.full(n, np.inf, dtype=np.float32)
    finals[m + 1] = 0.25
    d = {"offsets": np.array(offsets, dtype=np.uint32), "arcs": arr, "finals": finals, "start": 0, "props": 0}
    o = O.OFst.from_csr(d["offsets"].astype(np.uint64), arr, finals, 0, 0)
    o.compute_props()
    d["props"] = o.props
    p, o = both_from_dict(d)
    p.connect()
    o.connect()
    assert p.num_states() == 2 + m - m // 10
    assert_same(p, o, "hub connect")


def test_concurrent_calls_on_shared_handles_from_several_host_threads():
    """SURVEY.md §8b threading contract: calls are synchronous and re-entrant, concurrent calls on shared-const handles
    are safe (every call has its own CUDA stream, the error slot and the n-best staging buffer are per thread).
    Eight host threads run compose / shortest_path / 5-best on the same operands at once; every result must equal the
    sequential one bit for bit."""
    import threading
    import rustfst_b200 as R
    from rustfst_b200 import synth
    a, b = synth.workload("C2", scale=0.05)
    pa, _ = both_from_dict(a)
    pb, _ = both_from_dict(b)
    ref_c = pa.compose(pb)
    ref_sp = ref_c.shortest_path()
    ref_nb = ref_c.shortest_path(R.ShortestPathConfig(nshortest=5))
    want = (ref_c.to_bytes(), ref_sp.to_bytes(), ref_nb.to_bytes())
    errors = []

    def worker(k):
        try:
            for _ in range(3):
                c = pa.compose(pb)
                sp = c.shortest_path()
                nb = ref_c.shortest_path(R.ShortestPathConfig(nshortest=5))  # shared input handle
                got = (c.to_bytes(), sp.to_bytes(), nb.to_bytes())
                if got != want:
                    errors.append(f"thread {k}: result differs")
        except Exception as e:  # noqa: BLE001
            errors.append(f"thread {k}: {e!r}")

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_fst_reverse_matches_the_reference_kat_and_the_oracle():
    """fst_reverse (rustfst-ffi/src/algorithms/reverse.rs:14-29) on the device: the reference's Python KAT
    (test_reverse.py:4-54), every fixture machine and random machines, bit-exact incl. the property word."""
    import rustfst_b200 as R
    f = R.VectorFst()
    s1, s2, s3 = f.add_state(), f.add_state(), f.add_state()
    f.set_start(s1)
    f.set_final(s3, 1.0)
    f.add_tr(s1, R.Tr(1, 2, 1.0, s2))
    f.add_tr(s1, R.Tr(3, 4, 2.0, s2))
    f.add_tr(s2, R.Tr(5, 6, 1.5, s2))
    f.add_tr(s2, R.Tr(3, 5, 1.0, s3))
    e = R.VectorFst()
    t1, t2, t3, t4 = e.add_state(), e.add_state(), e.add_state(), e.add_state()
    e.set_start(t1)
    e.set_final(t2)
    e.add_tr(t1, R.Tr(0, 0, 1.0, t4))
    e.add_tr(t4, R.Tr(3, 5, 1.0, t3))
    e.add_tr(t3, R.Tr(1, 2, 1.0, t2))
    e.add_tr(t3, R.Tr(3, 4, 2.0, t2))
    e.add_tr(t3, R.Tr(5, 6, 1.5, t3))
    assert f.reverse() == e
    for name in FIXTURES:
        for which in ("raw", "compose"):
            p, o = both_from_path(golden_path(name, which))
            assert_same(p.reverse(), o.reverse(), f"{name}/{which} reverse")
    rng = np.random.default_rng(4242)
    for k in range(8):
        d = random_fst(rng, int(rng.integers(1, 60)), 5, 6, eps_prob=0.2, cyclic=(k % 2 == 0), acceptor=(k % 3 == 0))
        p, o = both_from_dict(d)
        assert_same(p.reverse(), o.reverse(), f"random reverse {k}")
    # no start state, no states
    g = R.VectorFst(); og = O.OFst()
    assert_same(g.reverse(), og.reverse(), "empty reverse")
    g.add_state(); og.add_state()
    assert_same(g.reverse(), og.reverse(), "startless reverse")


def test_connect_kat_through_cabi():
    """rustfst-python/tests/algorithms/test_connect.py:4-52 through fst_connect on the device."""
    import rustfst_b200 as R
    f = R.VectorFst()
    for _ in range(5):
        f.add_state()
    f.set_start(0)
    f.set_final(1, 0.0)
    for (src, il, ol, w, dst) in [(4, 1, 2, 1.0, 0), (0, 3, 4, 2.0, 1), (1, 4, 5, 3.0, 2), (2, 4, 6, 4.0, 3), (2, 7, 8, 5.0, 0)]:
        f.add_tr(src, R.Tr(il, ol, w, dst))
    e = R.VectorFst()
    for _ in range(3):
        e.add_state()
    e.set_start(0)
    e.set_final(1, 0.0)
    for (src, il, ol, w, dst) in [(0, 3, 4, 2.0, 1), (1, 4, 5, 3.0, 2), (2, 7, 8, 5.0, 0)]:
        e.add_tr(src, R.Tr(il, ol, w, dst))
    res = f.connect()
    assert f == e and res == e


def test_spread_transducer_and_window_dag_workloads_at_reduced_scale():
    """The two extra bench workloads (bench.py run_extras) at a size the oracle finishes in seconds: a composition whose
    searched side is spread over the whole transducer with fan-out start states (a hub in the first wave), and the
    shortest path of a DAG with skip-level arcs."""
    import rustfst_b200 as R
    from rustfst_b200 import synth
    a1 = synth.layered_acceptor(40_000, 400_000, 93, 3, 50, start_fanout=True)
    a2 = synth.bigram_transducer(40_000, 400_000, 93, 4, 50, out_vocab=20000, start_fanout=True, spread=True)
    pa, oa = both_from_dict(a1)
    pb, ob = both_from_dict(a2)
    got, st = R.compose_with_stats(pa, pb)
    exp, ost = O.compose(oa, ob, want_stats=True)
    assert st["arcs_emitted"] == ost["arcs_emitted"] and st["arcs_emitted"] > 10_000
    assert_same(got, exp, "spread compose")
    g = synth.window_dag(100_000, 1_000_000, 1000, 6, window=1000)
    pg, og = both_from_dict(g)
    sp, sst = R.shortestpath_with_stats(pg)
    assert sst["queue_kind"] == 0
    assert_same(sp, O.shortest_path(og), "window dag shortest path")
