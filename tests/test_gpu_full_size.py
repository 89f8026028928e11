"""Parity at BASELINE.json's FULL sizes: C2 and C4 against the oracle directly (it still finishes in seconds), C3
through size-independent properties (determinism, agreement of the two device back ends, trim idempotence,
compose(connect=false) + fst_connect == compose(connect=true), CSR well-formedness) plus one oracle comparison."""
import os

import numpy as np
import pytest

from tests import oracle_lib as O
from tests.parity_utils import assert_same, both_from_dict

pytestmark = pytest.mark.gpu


def test_c2_full_size_compose_and_shortest_path_match_oracle():
    import rustfst_b200 as R
    from rustfst_b200 import synth
    a, b = synth.workload("C2")
    pa, oa = both_from_dict(a)
    pb, ob = both_from_dict(b)
    o, ost = O.compose(oa, ob, want_stats=True)
    p, st = R.compose_with_stats(pa, pb)
    assert st["arcs_emitted"] == ost["arcs_emitted"] and st["states_expanded"] == ost["states_expanded"]
    assert_same(p, o, "C2 full size")
    sp, sst = R.shortestpath_with_stats(p)
    assert sst["path"] == 0
    assert_same(sp, O.shortest_path(o), "C2 full size lattice shortest path")


def test_c4_full_size_shortest_path_matches_oracle():
    import rustfst_b200 as R
    from rustfst_b200 import synth
    g = synth.workload("C4")
    p, o = both_from_dict(g)
    got, st = R.shortestpath_with_stats(p)
    assert st["path"] == 0 and st["queue_kind"] == 0
    expected, ost = O.shortest_path(o, want_stats=True)
    assert st["arcs_relaxed"] == ost["arcs_relaxed"]
    assert_same(got, expected, "C4 full size shortest path")


def test_c3_full_size_properties_and_oracle():
    import rustfst_b200 as R
    from rustfst_b200 import synth
    a, b = synth.workload("C3")
    pa, oa = both_from_dict(a)
    pb, ob = both_from_dict(b)
    r1, st1 = R.compose_with_stats(pa, pb)
    # well-formed CSR
    off, arcs, fin, start = r1.to_csr()
    n = len(fin)
    assert start == 0 and off[0] == 0 and off[-1] == len(arcs) and np.all(np.diff(off.astype(np.int64)) >= 0)
    assert n == st1["states_out"] and len(arcs) == st1["arcs_out"]
    assert int(arcs["nextstate"].max()) < n
    assert np.isfinite(arcs["weight"]).all()
    # deterministic: same bytes on a second run
    r2, _ = R.compose_with_stats(pa, pb)
    assert r1.to_bytes() == r2.to_bytes()
    # the multi-kernel back end produces the identical FST
    os.environ["B200_COMPOSE_IMPL"] = "waves"
    try:
        r3, st3 = R.compose_with_stats(pa, pb)
    finally:
        del os.environ["B200_COMPOSE_IMPL"]
    assert st3["emit_launches"] > 1 and st1["emit_launches"] == 1
    assert r1.to_bytes() == r3.to_bytes()
    # connect is idempotent; compose(connect=false) + fst_connect == compose(connect=true)
    untrimmed = R.compose_with_config(pa, pb, R.ComposeConfig(R.ComposeFilter.AUTOFILTER, False))
    assert untrimmed.num_states() == st1["states_expanded"] and untrimmed.num_trs_total() == st1["arcs_emitted"]
    untrimmed.connect()
    assert untrimmed.to_bytes() == r1.to_bytes()
    again = r1.copy().connect()
    assert again.to_bytes() == r1.to_bytes()
    # and the oracle agrees (about 10 s of CPU)
    assert_same(r1, O.compose(oa, ob), "C3 full size")


def test_c5_full_shape_batch_matches_oracle_on_a_256_acceptor_sample():
    """BASELINE.json configs[4] at its full shape: 8192 linear acceptors (200 arcs) against one 500K-state / 5M-arc
    transducer in ONE batched call (packed result, transducer resident in HBM); 256 of the results, spread over the
    batch, are compared bit-for-bit with the oracle composing that acceptor alone."""
    import rustfst_b200 as R
    from rustfst_b200 import synth
    n_t = 500_000
    t = synth.random_graph_transducer(n_t, 5_000_000, 5000, seed=5)
    rng = np.random.default_rng(77)
    t["finals"] = np.where(rng.random(n_t) < 0.5, rng.integers(0, 640, size=n_t) / 64.0, np.inf).astype(np.float32)
    pt, ot = both_from_dict(t)
    dt = R.DeviceFst.upload(pt)
    labels = synth.sample_path_labels_batch(t, 200, 8192, seed=100)
    dicts = [synth.linear_acceptor(labels[i], seed=100 + i) for i in range(8192)]
    accs = [synth.to_vector_fst(d) for d in dicts]
    pb, st = R.compose_batch_packed(accs, device_transducer=dt)
    assert len(pb) == 8192 and st["waves"] == 201 and st["emit_launches"] == 1
    back = R.PackedBatch.from_buffer(pb.to_numpy())
    assert back.info() == pb.info()
    nonempty = 0
    for i in range(0, 8192, 32):
        d = dicts[i]
        oa = O.OFst.from_csr(d["offsets"].astype(np.uint64), d["arcs"], d["finals"], d["start"], d["props"])
        expected = O.compose(oa, ot)
        assert_same(pb.result(i), expected, f"C5 batch item {i}")
        nonempty += expected.num_states > 0
    assert nonempty >= 128
