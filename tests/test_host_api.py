"""CPU tests of the host side of librustfst_b200 (no compute calls): the container mirrors the reference's
observable behaviour — property word maintained on every mutation, optional out-slots, error convention,
OpenFst binary I/O — checked against the oracle's independent restatement and the reference's binary fixtures."""
import ctypes as C
import os

import numpy as np
import pytest

import rustfst_b200 as R
from rustfst_b200.ffi import lib
from tests import oracle_lib as O
from tests.parity_utils import FIXTURES, GOLDEN, assert_same, golden_path, random_fst


def test_library_reports_version():
    assert b"sm_100a" in lib.b200_version()


def test_incremental_build_tracks_properties_like_the_reference():
    rng = np.random.default_rng(0)
    for trial in range(40):
        n = int(rng.integers(1, 8))
        p, o = R.VectorFst(), O.OFst()
        for _ in range(n):
            assert p.add_state() == o.add_state()
        ops = int(rng.integers(0, 25))
        for _ in range(ops):
            kind = rng.integers(0, 3)
            s = int(rng.integers(0, n))
            if kind == 0:
                il, ol = int(rng.integers(0, 4)), int(rng.integers(0, 4))
                w = float(rng.integers(0, 4)) / 2
                ns = int(rng.integers(0, n))
                p.add_tr(s, R.Tr(il, ol, w, ns)); o.add_tr(s, il, ol, w, ns)
            elif kind == 1:
                w = float(rng.integers(0, 3))
                p.set_final(s, w); o.set_final(s, w)
            else:
                p.set_start(s); o.set_start(s)
            assert p.properties == o.props, f"trial {trial}"
        assert_same(p, o, f"trial {trial}")


def test_null_properties_and_mutators():
    f = R.VectorFst()
    assert f.properties == 0x0000956A5A950000  # properties.rs null_properties()
    s = f.add_state()
    assert f.start() is None and f.final(s) is None and f.num_trs(s) == 0
    f.set_final(s, 1.5)
    assert f.final(s) == 1.5 and f.is_final(s)
    f.unset_final(s)
    assert f.final(s) is None
    with pytest.raises(ValueError, match="doesn't exist"):
        f.set_start(3)
    with pytest.raises(ValueError, match="doesn't exist"):
        f.add_tr(7, R.Tr(1, 1, 0.0, 0))
    f.delete_states()
    assert f.num_states() == 0 and f.properties == 0x0000956A5A950000


def test_optional_out_slots_are_left_untouched():
    f = R.VectorFst()
    f.add_state()
    slot = C.c_uint32(12345)
    assert lib.fst_start(f.ptr, C.byref(slot)) == 0 and slot.value == 12345
    w = C.c_float(7.25)
    assert lib.fst_final_weight(f.ptr, 0, C.byref(w)) == 0 and w.value == 7.25
    sym = C.c_void_p(0xDEAD)
    assert lib.fst_input_symbols(f.ptr, C.byref(sym)) == 0 and sym.value == 0xDEAD


def test_error_convention():
    f = R.VectorFst()
    assert lib.vec_fst_set_start(f.ptr, 3) == 1
    msg = C.c_char_p()
    assert lib.rustfst_ffi_get_last_error(C.byref(msg)) == 0
    assert b"doesn't exist" in C.string_at(msg)
    lib.rustfst_destroy_string(msg)
    msg = C.c_char_p()
    assert lib.rustfst_ffi_get_last_error(C.byref(msg)) == 0  # taken: the slot is now empty
    assert C.string_at(msg) == b"No error message"
    lib.rustfst_destroy_string(msg)
    assert lib.fst_destroy(None) == 0 and lib.tr_delete(None) == 0


@pytest.mark.parametrize("name", FIXTURES)
def test_binary_io_matches_oracle_and_roundtrips(name):
    for which in ("raw", "compose"):
        path = golden_path(name, which)
        p, o = R.VectorFst.read(path), O.OFst.from_path(path)
        assert_same(p, o, f"{name}/{which}")
        raw = open(path, "rb").read()
        assert p.to_bytes() == raw == o.to_bytes()
        q = R.VectorFst.from_bytes(raw)
        assert q == p and q.properties == p.properties


def test_reads_reference_vector_files_with_symbol_free_headers(tmp_path):
    # rustfst-tests-data/fst_020/patterns_fst.fst.in re-serialised by the oracle lives in tests/golden
    f = R.VectorFst.read(golden_path("fst_020", "raw"))
    assert f.num_states() == 66 and f.num_trs_total() == 83
    out = tmp_path / "x.fst"
    f.write(out)
    assert R.VectorFst.read(out) == f
    with pytest.raises(ValueError):
        R.VectorFst.from_bytes(b"not an fst")
    with pytest.raises(ValueError):
        R.VectorFst.read(tmp_path / "missing.fst")


def test_tr_sort_is_stable_and_sets_bits():
    rng = np.random.default_rng(3)
    d = random_fst(rng, 20, 8, 4, eps_prob=0.2)
    for ilabel in (True, False):
        p = R.VectorFst.from_csr(d["offsets"], d["arcs"], d["finals"], d["start"], d["props"])
        o = O.OFst.from_csr(d["offsets"].astype(np.uint64), d["arcs"], d["finals"], d["start"], d["props"])
        p.tr_sort(ilabel); o.tr_sort(ilabel)
        assert_same(p, o, f"tr_sort ilabel={ilabel}")


def test_equals_is_approximate_on_weights_and_ignores_properties():
    a, b = R.VectorFst(), R.VectorFst()
    for f, w in ((a, 1.0), (b, 1.0 + 1.0 / 2048)):
        f.add_state(); f.add_state(); f.set_start(0); f.set_final(1, w); f.add_tr(0, R.Tr(1, 2, w, 1))
    assert a == b  # |dw| <= KDELTA = 1/1024 (semiring.rs:159-168)
    b.properties = 0
    assert a == b
    c = a.copy()
    c.add_tr(1, R.Tr(1, 1, 0.0, 0))
    assert a != c


def test_iterators_and_trs():
    f = R.VectorFst()
    f.add_state(); f.add_state(); f.set_start(0)
    f.add_tr(0, R.Tr(3, 4, 0.5, 1)); f.add_tr(0, R.Tr(5, 6, 1.5, 0))
    got = [(t.ilabel, t.olabel, t.weight, t.next_state) for t in f.trs(0)]
    assert got == [(3, 4, 0.5, 1), (5, 6, 1.5, 0)]
    assert list(f.states()) == [0, 1]
    t = R.Tr(1, 2, 3.0, 4)
    t.ilabel = 9; t.weight = 0.25; t.next_state = 2
    assert (t.ilabel, t.olabel, t.weight, t.next_state) == (9, 2, 0.25, 2)
    assert str(f) == "0\t1\t3\t4\t0.5\n0\t0\t5\t6\t1.5\n"


def test_compute_entry_points_fail_loudly_without_a_gpu():
    if R.device_count() > 0:
        pytest.skip("a GPU is present")
    a = R.VectorFst.read(golden_path("fst_003", "raw"))
    b = R.VectorFst.read(golden_path("fst_003", "compose"))
    with pytest.raises(ValueError, match="no CUDA device"):
        a.compose(b)
    with pytest.raises(ValueError, match="no CUDA device"):
        a.shortest_path()
    with pytest.raises(ValueError, match="no CUDA device"):
        a.connect()


def _product_queue_plan(p):
    n = p.num_states()
    kind, n_scc = C.c_int32(), C.c_uint32()
    order = np.zeros(max(1, n), dtype=np.uint32)
    fifo = np.zeros(max(1, n), dtype=np.uint8)
    assert lib.b200_shortest_path_queue_plan(p.ptr, C.byref(kind), order.ctypes.data, fifo.ctypes.data,
                                             C.byref(n_scc)) == 0
    return kind.value, order[:n], fifo[:n_scc.value]


@pytest.mark.parametrize("name", FIXTURES)
def test_queue_plan_matches_auto_queue_on_fixtures(name):
    """AutoQueue selection + DFS orders (auto_queue.rs:23-99, top_sort.rs, scc_visitors.rs) are host logic: the
    product's plan must equal the oracle's restatement on every fixture machine (cyclic, epsilon-rich)."""
    for which in ("raw", "compose"):
        p, o = R.VectorFst.read(golden_path(name, which)), O.OFst.from_path(golden_path(name, which))
        pk, po, pf = _product_queue_plan(p)
        ok, oo, of = O.queue_plan(o)
        assert pk == ok, f"{name}/{which}: kind {pk} != {ok}"
        if ok in (1, 3):
            assert np.array_equal(po, oo), f"{name}/{which}: order/scc differ"
        if ok == 3:
            assert np.array_equal(pf, of)


@pytest.mark.parametrize("seed", range(12))
def test_queue_plan_fuzz(seed):
    rng = np.random.default_rng(5000 + seed)
    d = random_fst(rng, int(rng.integers(1, 80)), 5, 6, eps_prob=0.1, cyclic=(seed % 3 != 0),
                   weight_grid=(seed % 4 != 1))
    if seed % 5 == 0:   # unknown properties force the SCC branch
        d["props"] = 0
    p = R.VectorFst.from_csr(d["offsets"], d["arcs"], d["finals"], d["start"], d["props"])
    o = O.OFst.from_csr(d["offsets"].astype(np.uint64), d["arcs"], d["finals"], d["start"], d["props"])
    pk, po, pf = _product_queue_plan(p)
    ok, oo, of = O.queue_plan(o)
    assert pk == ok
    if ok in (1, 3):
        assert np.array_equal(po, oo)
    if ok == 3:
        assert np.array_equal(pf, of)


def test_large_machine_io_round_trip_uses_the_threaded_copy_path(tmp_path):
    """Above 65 536 states parse / store copy on several host threads; bytes must equal the oracle's serialisation and
    the machine must survive file and in-memory round trips."""
    import numpy as np
    import rustfst_b200 as R
    from rustfst_b200 import synth
    from tests import oracle_lib as O
    g = synth.layered_acceptor(150_000, 900_000, 50, 11, 30)
    v = synth.to_vector_fst(g)
    o = O.OFst.from_csr(g["offsets"].astype(np.uint64), g["arcs"], g["finals"], g["start"], g["props"])
    b = v.to_bytes()
    assert b == o.to_bytes()
    w = R.VectorFst.from_bytes(b)
    assert w == v and w.properties == v.properties
    path = tmp_path / "big.fst"
    v.write(path)
    assert path.read_bytes() == b
    r = R.VectorFst.read(path)
    ro, ra, rf, rs = r.to_csr()
    assert np.array_equal(ro, g["offsets"]) and ra.tobytes() == g["arcs"].tobytes() and rf.tobytes() == g["finals"].tobytes()
    assert rs == g["start"]
