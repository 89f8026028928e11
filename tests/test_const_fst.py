"""ConstFst handles and the OpenFst binary "const" format: const_fst_from_path / write_file /
equals / copy / display (rustfst-ffi/src/fst/const_fst.rs), host-only, no GPU involved."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

import rustfst_b200 as R
from rustfst_b200.ffi import lib
from tests import oracle_lib as O
from tests.parity_utils import golden_path

CONST_FIXTURE = os.path.join(os.path.dirname(__file__), "golden", "fst_012_hcl_const.fst.in")


def test_openfst_written_aligned_const_file_equals_its_vector_twin():
    """rustfst-tests-data/fst_012/hcl.fst.in is a const file written by OpenFst (version 1: 16-byte aligned sections);
    tests/golden/fst_012_raw.fst is the same machine in the vector format."""
    c = R.ConstFst.read(CONST_FIXTURE)
    v = R.VectorFst.read(golden_path("fst_012", "raw"))
    assert (c.num_states(), c.num_trs_total()) == (215, 942)
    assert c.properties == 0x65a5656a0000  # header word 0x65a5656a0001 without the binary EXPANDED bit
    co, ca, cf, cs = c.to_csr()
    vo, va, vf, vs = v.to_csr()
    assert np.array_equal(co, vo) and ca.tobytes() == va.tobytes() and cf.tobytes() == vf.tobytes() and cs == vs
    assert str(c) == str(v)
    assert c.start() == 0 and c.num_trs(0) == v.num_trs(0)
    assert [(t.ilabel, t.olabel, t.next_state) for t in c.trs(3)] == [(t.ilabel, t.olabel, t.next_state) for t in v.trs(3)]


def test_const_write_read_round_trip_and_oracle_cross_check(tmp_path):
    c = R.ConstFst.read(CONST_FIXTURE)
    out = tmp_path / "packed.fst"
    c.write(out)
    raw = out.read_bytes()
    # header: magic, "const", "standard", version 2 (packed), flags 0, properties | EXPANDED
    assert struct.unpack_from("<i", raw, 0)[0] == 2125659606
    assert raw[4:4 + 4 + 5] == struct.pack("<i", 5) + b"const"
    assert struct.unpack_from("<i", raw, 4 + 9 + 12)[0] == 2
    assert struct.unpack_from("<Q", raw, 4 + 9 + 12 + 8)[0] == (0x65a5656a0000 | 1)
    c2 = R.ConstFst.read(out)
    assert c2 == c and c.copy() == c and str(c2) == str(c)
    # the oracle's reader (const_fst/serializable_fst.rs restated independently) accepts what we wrote
    o = O.OFst.from_bytes(raw)
    assert o == O.OFst.from_path(golden_path("fst_012", "raw"))


def _const_bytes(version, start, states, arcs):
    """Hand-written const file: states = [(final or inf, pos, ntrs, nieps, noeps)], arcs = [(il, ol, w, ns)]."""
    b = struct.pack("<i", 2125659606) + struct.pack("<i", 5) + b"const" + struct.pack("<i", 8) + b"standard"
    b += struct.pack("<iIQqqq", version, 0, 0x1, start, len(states), len(arcs))
    if version == 1 and states and len(b) % 16:
        b += b"\0" * (16 - len(b) % 16)
    for fw, pos, ntrs, nie, noe in states:
        b += struct.pack("<fiiii", fw, pos, ntrs, nie, noe)
    if version == 1 and arcs and len(b) % 16:
        b += b"\0" * (16 - len(b) % 16)
    for il, ol, w, ns in arcs:
        b += struct.pack("<iifi", il, ol, w, ns)
    return b


def test_aligned_and_packed_layouts_parse_to_the_same_machine(tmp_path):
    states = [(float("inf"), 0, 2, 0, 1), (0.5, 2, 1, 1, 0), (float("inf"), 3, 0, 0, 0)]
    arcs = [(1, 0, 1.5, 1), (2, 3, 0.25, 2), (0, 4, 2.0, 1)]
    paths = []
    for version in (1, 2):
        p = tmp_path / f"v{version}.fst"
        p.write_bytes(_const_bytes(version, 0, states, arcs))
        paths.append(p)
    a, b = R.ConstFst.read(paths[0]), R.ConstFst.read(paths[1])
    assert a == b and a.num_states() == 3 and a.num_trs_total() == 3
    assert a.final(1) == 0.5 and a.final(0) is None and a.start() == 0
    assert [(t.ilabel, t.olabel, t.weight, t.next_state) for t in a.trs(0)] == [(1, 0, 1.5, 1), (2, 3, 0.25, 2)]
    # truncated file / wrong type -> the reference's single error message
    bad = tmp_path / "bad.fst"
    bad.write_bytes(_const_bytes(2, 0, states, arcs)[:-7])
    with pytest.raises(ValueError, match="Error while parsing binary ConstFst"):
        R.ConstFst.read(bad)
    with pytest.raises(ValueError, match="Error while parsing binary ConstFst"):
        R.ConstFst.read(golden_path("fst_012", "raw"))  # a vector file


def test_handle_kinds_are_not_interchangeable():
    """as_fst! downcasts (rustfst-ffi/src/fst/mod.rs:99-111; algorithms/compose.rs:315-321)."""
    c = R.ConstFst.read(CONST_FIXTURE)
    v = R.VectorFst.read(golden_path("fst_012", "raw"))
    n = C.c_size_t()
    with pytest.raises(ValueError, match=r"Could not downcast to VectorFst<TropicalWeight> FST"):
        R.check_ffi_error(lib.vec_fst_num_states(c.ptr, C.byref(n)), "num_states")
    s = C.c_char_p()
    with pytest.raises(ValueError, match=r"Could not downcast to ConstFst<TropicalWeight> FST"):
        R.check_ffi_error(lib.const_fst_display(v.ptr, C.byref(s)), "display")
    out = C.c_void_p()
    for call in (lambda: lib.fst_compose(c.ptr, v.ptr, C.byref(out)), lambda: lib.fst_shortest_path(c.ptr, C.byref(out)),
                 lambda: lib.fst_connect(c.ptr), lambda: lib.fst_reverse(c.ptr, C.byref(out)),
                 lambda: lib.fst_tr_sort(c.ptr, True)):
        with pytest.raises(ValueError, match="Could not downcast to vector FST"):
            R.check_ffi_error(call(), "algorithm on a const handle")


def _props_both(d):
    """Property word computed by the product from the content (all bits unknown beforehand) and by the oracle."""
    p = R.VectorFst.from_csr(d["offsets"], d["arcs"], d["finals"], d["start"], 0)
    o = O.OFst.from_csr(d["offsets"].astype(np.uint64), d["arcs"], d["finals"], d["start"], 0)
    o.compute_props()
    return p.compute_properties(), o.props


@pytest.mark.parametrize("seed", range(40))
def test_compute_properties_matches_the_oracle_on_random_machines(seed):
    """b200_fst_compute_properties = compute_and_update_properties_all (compute_fst_properties.rs:14-208): cyclic /
    acyclic machines, epsilons, acceptors, unreachable and dead states, weighted cycles."""
    from tests.parity_utils import random_fst
    rng = np.random.default_rng(9000 + seed)
    d = random_fst(rng, int(rng.integers(1, 30)), int(rng.integers(1, 5)), int(rng.integers(1, 6)), eps_prob=0.2,
                   acceptor=(seed % 3 == 0), sort=[None, "ilabel", "olabel"][seed % 3], cyclic=(seed % 2 == 0),
                   weight_grid=True, final_prob=0.3)
    got, want = _props_both(d)
    assert got == want, f"{got:#x} != {want:#x}"


def test_compute_properties_on_fixtures_and_degenerate_machines():
    from tests.parity_utils import FIXTURES
    for name in FIXTURES:
        for which in ("raw", "compose"):
            v = R.VectorFst.read(golden_path(name, which))
            off, arcs, fin, start = v.to_csr()
            got, want = _props_both({"offsets": off, "arcs": arcs, "finals": fin, "start": start})
            assert got == want, f"{name}/{which}: {got:#x} != {want:#x}"
            assert got == v.properties, "the fixture files carry fully computed words"
    # no start state (the DFS group keeps the visitor's initial word), empty machine, a string, a self loop on the start
    from rustfst_b200.fst import TR_DTYPE
    one = np.zeros(1, dtype=TR_DTYPE); one[0] = (1, 1, 2.5, 0)
    cases = [
        {"offsets": np.array([0, 1], np.uint32), "arcs": one, "finals": np.array([np.inf], np.float32), "start": None},
        {"offsets": np.array([0], np.uint32), "arcs": np.zeros(0, TR_DTYPE), "finals": np.zeros(0, np.float32), "start": None},
        {"offsets": np.array([0, 1], np.uint32), "arcs": one, "finals": np.array([0.0], np.float32), "start": 0},
    ]
    chain = np.zeros(2, dtype=TR_DTYPE); chain[0] = (1, 1, 0.0, 1); chain[1] = (2, 2, 0.0, 2)
    cases.append({"offsets": np.array([0, 1, 2, 2], np.uint32), "arcs": chain,
                  "finals": np.array([np.inf, np.inf, 0.0], np.float32), "start": 0})
    for d in cases:
        got, want = _props_both(d)
        assert got == want, f"{got:#x} != {want:#x}"


def test_const_from_vector_fst_freezes_all_properties(tmp_path):
    """const_fst_from_vec_fst (const_fst.rs:157-170, converters.rs:7-37): a copy with every property computed; the
    source handle keeps its own word."""
    v = R.VectorFst()
    s0, s1 = v.add_state(), v.add_state()
    v.set_start(s0)
    v.set_final(s1, 1.5)
    v.add_tr(s0, R.Tr(2, 2, 0.5, s1))
    v.add_tr(s0, R.Tr(1, 1, 0.25, s1))
    before = v.properties
    c = R.ConstFst.from_vector_fst(v)
    assert v.properties == before
    o = O.OFst()
    o.add_state(); o.add_state(); o.set_start(0); o.set_final(1, 1.5)
    o.add_tr(0, 2, 2, 0.5, 1); o.add_tr(0, 1, 1, 0.25, 1)
    o.compute_props()
    assert c.properties == o.props
    assert str(c) == str(v) and c.num_states() == 2
    out = tmp_path / "c.fst"
    c.write(out)
    assert R.ConstFst.read(out) == c
