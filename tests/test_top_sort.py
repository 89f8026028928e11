"""fst_top_sort (rustfst-ffi/src/algorithms/top_sort.rs:13-21): host-side in this library (the numbering is the finish
order of the reference's sequential DFS), so the whole check runs without a GPU."""
import ctypes as C

import numpy as np
import pytest

import rustfst_b200 as R
from rustfst_b200 import props as P
from rustfst_b200.ffi import lib
from tests import oracle_lib as O
from tests.parity_utils import assert_same, both_from_dict, random_fst


def test_top_sort_kat_from_the_reference_python_tests():
    """rustfst-python/tests/algorithms/test_top_sort.py:4-20"""
    for cls, mk in ((R.VectorFst, lambda il, ol, w, d: (R.Tr(il, ol, w, d),)), (O.OFst, lambda il, ol, w, d: (il, ol, w, d))):
        f = cls()
        s1, s2 = f.add_state(), f.add_state()
        f.set_start(s2)
        f.set_final(s1, 0.0)
        f.add_tr(s2, *mk(1, 2, 1.0, s1))
        start_before = f.start() if callable(getattr(f, "start")) else f.start
        assert start_before == s2
        f.top_sort()
        start_after = f.start() if callable(getattr(f, "start")) else f.start
        assert start_after == s1


@pytest.mark.parametrize("seed", range(30))
def test_top_sort_matches_the_oracle(seed):
    rng = np.random.default_rng(3000 + seed)
    cyclic = seed % 5 == 0
    d = random_fst(rng, int(rng.integers(1, 40)), 4, 5, eps_prob=0.1, cyclic=cyclic)
    if seed % 3 == 0 and d["num_states"] > 2:
        d["start"] = int(rng.integers(0, d["num_states"]))  # unreachable states become extra DFS roots
    p, o = both_from_dict(d)
    p.top_sort()
    o.top_sort()
    assert_same(p, o, f"top_sort seed={seed}")
    if p.properties & P.TOP_SORTED:
        off, arcs, fin, start = p.to_csr()
        src = np.repeat(np.arange(len(fin)), np.diff(off.astype(np.int64)))
        assert np.all(arcs["nextstate"] > src), "all transitions go from lower to higher state ids"
        kind = C.c_int32()
        R.check_ffi_error(lib.b200_shortest_path_queue_plan(p.ptr, C.byref(kind), None, None, None), "plan")
        assert kind.value == 0, "a top-sorted machine takes the StateOrderQueue route (no DFS per shortest-path call)"


def test_top_sort_without_a_start_state_reports_the_reference_error():
    f = R.VectorFst()
    f.add_state(); f.add_state()
    with pytest.raises(ValueError, match=r"StateSort : Bad order vector size : 0\. Expected 2"):
        f.top_sort()
    R.VectorFst().top_sort()  # empty machine: nothing to do
