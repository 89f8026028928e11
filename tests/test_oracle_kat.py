"""Pins the CPU oracle against every known-answer test the reference holds in-tree for this path.

Sources (all under /root/reference):
  * rustfst-python/tests/algorithms/test_compose.py:13-81   default compose (AutoFilter, connect)
  * rustfst-python/tests/algorithms/test_compose.py:84-154  TRIVIALFILTER compose config
  * rustfst-python/tests/algorithms/test_shortest_path.py:5-51
  * rustfst/src/algorithms/compose/compose_static.rs:313-321 doc-test  fst![1,2=>2,3] o fst![2,3=>3,4]
  * SURVEY.md App. A hand-simulated sizes for the fixtures (second independent reading)
The OpenFst-generated goldens for fst_000..020 are git-ignored upstream and absent here, so these KATs are the
only reference-produced answers available ("parity pinned by in-tree KATs").
"""
import os

import pytest

from tests import oracle_lib as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def build(n, start, finals, arcs):
    f = O.OFst()
    for _ in range(n):
        f.add_state()
    if start is not None:
        f.set_start(start)
    for s, w in finals:
        f.set_final(s, w)
    for (s, il, ol, w, ns) in arcs:
        f.add_tr(s, il, ol, w, ns)
    return f


def kat_fst1():
    return build(3, 0, [(1, 0.0), (2, 0.0)], [(0, 1, 2, 1.0, 1), (0, 1, 4, 2.0, 2), (1, 3, 5, 2.0, 1)])


def kat_fst2():
    return build(3, 0, [(2, 0.0)], [(0, 2, 6, 1.0, 1), (1, 5, 7, 2.5, 2), (2, 5, 8, 1.5, 2), (0, 4, 9, 3.0, 2)])


def kat_expected():
    return build(4, 0, [(2, 0.0), (3, 0.0)],
                 [(0, 1, 6, 2.0, 1), (0, 1, 9, 5.0, 2), (1, 3, 7, 4.5, 3), (3, 3, 8, 3.5, 3)])


def test_compose_default_kat():
    assert O.compose(kat_fst1(), kat_fst2()) == kat_expected()


def test_compose_trivial_filter_kat():
    assert O.compose(kat_fst1(), kat_fst2(), filter=2, connect=True) == kat_expected()


def test_compose_doc_test_kat():
    # utils::transducer(&[1,2],&[2,3]) : linear, final = last state, weight one
    a = build(3, 0, [(2, 0.0)], [(0, 1, 2, 0.0, 1), (1, 2, 3, 0.0, 2)])
    b = build(3, 0, [(2, 0.0)], [(0, 2, 3, 0.0, 1), (1, 3, 4, 0.0, 2)])
    e = build(3, 0, [(2, 0.0)], [(0, 1, 3, 0.0, 1), (1, 2, 4, 0.0, 2)])
    assert O.compose(a, b) == e


def test_shortest_path_kat():
    f = build(4, 0, [(3, 2.0)], [(0, 1, 1, 3.0, 1), (1, 2, 2, 2.0, 1), (1, 3, 3, 4.0, 3), (0, 4, 4, 5.0, 2),
                                 (2, 5, 5, 4.0, 3)])
    e = build(3, 2, [(0, 2.0)], [(2, 1, 1, 3.0, 1), (1, 3, 3, 4.0, 0)])
    assert O.shortest_path(f, nshortest=1, unique=True) == e


def test_unsorted_errors():
    a = build(2, 0, [(1, 0.0)], [(0, 1, 5, 0.0, 1), (0, 1, 2, 0.0, 1)])  # olabels 5,2: not sorted
    b = build(2, 0, [(1, 0.0)], [(0, 7, 1, 0.0, 1), (0, 2, 1, 0.0, 1)])  # ilabels 7,2: not sorted
    with pytest.raises(O.OracleError, match="sort"):
        O.compose(a, b)


def fixture(name, which):
    return O.OFst.from_path(os.path.join(GOLDEN, f"{name}_{which}.fst"))


@pytest.mark.parametrize("name,pre,post", [
    ("fst_012", (35, 65), (5, 5)),
    ("fst_013", (17, 56), (17, 56)),
    ("fst_014", None, (5, 5)),
    ("fst_003", (4, 3), (4, 3)),
])
def test_fixture_sizes_match_survey_simulation(name, pre, post):
    a, b = fixture(name, "raw"), fixture(name, "compose")
    if pre is not None:
        r = O.compose(a, b, filter=0, connect=False)
        assert (r.num_states, r.num_trs) == pre
    r = O.compose(a, b, filter=0, connect=True)
    assert (r.num_states, r.num_trs) == post


def test_fst_003_pair_values():
    r = O.compose(fixture("fst_003", "raw"), fixture("fst_003", "compose"))
    e = build(4, 0, [(3, 0.7 + 1.2)], [(0, 12, 2, 0.3 + 1.7, 1), (1, 0, 1, 1.2, 2), (2, 5, 5, 0.1 + 1.8, 3)])
    assert r == e


def test_fst_003_cross_004_is_empty():
    r = O.compose(fixture("fst_003", "raw"), fixture("fst_004", "raw"))
    assert r.num_states == 0 and r.start is None


def test_bytes_roundtrip():
    for name in ("fst_012", "fst_020", "fst_000"):
        f = fixture(name, "raw")
        g = O.OFst.from_bytes(f.to_bytes())
        assert f == g and f.props == g.props


# ---- sigma matcher KATs (rustfst-python/tests/algorithms/test_compose.py:157-214, sigma_matcher.rs:487-597)
def acceptor_fst(labels):
    """rustfst utils::acceptor: linear, weight one, last state final."""
    n = len(labels)
    return build(n + 1, 0, [(n, 0.0)], [(i, l, l, 0.0, i + 1) for i, l in enumerate(labels)])


def test_sigma_compose_kat():
    # symt: <eps>=0 play=1 david=2 queen=3 please=4 <sigma>=5
    query = acceptor_fst([1, 3, 4])
    sigma = acceptor_fst([1, 5, 4])
    res = O.compose_sigma(query, sigma, filter=3, connect=True, sigma2=(5, 0, None))
    assert res == query


def test_sigma_compose_with_allowlist_kat():
    # symt: <eps>=0 play=1 bowie=2 queen=3 radiohead=4 please=5 <sigma>=6; allowlist = [queen, bowie]
    sigma = acceptor_fst([1, 6, 5])
    for artist, ok in ((3, True), (2, True), (4, False)):
        q = acceptor_fst([1, artist, 5])
        res = O.compose_sigma(q, sigma, filter=3, connect=True, sigma2=(6, 0, [3, 2]))
        assert (res == q) is ok


def test_sigma_matcher_2_kat():
    """sigma_matcher.rs:548-597: left o right with a sigma matcher on the right has exactly 4 string paths."""
    import os
    ref = "/root/reference/rustfst-tests-data/sigma-matcher-2"
    if not os.path.exists(ref):
        pytest.skip("reference checkout not available")
    left, right = O.OFst.from_path(f"{ref}/left.fst"), O.OFst.from_path(f"{ref}/right.fst")
    left.tr_sort(ilabel=False); right.tr_sort(ilabel=True)
    # <sigma> label: read from the symbol table text? the binary symt lists it; the test uses get_label("<sigma>")
    sig = _sigma_label(f"{ref}/symt.bin")
    res = O.compose_sigma(left, right, filter=3, connect=False, sigma2=(sig, 0, None))
    assert O.count_paths(res) == 4


def _sigma_label(path):
    import struct
    b = open(path, "rb").read()
    off = 4
    n = struct.unpack_from("<i", b, off)[0]; off += 4 + n
    off += 8
    cnt = struct.unpack_from("<q", b, off)[0]; off += 8
    for _ in range(cnt):
        n = struct.unpack_from("<i", b, off)[0]; off += 4
        sym = b[off:off + n].decode(); off += n
        key = struct.unpack_from("<q", b, off)[0]; off += 8
        if sym == "<sigma>":
            return key
    raise AssertionError("no <sigma>")


def test_sigma_errors():
    query, sigma = acceptor_fst([1, 3, 4]), acceptor_fst([1, 5, 4])
    with pytest.raises(O.OracleError, match="AutoFilter"):
        O.compose_sigma(query, sigma, filter=0, connect=True, sigma2=(5, 0, None))
    with pytest.raises(O.OracleError, match="sigma_label"):
        O.compose_sigma(query, sigma, filter=3, connect=True, sigma2=(0, 0, None))
    with pytest.raises(O.OracleError, match="bad label"):
        O.compose_sigma(acceptor_fst([1, 5, 4]), sigma, filter=3, connect=True, sigma2=(5, 0, None))


def test_reverse_kat_from_the_reference_python_tests():
    """rustfst-python/tests/algorithms/test_reverse.py:4-54 — pins reverse(), which the n-best route is built on:
    superinitial state 0, old state s becomes s + 1, in-arcs of a state in (source state, arc position) order."""
    f = O.OFst()
    s1, s2, s3 = f.add_state(), f.add_state(), f.add_state()
    f.set_start(s1)
    f.set_final(s3, 1.0)
    f.add_tr(s1, 1, 2, 1.0, s2)
    f.add_tr(s1, 3, 4, 2.0, s2)
    f.add_tr(s2, 5, 6, 1.5, s2)
    f.add_tr(s2, 3, 5, 1.0, s3)
    e = O.OFst()
    t1, t2, t3, t4 = e.add_state(), e.add_state(), e.add_state(), e.add_state()
    e.set_start(t1)
    e.set_final(t2, 0.0)
    e.add_tr(t1, 0, 0, 1.0, t4)
    e.add_tr(t4, 3, 5, 1.0, t3)
    e.add_tr(t3, 1, 2, 1.0, t2)
    e.add_tr(t3, 3, 4, 2.0, t2)
    e.add_tr(t3, 5, 6, 1.5, t3)
    assert f.reverse() == e


def _build(cls_new, add_tr, arcs, n_states, start, finals):
    f = cls_new()
    for _ in range(n_states):
        f.add_state()
    f.set_start(start)
    for st, w in finals:
        f.set_final(st, w)
    for (src, il, ol, w, dst) in arcs:
        add_tr(f, src, il, ol, w, dst)
    return f


def test_connect_kat_from_the_reference_python_tests():
    """rustfst-python/tests/algorithms/test_connect.py:4-52 — pins connect (trim + order-preserving renumbering)."""
    add = lambda f, s, il, ol, w, d: f.add_tr(s, il, ol, w, d)  # noqa: E731
    f = _build(O.OFst, add, [(4, 1, 2, 1.0, 0), (0, 3, 4, 2.0, 1), (1, 4, 5, 3.0, 2), (2, 4, 6, 4.0, 3), (2, 7, 8, 5.0, 0)],
               5, 0, [(1, 0.0)])
    e = _build(O.OFst, add, [(0, 3, 4, 2.0, 1), (1, 4, 5, 3.0, 2), (2, 7, 8, 5.0, 0)], 3, 0, [(1, 0.0)])
    f.connect()
    assert f == e


def test_tr_sort_kats_from_the_reference_python_tests():
    """rustfst-python/tests/algorithms/test_tr_sort.py:4-95 — stable sort by ilabel / olabel."""
    add = lambda f, s, il, ol, w, d: f.add_tr(s, il, ol, w, d)  # noqa: E731
    arcs = [(0, 1, 2, 1.0, 1), (0, 3, 3, 2.0, 1), (0, 1, 5, 3.0, 1), (0, 2, 6, 4.0, 1)]
    f = _build(O.OFst, add, arcs, 2, 0, [(1, 0.0)])
    f.tr_sort(True)
    assert f == _build(O.OFst, add, [(0, 1, 2, 1.0, 1), (0, 1, 5, 3.0, 1), (0, 2, 6, 4.0, 1), (0, 3, 3, 2.0, 1)], 2, 0, [(1, 0.0)])
    g = _build(O.OFst, add, arcs, 2, 0, [(1, 0.0)])
    g.tr_sort(False)
    assert g == _build(O.OFst, add, [(0, 1, 2, 1.0, 1), (0, 3, 3, 2.0, 1), (0, 1, 5, 3.0, 1), (0, 2, 6, 4.0, 1)], 2, 0, [(1, 0.0)])


def test_tr_sort_kats_through_the_cabi_host_container():
    """The same KATs through fst_tr_sort of the product library: machines this small are sorted by the host container
    (the device path needs >= 64K arcs), so the check runs without a GPU."""
    import rustfst_b200 as R
    add = lambda f, s, il, ol, w, d: f.add_tr(s, R.Tr(il, ol, w, d))  # noqa: E731
    arcs = [(0, 1, 2, 1.0, 1), (0, 3, 3, 2.0, 1), (0, 1, 5, 3.0, 1), (0, 2, 6, 4.0, 1)]
    f = _build(R.VectorFst, add, arcs, 2, 0, [(1, 0.0)])
    f.tr_sort()
    assert f == _build(R.VectorFst, add, [(0, 1, 2, 1.0, 1), (0, 1, 5, 3.0, 1), (0, 2, 6, 4.0, 1), (0, 3, 3, 2.0, 1)], 2, 0, [(1, 0.0)])
    g = _build(R.VectorFst, add, arcs, 2, 0, [(1, 0.0)])
    g.tr_sort(ilabel_cmp=False)
    assert g == _build(R.VectorFst, add, [(0, 1, 2, 1.0, 1), (0, 3, 3, 2.0, 1), (0, 1, 5, 3.0, 1), (0, 2, 6, 4.0, 1)], 2, 0, [(1, 0.0)])
