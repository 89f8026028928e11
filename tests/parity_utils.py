"""Helpers shared by the parity tests: build the same FST on the product side (rustfst_b200.VectorFst over the
C-ABI) and on the oracle side (tests/oracle_lib.OFst), and compare results bit-for-bit."""
import os

import numpy as np

from tests import oracle_lib as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FIXTURES = ["fst_000", "fst_001", "fst_002", "fst_003", "fst_004", "fst_006", "fst_007", "fst_008", "fst_009",
            "fst_012", "fst_013", "fst_014", "fst_015", "fst_016", "fst_017", "fst_018", "fst_019", "fst_020"]


def golden_path(name, which):
    return os.path.join(GOLDEN, f"{name}_{which}.fst")


def both_from_path(path):
    import rustfst_b200 as R
    return R.VectorFst.read(path), O.OFst.from_path(path)


def both_from_dict(d):
    import rustfst_b200 as R
    p = R.VectorFst.from_csr(d["offsets"], d["arcs"], d["finals"], d["start"], d["props"])
    o = O.OFst.from_csr(d["offsets"].astype(np.uint64), d["arcs"], d["finals"], d["start"], d["props"])
    return p, o


def assert_same(prod, orc, what="", check_props=True):
    """Bit-exact comparison: #states, start, offsets, every arc field (weights by bit pattern), finals, properties."""
    po, pa, pf, ps = prod.to_csr()
    oo, oa, of, = orc.to_csr()
    assert len(pf) == len(of), f"{what}: num_states {len(pf)} != {len(of)}"
    assert ps == orc.start, f"{what}: start {ps} != {orc.start}"
    assert np.array_equal(po.astype(np.uint64), oo), f"{what}: state->arc offsets differ"
    for field in ("ilabel", "olabel", "nextstate"):
        assert np.array_equal(pa[field], oa[field]), f"{what}: arc field {field} differs"
    assert np.array_equal(pa["weight"].view(np.uint32), oa["weight"].view(np.uint32)), f"{what}: arc weights differ"
    assert np.array_equal(pf.view(np.uint32), of.view(np.uint32)), f"{what}: final weights differ"
    if check_props:
        assert prod.properties == orc.props, f"{what}: properties {prod.properties:#x} != {orc.props:#x}"


def random_fst(rng, n_states, max_arcs, n_labels, eps_prob=0.0, acceptor=False, sort=None, cyclic=True,
               weight_grid=True, final_prob=0.3):
    """proptest-style random FST (rustfst/src/proptest_fst/mod.rs:7-10 idea) as a CSR dict."""
    from rustfst_b200 import props as P
    from rustfst_b200.fst import TR_DTYPE
    offsets = [0]
    rows = []
    for s in range(n_states):
        k = int(rng.integers(0, max_arcs + 1))
        arcs = []
        for _ in range(k):
            il = 0 if rng.random() < eps_prob else int(rng.integers(1, n_labels + 1))
            ol = il if acceptor else (0 if rng.random() < eps_prob else int(rng.integers(1, n_labels + 1)))
            w = float(rng.integers(0, 64)) / 8.0 if weight_grid else float(np.float32(rng.random() * 10))
            if cyclic:
                ns = int(rng.integers(0, n_states))
            else:
                if s + 1 >= n_states:
                    continue
                ns = int(rng.integers(s + 1, n_states))
            arcs.append((il, ol, w, ns))
        if sort == "ilabel":
            arcs.sort(key=lambda a: a[0])
        elif sort == "olabel":
            arcs.sort(key=lambda a: a[1])
        rows.extend(arcs)
        offsets.append(len(rows))
    arr = np.zeros(len(rows), dtype=TR_DTYPE)
    for i, (il, ol, w, ns) in enumerate(rows):
        arr[i] = (il, ol, w, ns)
    finals = np.full(n_states, np.inf, dtype=np.float32)
    for s in range(n_states):
        if rng.random() < final_prob:
            finals[s] = float(rng.integers(0, 64)) / 8.0 if weight_grid else float(np.float32(rng.random() * 10))
    d = {"offsets": np.array(offsets, dtype=np.uint32), "arcs": arr, "finals": finals,
         "start": 0 if n_states else None, "props": 0, "num_states": n_states}
    # let the oracle compute the full property word (as rustfst-tests-data/main.cpp does for its operands)
    o = O.OFst.from_csr(d["offsets"].astype(np.uint64), arr, finals, d["start"], 0)
    o.compute_props()
    d["props"] = o.props
    return d
