"""Synthetic acyclic FST generators for the BASELINE.json workloads (seeded, numpy only).

All machines are LAYERED: state 0 is the start (level 0), levels 1..L hold w = (N-1)//L states each, every arc
goes from level l to level l+1, so the machines are acyclic and topologically sorted by state id, every
composition proceeds level by level (BFS depth = L) and the size of the product is controlled by construction:

  * `layered_acceptor`   : out-degree ~ A/N, targets uniform in the next level, labels uniform in [1, V].
  * `bigram_transducer`  : LM-like: the target of an arc is determined by its input label
                           (slot (ilabel-1) % w of the next level), each state carries ~A/N distinct input labels;
                           output labels are independent.  Composing a random acceptor with it pairs every
                           acceptor state with at most (in-degree) transducer states, so |product| ~ 10 x N
                           instead of N^2 — SURVEY.md's uniform-random recipe with V = 20000 yields an EMPTY
                           product (10*10/20000 expected matches per state), which benchmarks nothing.

Weights are dyadic (k/64) so every path sum is exact in f32 (ties are exact, never "near").
Arcs are label-sorted per state (fst1 by olabel, fst2 by ilabel) and the property word says so.
"""
import os

import numpy as np

try:
    from . import props as P
except ImportError:  # loaded by file path (bench.py --impl reference must not map the product library)
    import importlib.util as _ilu
    _spec = _ilu.spec_from_file_location("_b200_props", os.path.join(os.path.dirname(os.path.abspath(__file__)), "props.py"))
    P = _ilu.module_from_spec(_spec)
    _spec.loader.exec_module(P)

TR_DTYPE = np.dtype([("ilabel", "<u4"), ("olabel", "<u4"), ("weight", "<f4"), ("nextstate", "<u4")])


def _level_layout(n_states, levels):
    w = max(1, (n_states - 1) // levels)
    n = 1 + w * levels
    base = np.concatenate([[0], 1 + w * np.arange(levels)]).astype(np.int64)  # first state id of each level
    return n, w, base


def _degrees(rng, n_src, n_arcs):
    mean = max(1.0, n_arcs / max(1, n_src))
    deg = 1 + rng.poisson(mean - 1.0, size=n_src).astype(np.int64)
    return deg


def _weights(rng, size, continuous):
    """Dyadic grid k/64 (every path sum exact in f32, ties exact) or continuous U[0, 10) (near-ties within KDELTA)."""
    if continuous:
        return (rng.random(size) * 10.0).astype(np.float32)
    return rng.integers(0, 640, size=size).astype(np.float32) / np.float32(64.0)


def _finish(n, offsets, arcs, finals, acceptor, sorted_i, sorted_o):
    pr = (P.ACYCLIC | P.INITIAL_ACYCLIC | P.TOP_SORTED | P.NO_EPSILONS | P.NO_I_EPSILONS | P.NO_O_EPSILONS |
          P.WEIGHTED | P.ACCESSIBLE | P.UNWEIGHTED_CYCLES | P.NOT_STRING)
    pr |= P.ACCEPTOR if acceptor else P.NOT_ACCEPTOR
    if sorted_i:
        pr |= P.I_LABEL_SORTED
    if sorted_o:
        pr |= P.O_LABEL_SORTED
    return {"offsets": offsets.astype(np.uint32), "arcs": arcs, "finals": finals.astype(np.float32), "start": 0,
            "props": int(pr), "num_states": int(n)}


def layered_acceptor(n_states, n_arcs, vocab, seed, levels=32, start_fanout=False, continuous=False):
    """Random layered acceptor (ilabel == olabel), arcs sorted by label.  With start_fanout the start state has one
    arc to every state of level 1 (a wide lattice from the first wave on) instead of ~A/N arcs."""
    rng = np.random.default_rng(seed)
    n, w, base = _level_layout(n_states, levels)
    n_src = n - w  # states of the last level have no arcs
    deg = _degrees(rng, n_src, n_arcs)
    if start_fanout:
        deg[0] = w
    offsets = np.zeros(n + 1, dtype=np.int64)
    offsets[1:n_src + 1] = np.cumsum(deg)
    offsets[n_src + 1:] = offsets[n_src]
    total = int(offsets[n_src])
    src = np.repeat(np.arange(n_src, dtype=np.int64), deg)
    level = np.where(src == 0, 0, 1 + (src - 1) // w)
    label = rng.integers(1, vocab + 1, size=total, dtype=np.int64)
    target = base[level + 1] + rng.integers(0, w, size=total, dtype=np.int64)
    if start_fanout:
        target[:w] = base[1] + np.arange(w)
    weight = _weights(rng, total, continuous)
    order = np.lexsort((label, src))  # stable: by src, then label
    arcs = np.zeros(total, dtype=TR_DTYPE)
    arcs["ilabel"] = label[order]; arcs["olabel"] = label[order]
    arcs["weight"] = weight[order]; arcs["nextstate"] = target[order]
    finals = np.full(n, np.inf, dtype=np.float32)
    finals[n - w:] = _weights(rng, w, continuous)
    return _finish(n, offsets, arcs, finals, True, True, True)


def bigram_transducer(n_states, n_arcs, vocab, seed, levels=32, out_vocab=None, start_fanout=False,
                      continuous=False, spread=False):
    """Layered transducer whose arc target is a function of the input label (LM-like); ilabel-sorted.
    With start_fanout the start state carries every input label once.  With spread the target also depends on the
    source state (slot = mix(ilabel, source) % w), so ALL w states of a level are reached and the arc lists a
    composition searches are spread over the whole machine (HBM-resident) instead of `vocab` slots per level."""
    rng = np.random.default_rng(seed)
    out_vocab = out_vocab or vocab
    n, w, base = _level_layout(n_states, levels)
    n_src = n - w
    deg = np.minimum(_degrees(rng, n_src, n_arcs), vocab)
    if start_fanout:
        deg[0] = vocab
    offsets = np.zeros(n + 1, dtype=np.int64)
    offsets[1:n_src + 1] = np.cumsum(deg)
    offsets[n_src + 1:] = offsets[n_src]
    total = int(offsets[n_src])
    src = np.repeat(np.arange(n_src, dtype=np.int64), deg)
    level = np.where(src == 0, 0, 1 + (src - 1) // w)
    ilabel = rng.integers(1, vocab + 1, size=total, dtype=np.int64)
    olabel = rng.integers(1, out_vocab + 1, size=total, dtype=np.int64)
    if start_fanout:
        ilabel[:vocab] = np.arange(1, vocab + 1)
    # a per-level rotation keeps consecutive levels from using the same slots for the same label
    if spread:
        mix = (ilabel * 0x9E3779B1 + src * 0x85EBCA77) & 0xFFFFFFFF
        mix ^= mix >> 15
        target = base[level + 1] + (mix * 0x2C1B3C6D & 0xFFFFFFFF) % w
    else:
        target = base[level + 1] + (ilabel - 1 + 7919 * level) % w
    weight = _weights(rng, total, continuous)
    order = np.lexsort((olabel, ilabel, src))
    arcs = np.zeros(total, dtype=TR_DTYPE)
    arcs["ilabel"] = ilabel[order]; arcs["olabel"] = olabel[order]
    arcs["weight"] = weight[order]; arcs["nextstate"] = target[order]
    finals = np.full(n, np.inf, dtype=np.float32)
    finals[n - w:] = _weights(rng, w, continuous)
    return _finish(n, offsets, arcs, finals, False, True, False)


def window_dag(n_states, n_arcs, vocab, seed, window=1000, continuous=False):
    """SURVEY.md 8d's acyclic acceptor: state s has ~A/N arcs to targets uniform in (s, min(N-1, s+window)], so arcs
    skip levels (a label-correcting relaxation revisits states) and the machine is topologically sorted by state id.
    The last 1 % of the states are final."""
    rng = np.random.default_rng(seed)
    n = int(n_states)
    n_src = n - 1
    deg = _degrees(rng, n_src, n_arcs)
    offsets = np.zeros(n + 1, dtype=np.int64)
    offsets[1:n_src + 1] = np.cumsum(deg)
    offsets[n_src + 1:] = offsets[n_src]
    total = int(offsets[n_src])
    src = np.repeat(np.arange(n_src, dtype=np.int64), deg)
    span = np.minimum(window, n - 1 - src)
    target = src + 1 + (rng.random(total) * span).astype(np.int64)
    target = np.minimum(target, n - 1)
    label = rng.integers(1, vocab + 1, size=total, dtype=np.int64)
    weight = _weights(rng, total, continuous)
    order = np.lexsort((label, src))
    arcs = np.zeros(total, dtype=TR_DTYPE)
    arcs["ilabel"] = label[order]; arcs["olabel"] = label[order]
    arcs["weight"] = weight[order]; arcs["nextstate"] = target[order]
    finals = np.full(n, np.inf, dtype=np.float32)
    nf = max(1, n // 100)
    finals[n - nf:] = _weights(rng, nf, continuous)
    return _finish(n, offsets, arcs, finals, True, True, True)


def linear_acceptor(labels, seed):
    """String acceptor: len(labels)+1 states, arc i carries labels[i] (rustfst utils::acceptor shape, weighted)."""
    rng = np.random.default_rng(seed)
    m = len(labels)
    arcs = np.zeros(m, dtype=TR_DTYPE)
    arcs["ilabel"] = labels; arcs["olabel"] = labels
    arcs["weight"] = rng.integers(0, 640, size=m).astype(np.float32) / np.float32(64.0)
    arcs["nextstate"] = np.arange(1, m + 1)
    offsets = np.concatenate([np.arange(m + 1), [m]])
    finals = np.full(m + 1, np.inf, dtype=np.float32)
    finals[m] = 0.0
    d = _finish(m + 1, offsets, arcs, finals, True, True, True)
    return d


def random_graph_transducer(n_states, n_arcs, vocab, seed):
    """Unstructured (cyclic, self-loops allowed) transducer for the batched workload; ilabel-sorted."""
    rng = np.random.default_rng(seed)
    deg = _degrees(rng, n_states, n_arcs)
    offsets = np.concatenate([[0], np.cumsum(deg)])
    total = int(offsets[-1])
    src = np.repeat(np.arange(n_states, dtype=np.int64), deg)
    ilabel = rng.integers(1, vocab + 1, size=total, dtype=np.int64)
    olabel = rng.integers(1, vocab + 1, size=total, dtype=np.int64)
    target = rng.integers(0, n_states, size=total, dtype=np.int64)
    weight = rng.integers(0, 640, size=total).astype(np.float32) / np.float32(64.0)
    order = np.lexsort((olabel, ilabel, src))
    arcs = np.zeros(total, dtype=TR_DTYPE)
    arcs["ilabel"] = ilabel[order]; arcs["olabel"] = olabel[order]
    arcs["weight"] = weight[order]; arcs["nextstate"] = target[order]
    finals = np.where(rng.random(n_states) < 0.05, rng.integers(0, 640, size=n_states) / 64.0, np.inf)
    pr = (P.NOT_ACCEPTOR | P.NO_EPSILONS | P.NO_I_EPSILONS | P.NO_O_EPSILONS | P.WEIGHTED | P.I_LABEL_SORTED |
          P.CYCLIC | P.NOT_TOP_SORTED | P.NOT_STRING)
    return {"offsets": offsets.astype(np.uint32), "arcs": arcs, "finals": finals.astype(np.float32), "start": 0,
            "props": int(pr), "num_states": int(n_states)}


def sample_path_labels(t, length, seed):
    """Input labels along a random walk in transducer `t` (so that the string is accepted up to dead ends)."""
    rng = np.random.default_rng(seed)
    off, arcs = t["offsets"], t["arcs"]
    s = t["start"]
    labels = np.zeros(length, dtype=np.int64)
    for i in range(length):
        lo, hi = int(off[s]), int(off[s + 1])
        if hi == lo:
            labels[i:] = rng.integers(1, 1 + int(arcs["ilabel"].max()), size=length - i)
            break
        a = arcs[rng.integers(lo, hi)]
        labels[i] = a["ilabel"]
        s = int(a["nextstate"])
    return labels


def sample_path_labels_batch(t, length, count, seed):
    """`count` label strings at once (vectorised random walks in `t`; the rest of a walk that hits a dead end is filled
    with random labels).  Returns an int64 array (count, length)."""
    rng = np.random.default_rng(seed)
    off, arcs = t["offsets"].astype(np.int64), t["arcs"]
    vmax = int(arcs["ilabel"].max()) if len(arcs) else 1
    s = np.full(count, t["start"], dtype=np.int64)
    alive = np.ones(count, dtype=bool)
    labels = rng.integers(1, 1 + vmax, size=(count, length), dtype=np.int64)
    for i in range(length):
        lo, hi = off[s], off[s + 1]
        alive &= hi > lo
        pick = lo + (rng.random(count) * np.maximum(hi - lo, 1)).astype(np.int64)
        pick = np.minimum(pick, np.maximum(hi - 1, 0))
        a = arcs[np.where(alive, pick, 0)]
        labels[alive, i] = a["ilabel"][alive]
        s = np.where(alive, a["nextstate"].astype(np.int64), s)
    return labels


def to_vector_fst(d):
    from .fst import VectorFst
    return VectorFst.from_csr(d["offsets"], d["arcs"], d["finals"], d["start"], d["props"])


def workload(name, scale=1.0):
    """The BASELINE.json configurations (C2..C4) as generator calls; `scale` shrinks them for tests."""
    if name == "C2":
        n, a, lv, v = int(100_000 * scale), int(1_000_000 * scale), 25, 32
        return layered_acceptor(n, a, v, 1, lv), bigram_transducer(n, a, v, 2, lv, out_vocab=2000)
    if name == "C3":
        n, a, lv, v = int(1_000_000 * scale), int(10_000_000 * scale), 50, 32
        return layered_acceptor(n, a, v, 3, lv), bigram_transducer(n, a, v, 4, lv, out_vocab=20000)
    if name == "C4":
        n, a, lv, v = int(5_000_000 * scale), int(50_000_000 * scale), 50, 1000
        return layered_acceptor(n, a, v, 6, lv)
    raise ValueError(name)
