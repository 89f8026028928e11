"""ctypes binding of the CPU oracle (oracle/liboracle.so). TEST INFRASTRUCTURE ONLY.

The oracle restates the reference's compose / shortest_path (see oracle/oracle.hpp); tests use it as the
checker for the CUDA path.  It is built on demand with `make -C oracle`.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

TR_DTYPE = np.dtype([("ilabel", "<u4"), ("olabel", "<u4"), ("weight", "<f4"), ("nextstate", "<u4")])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        src = [os.path.join(ROOT, "oracle", f) for f in ("oracle.hpp", "oracle_capi.cpp")]
        if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in src):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"],
                                  stdout=subprocess.DEVNULL)
        L = C.CDLL(path)
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_fst_new.restype = C.c_void_p
        L.oracle_fst_free.argtypes = [C.c_void_p]
        L.oracle_fst_add_state.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        L.oracle_fst_set_start.argtypes = [C.c_void_p, C.c_uint32]
        L.oracle_fst_set_final.argtypes = [C.c_void_p, C.c_uint32, C.c_float]
        L.oracle_fst_add_tr.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_uint32]
        L.oracle_fst_tr_sort.argtypes = [C.c_void_p, C.c_int]
        L.oracle_fst_connect.argtypes = [C.c_void_p]
        L.oracle_fst_top_sort.argtypes = [C.c_void_p]
        L.oracle_fst_reverse.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.oracle_fst_compute_props.argtypes = [C.c_void_p]
        L.oracle_fst_props.argtypes = [C.c_void_p]
        L.oracle_fst_props.restype = C.c_uint64
        L.oracle_fst_set_props.argtypes = [C.c_void_p, C.c_uint64]
        L.oracle_fst_num_states.argtypes = [C.c_void_p]
        L.oracle_fst_num_states.restype = C.c_uint64
        L.oracle_fst_num_trs.argtypes = [C.c_void_p]
        L.oracle_fst_num_trs.restype = C.c_uint64
        L.oracle_fst_start.argtypes = [C.c_void_p]
        L.oracle_fst_start.restype = C.c_int64
        L.oracle_fst_equal.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_fst_from_bytes.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_void_p)]
        L.oracle_fst_to_bytes.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.oracle_fst_to_csr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_fst_from_csr.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint64,
                                          C.POINTER(C.c_void_p)]
        L.oracle_compose.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        L.oracle_compose_sigma.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.oracle_queue_plan.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
        L.oracle_count_paths.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.oracle_shortest_path.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_float, C.POINTER(C.c_void_p),
                                           C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.c_void_p]
        _LIB = L
    return _LIB


class OracleError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise OracleError(lib().oracle_last_error().decode())


class OFst:
    """Oracle-side VectorFst<TropicalWeight> (mirrors the mutation API used by the reference's tests)."""

    def __init__(self, ptr=None):
        self.ptr = C.c_void_p(lib().oracle_fst_new()) if ptr is None else ptr

    def __del__(self):
        try:
            lib().oracle_fst_free(self.ptr)
        except Exception:
            pass

    # -- construction
    def add_state(self):
        out = C.c_uint32()
        _check(lib().oracle_fst_add_state(self.ptr, C.byref(out)))
        return out.value

    def set_start(self, s):
        _check(lib().oracle_fst_set_start(self.ptr, s))

    def set_final(self, s, w=0.0):
        _check(lib().oracle_fst_set_final(self.ptr, s, w))

    def add_tr(self, s, il, ol, w, ns):
        _check(lib().oracle_fst_add_tr(self.ptr, s, il, ol, w, ns))

    def tr_sort(self, ilabel=True):
        _check(lib().oracle_fst_tr_sort(self.ptr, 1 if ilabel else 0))

    def connect(self):
        _check(lib().oracle_fst_connect(self.ptr))

    def top_sort(self):
        _check(lib().oracle_fst_top_sort(self.ptr))

    def reverse(self):
        out = C.c_void_p()
        _check(lib().oracle_fst_reverse(self.ptr, C.byref(out)))
        return OFst(out)

    def compute_props(self):
        _check(lib().oracle_fst_compute_props(self.ptr))

    # -- inspection
    @property
    def props(self):
        return lib().oracle_fst_props(self.ptr)

    @props.setter
    def props(self, p):
        lib().oracle_fst_set_props(self.ptr, p)

    @property
    def num_states(self):
        return lib().oracle_fst_num_states(self.ptr)

    @property
    def num_trs(self):
        return lib().oracle_fst_num_trs(self.ptr)

    @property
    def start(self):
        s = lib().oracle_fst_start(self.ptr)
        return None if s < 0 else s

    def __eq__(self, other):
        return bool(lib().oracle_fst_equal(self.ptr, other.ptr))

    # -- I/O
    @staticmethod
    def from_bytes(b):
        out = C.c_void_p()
        _check(lib().oracle_fst_from_bytes(b, len(b), C.byref(out)))
        return OFst(out)

    @staticmethod
    def from_path(path):
        with open(path, "rb") as f:
            return OFst.from_bytes(f.read())

    def to_bytes(self):
        size = C.c_uint64()
        _check(lib().oracle_fst_to_bytes(self.ptr, None, 0, C.byref(size)))
        buf = C.create_string_buffer(size.value)
        _check(lib().oracle_fst_to_bytes(self.ptr, buf, size.value, C.byref(size)))
        return buf.raw

    def to_csr(self):
        n, a = self.num_states, self.num_trs
        offsets = np.zeros(n + 1, dtype=np.uint64)
        arcs = np.zeros(a, dtype=TR_DTYPE)
        finals = np.zeros(n, dtype=np.float32)
        _check(lib().oracle_fst_to_csr(self.ptr, offsets.ctypes.data, arcs.ctypes.data, finals.ctypes.data))
        return offsets, arcs, finals

    @staticmethod
    def from_csr(offsets, arcs, finals, start, props):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        arcs = np.ascontiguousarray(arcs, dtype=TR_DTYPE)
        finals = np.ascontiguousarray(finals, dtype=np.float32)
        out = C.c_void_p()
        _check(lib().oracle_fst_from_csr(len(finals), offsets.ctypes.data, arcs.ctypes.data, finals.ctypes.data,
                                         -1 if start is None else int(start), int(props), C.byref(out)))
        return OFst(out)


def compose(a, b, filter=0, connect=True, want_stats=False):
    out = C.c_void_p()
    stats = (C.c_uint64 * 3)()
    secs = C.c_double()
    _check(lib().oracle_compose(a.ptr, b.ptr, int(filter), 1 if connect else 0, C.byref(out), stats, C.byref(secs)))
    r = OFst(out)
    if want_stats:
        return r, {"states_expanded": stats[0], "arcs_iterated": stats[1], "arcs_emitted": stats[2],
                   "seconds": secs.value}
    return r


def _flat_sigma(cfg):
    if cfg is None:
        return [0, 0, 0, 0]
    label, mode, allowed = cfg
    allowed = list(allowed or [])
    return [1, label, mode, len(allowed)] + allowed


def compose_sigma(a, b, filter, connect, sigma1=None, sigma2=None):
    """sigma1 / sigma2 = (sigma_label, rewrite_mode, allowed list or None) for matcher1 / matcher2."""
    flat = np.array(_flat_sigma(sigma1) + _flat_sigma(sigma2), dtype=np.uint32)
    out = C.c_void_p()
    _check(lib().oracle_compose_sigma(a.ptr, b.ptr, int(filter), 1 if connect else 0, flat.ctypes.data, C.byref(out)))
    return OFst(out)


def queue_plan(a):
    n = a.num_states
    kind, n_scc = C.c_int32(), C.c_uint32()
    order = np.zeros(max(1, n), dtype=np.uint32)
    fifo = np.zeros(max(1, n), dtype=np.uint8)
    _check(lib().oracle_queue_plan(a.ptr, C.byref(kind), order.ctypes.data, fifo.ctypes.data, C.byref(n_scc)))
    return kind.value, order[:n], fifo[:n_scc.value]


def count_paths(a):
    n = C.c_uint64()
    _check(lib().oracle_count_paths(a.ptr, C.byref(n)))
    return n.value


def shortest_path(a, nshortest=1, unique=False, delta=1e-6, want_stats=False, want_distance=False):
    out = C.c_void_p()
    stats = (C.c_uint64 * 2)()
    secs = C.c_double()
    dist = np.zeros(a.num_states, dtype=np.float32) if want_distance else None
    _check(lib().oracle_shortest_path(a.ptr, nshortest, 1 if unique else 0, delta, C.byref(out), stats,
                                      C.byref(secs), dist.ctypes.data if want_distance else None))
    r = OFst(out)
    extra = {"arcs_relaxed": stats[0], "states_dequeued": stats[1], "seconds": secs.value}
    if want_distance:
        extra["distance"] = dist
    if want_stats or want_distance:
        return r, extra
    return r
