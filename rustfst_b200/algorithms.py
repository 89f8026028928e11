"""compose / shortest path — mirror of rustfst-python/rustfst/algorithms/{compose,shortest_path}.py."""
from __future__ import annotations

import ctypes as C
from enum import Enum
from typing import List, Optional, Union

from .ffi import CIntArray, ComposeStats, SsspStats, check_ffi_error, lib
from .fst import VectorFst

KSHORTESTDELTA = 1e-6


class MatcherRewriteMode(Enum):  # compose.py:17-20
    AUTO = 0
    ALWAYS = 1
    NEVER = 2


class MatcherConfig:  # compose.py:27-55
    def __init__(self, sigma_label: int, rewrite_mode: MatcherRewriteMode = MatcherRewriteMode.AUTO,
                 sigma_allowed_matches: Optional[List[int]] = None):
        array = sigma_allowed_matches or []
        arr = CIntArray()
        arr.size = len(array)
        self._keep = (C.c_uint32 * len(array))(*array)
        arr.data = C.cast(self._keep, C.POINTER(C.c_uint32))
        self.ptr = C.c_void_p()
        check_ffi_error(lib.fst_matcher_config_new(sigma_label, rewrite_mode.value, arr, C.byref(self.ptr)),
                        "Error creating MatcherConfig")

    def __del__(self):
        try:
            lib.fst_matcher_config_destroy(self.ptr)
        except Exception:
            pass


class ComposeFilter(Enum):  # compose.py:58-65
    AUTOFILTER = 0
    NULLFILTER = 1
    TRIVIALFILTER = 2
    SEQUENCEFILTER = 3
    ALTSEQUENCEFILTER = 4
    MATCHFILTER = 5
    NOMATCHFILTER = 6


class ComposeConfig:  # compose.py:68-107
    def __init__(self, compose_filter: ComposeFilter = ComposeFilter.AUTOFILTER, connect: bool = True,
                 matcher1_config: Optional[MatcherConfig] = None, matcher2_config: Optional[MatcherConfig] = None):
        self.ptr = C.c_void_p()
        value = compose_filter.value if isinstance(compose_filter, ComposeFilter) else int(compose_filter)
        check_ffi_error(lib.fst_compose_config_new(value, bool(connect),
                                                   matcher1_config.ptr if matcher1_config else None,
                                                   matcher2_config.ptr if matcher2_config else None,
                                                   C.byref(self.ptr)), "Error creating ComposeConfig")

    def __del__(self):
        try:
            lib.fst_compose_config_destroy(self.ptr)
        except Exception:
            pass


def compose(fst: VectorFst, other_fst: VectorFst) -> VectorFst:  # compose.py:110-125
    out = C.c_void_p()
    check_ffi_error(lib.fst_compose(fst.ptr, other_fst.ptr, C.byref(out)), "Error Composing FSTs")
    return VectorFst(ptr=out)


def compose_with_config(fst: VectorFst, other_fst: VectorFst, config: ComposeConfig) -> VectorFst:  # :128-148
    out = C.c_void_p()
    check_ffi_error(lib.fst_compose_with_config(fst.ptr, other_fst.ptr, config.ptr, C.byref(out)),
                    "Error Composing FSTs")
    return VectorFst(ptr=out)


def compose_with_stats(fst: VectorFst, other_fst: VectorFst, config: Optional[ComposeConfig] = None):
    """b200 addition: same call through host buffers, plus device counters/timings."""
    out = C.c_void_p()
    st = ComposeStats()
    check_ffi_error(lib.b200_compose_with_stats(fst.ptr, other_fst.ptr, config.ptr if config else None,
                                                C.byref(out), C.byref(st)), "Error Composing FSTs")
    return VectorFst(ptr=out), st.as_dict()


class ShortestPathConfig:  # shortest_path.py:15-41
    def __init__(self, nshortest: int = 1, unique: bool = False, delta: Union[float, None] = None):
        if delta is None:
            delta = KSHORTESTDELTA
        self.ptr = C.c_void_p()
        check_ffi_error(lib.fst_shortest_path_config_new(delta, nshortest, bool(unique), C.byref(self.ptr)),
                        "Error creating ShortestPathConfig")

    def __del__(self):
        try:
            lib.b200_shortest_path_config_destroy(self.ptr)
        except Exception:
            pass


def shortestpath(fst: VectorFst) -> VectorFst:  # shortest_path.py:44-57
    out = C.c_void_p()
    check_ffi_error(lib.fst_shortest_path(fst.ptr, C.byref(out)), "Error computing shortest path")
    return VectorFst(ptr=out)


def shortestpath_with_config(fst: VectorFst, config: ShortestPathConfig) -> VectorFst:  # shortest_path.py:60-76
    out = C.c_void_p()
    check_ffi_error(lib.fst_shortest_path_with_config(fst.ptr, config.ptr, C.byref(out)),
                    "Error computing shortest path")
    return VectorFst(ptr=out)


def shortestpath_with_stats(fst: VectorFst, config: Optional[ShortestPathConfig] = None, force_serial=False):
    out = C.c_void_p()
    st = SsspStats()
    check_ffi_error(lib.b200_shortest_path_with_stats(fst.ptr, config.ptr if config else None, C.byref(out),
                                                      C.byref(st), bool(force_serial)),
                    "Error computing shortest path")
    return VectorFst(ptr=out), st.as_dict()


class DeviceFst:
    """An FST resident in HBM (b200 addition): inputs uploaded once, results left on the device."""

    def __init__(self, ptr, host: Optional[VectorFst] = None):
        self.ptr = ptr
        self.host = host

    @classmethod
    def upload(cls, fst: VectorFst) -> "DeviceFst":
        p = C.c_void_p()
        check_ffi_error(lib.b200_device_fst_upload(fst.ptr, C.byref(p)), "Error uploading FST")
        return cls(p, fst)

    def download(self) -> VectorFst:
        p = C.c_void_p()
        check_ffi_error(lib.b200_device_fst_download(self.ptr, C.byref(p)), "Error downloading FST")
        return VectorFst(p)

    def info(self):
        n, a, pr = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check_ffi_error(lib.b200_device_fst_info(self.ptr, C.byref(n), C.byref(a), C.byref(pr)), "info failed")
        return n.value, a.value, pr.value

    def isomorphic(self, other: "DeviceFst") -> Optional[bool]:
        """isomorphic() of two machines in HBM (no download).  None = undecided on the device (the reference's answer
        depends on its visiting order there): use VectorFst.isomorphic on the host copies."""
        r = C.c_int32()
        check_ffi_error(lib.b200_device_isomorphic(self.ptr, other.ptr, C.byref(r)), "Error during isomorphic")
        return None if r.value < 0 else bool(r.value)

    def __del__(self):
        try:
            lib.b200_device_fst_destroy(self.ptr)
        except Exception:
            pass


def device_compose(a: DeviceFst, b: DeviceFst, config: Optional[ComposeConfig] = None):
    out = C.c_void_p()
    st = ComposeStats()
    check_ffi_error(lib.b200_device_compose(a.ptr, b.ptr, config.ptr if config else None, C.byref(out), C.byref(st)),
                    "Error Composing FSTs")
    return DeviceFst(out), st.as_dict()


def device_shortest_path(d: DeviceFst, plan_from: Optional[VectorFst] = None, force_serial=False,
                         config: Optional[ShortestPathConfig] = None):
    host = plan_from or d.host
    hp = host.ptr if host is not None else None
    out = C.c_void_p()
    st = SsspStats()
    if config is None:
        rc = lib.b200_device_shortest_path(d.ptr, hp, C.byref(out), C.byref(st), bool(force_serial))
    else:
        rc = lib.b200_device_shortest_path_with_config(d.ptr, hp, config.ptr, C.byref(out), C.byref(st),
                                                       bool(force_serial))
    check_ffi_error(rc, "Error computing shortest path")
    return VectorFst(out), st.as_dict()


class PackedBatch:
    """The results of a batched composition as ONE block (b200_compose_batch_packed): fetch results one by one with
    result(i), or move the whole block as bytes (what a rank sends to rank 0 in the sharded mode)."""

    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        if getattr(self, "ptr", None):
            lib.b200_packed_batch_destroy(self.ptr)
            self.ptr = None

    def info(self):
        n, st, tr, by = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        check_ffi_error(lib.b200_packed_batch_info(self.ptr, C.byref(n), C.byref(st), C.byref(tr), C.byref(by)), "info")
        return {"n": n.value, "num_states": st.value, "num_trs": tr.value, "bytes": by.value}

    def __len__(self):
        return self.info()["n"]

    def result(self, i: int) -> VectorFst:
        out = C.c_void_p()
        check_ffi_error(lib.b200_packed_batch_get(self.ptr, i, C.byref(out)), "Error fetching a batch result")
        return VectorFst(out)

    def to_numpy(self, out=None):
        """The block as a uint8 array (one copy).  `out`: a caller-owned uint8 array to serialise into (e.g. a view of a
        page-locked buffer that is reused from call to call); the returned array is the used prefix of it."""
        import numpy as np
        nbytes = self.info()["bytes"]
        if out is None:
            buf = np.empty(nbytes, dtype=np.uint8)
        else:
            if out.dtype != np.uint8 or out.ndim != 1 or out.nbytes < nbytes or not out.flags["C_CONTIGUOUS"]:
                raise ValueError(f"PackedBatch.to_numpy: `out` must be a contiguous uint8 array of at least {nbytes} bytes")
            buf = out[:nbytes]
        check_ffi_error(lib.b200_packed_batch_serialize(self.ptr, buf.ctypes.data, buf.nbytes), "Error serialising")
        return buf

    def to_bytes(self) -> bytes:
        return self.to_numpy().tobytes()

    @classmethod
    def from_buffer(cls, buf) -> "PackedBatch":
        import numpy as np
        a = np.frombuffer(buf, dtype=np.uint8)
        out = C.c_void_p()
        check_ffi_error(lib.b200_packed_batch_deserialize(a.ctypes.data, a.nbytes, C.byref(out)), "Error deserialising")
        return cls(out)


class AcceptorBatch:
    """The `const CFst* const*` argument of the batched entry points, built once for a list of acceptors (a C caller
    already holds such an array; building it from 8192 Python wrappers costs ~3 ms per call otherwise)."""

    def __init__(self, acceptors: List[VectorFst]):
        self.acceptors = list(acceptors)  # keeps the handles alive
        self.n = len(self.acceptors)
        self.array = (C.c_void_p * self.n)(*[getattr(a.ptr, "value", a.ptr) for a in self.acceptors])


def compose_batch_packed(acceptors, transducer: Optional[VectorFst] = None,
                         config: Optional[ComposeConfig] = None, device_transducer: Optional["DeviceFst"] = None):
    """acceptors[i] o transducer for all i in one device BFS; `acceptors` is a list of VectorFst or an AcceptorBatch; the
    transducer is a host FST (uploaded by the call) or a DeviceFst that stays resident in HBM across calls."""
    if not isinstance(acceptors, AcceptorBatch):
        acceptors = AcceptorBatch(acceptors)
    n, ins = acceptors.n, acceptors.array
    out = C.c_void_p()
    st = ComposeStats()
    check_ffi_error(lib.b200_compose_batch_packed(ins, n, transducer.ptr if transducer is not None else None,
                                                  device_transducer.ptr if device_transducer is not None else None,
                                                  config.ptr if config else None, C.byref(out), C.byref(st)),
                    "Error in batched compose")
    return PackedBatch(out), st.as_dict()


def compose_batch(acceptors: List[VectorFst], transducer: VectorFst, config: Optional[ComposeConfig] = None):
    n = len(acceptors)
    ins = (C.c_void_p * n)(*[getattr(a.ptr, "value", a.ptr) for a in acceptors])
    outs = (C.c_void_p * n)()
    st = ComposeStats()
    check_ffi_error(lib.b200_compose_batch(ins, n, transducer.ptr, config.ptr if config else None, outs, C.byref(st)),
                    "Error in batched compose")
    # thousands of results: build the wrappers without going through __init__ (raw handle values are fine as `ptr`,
    # every entry point declares its argument types)
    results = []
    new = VectorFst.__new__
    for p in outs:
        v = new(VectorFst)
        v.ptr = p
        results.append(v)
    return results, st.as_dict()
