"""VectorFst / Tr — host-side mirror of rustfst-python/rustfst/fst/vector_fst.py and rustfst/tr.py over the C-ABI
of librustfst_b200.so (same class and method names, same argument meaning and error behaviour)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from .ffi import CArrayU8, CTr, check_ffi_error, lib

TR_DTYPE = np.dtype([("ilabel", "<u4"), ("olabel", "<u4"), ("weight", "<f4"), ("nextstate", "<u4")])


def weight_one() -> float:
    w = C.c_float()
    check_ffi_error(lib.fst_weight_one(C.byref(w)), "weight_one failed")
    return w.value


def weight_zero() -> float:
    w = C.c_float()
    check_ffi_error(lib.fst_weight_zero(C.byref(w)), "weight_zero failed")
    return w.value


class Tr:
    """rustfst-python/rustfst/tr.py:17-153"""

    def __init__(self, ilabel=None, olabel=None, weight=None, nextstate=None):
        if ilabel is not None and olabel is None and weight is None and nextstate is None:
            self._ptr = ilabel  # wrap an existing pointer
        else:
            if weight is None:
                weight = weight_one()
            ptr = C.c_void_p()
            check_ffi_error(lib.tr_new(ilabel, olabel, weight, nextstate, C.byref(ptr)),
                            "Something went wrong when creating the Tr struct")
            self._ptr = ptr

    @property
    def ptr(self):
        return self._ptr

    def _get(self, fn, ctype):
        v = ctype()
        check_ffi_error(fn(self._ptr, C.byref(v)), "Something went wrong when reading Tr")
        return v.value

    ilabel = property(lambda self: int(self._get(lib.tr_ilabel, C.c_uint32)),
                      lambda self, v: check_ffi_error(lib.tr_set_ilabel(self._ptr, v), "tr_set_ilabel"))
    olabel = property(lambda self: int(self._get(lib.tr_olabel, C.c_uint32)),
                      lambda self, v: check_ffi_error(lib.tr_set_olabel(self._ptr, v), "tr_set_olabel"))
    weight = property(lambda self: self._get(lib.tr_weight, C.c_float),
                      lambda self, v: check_ffi_error(lib.tr_set_weight(self._ptr, v), "tr_set_weight"))
    next_state = property(lambda self: int(self._get(lib.tr_next_state, C.c_uint32)),
                          lambda self, v: check_ffi_error(lib.tr_set_next_state(self._ptr, v), "tr_set_next_state"))

    def __eq__(self, other):
        return (self.ilabel == other.ilabel and self.olabel == other.olabel and self.weight == other.weight
                and self.next_state == other.next_state)

    def __repr__(self):
        return f"<Tr ilabel={self.ilabel}, olabel={self.olabel}, weight={self.weight}, next_state={self.next_state}>"

    def __del__(self):
        try:
            lib.tr_delete(self._ptr)
        except Exception:
            pass


class TrsIterator:
    """rustfst-python/rustfst/iterators.py (TrsIterator)"""

    def __init__(self, fst: "VectorFst", state: int):
        self._ptr = C.c_void_p()
        check_ffi_error(lib.trs_iterator_new(fst.ptr, state, C.byref(self._ptr)), "trs_iterator_new failed")
        if not self._ptr:
            raise ValueError(f"State {state} doesn't exist")

    def done(self) -> bool:
        d = C.c_size_t()
        check_ffi_error(lib.trs_iterator_done(self._ptr, C.byref(d)), "trs_iterator_done failed")
        return bool(d.value)

    def __iter__(self):
        return self

    def __next__(self) -> Tr:
        p = C.c_void_p()
        check_ffi_error(lib.trs_iterator_next(self._ptr, C.byref(p)), "trs_iterator_next failed")
        if not p:
            raise StopIteration
        return Tr(p)

    def reset(self):
        check_ffi_error(lib.trs_iterator_reset(self._ptr), "trs_iterator_reset failed")

    def __del__(self):
        try:
            lib.trs_iterator_destroy(self._ptr)
        except Exception:
            pass


class VectorFst:
    """rustfst-python/rustfst/fst/vector_fst.py:33-640 (the part that feeds and inspects compose / shortest path)."""

    def __init__(self, ptr=None):
        if ptr is None:
            ptr = C.c_void_p()
            check_ffi_error(lib.vec_fst_new(C.byref(ptr)), "Something went wrong when creating the Fst struct")
        self.ptr = ptr

    def __del__(self):
        try:
            lib.fst_destroy(self.ptr)
        except Exception:
            pass

    # ---- construction
    def add_state(self) -> int:
        s = C.c_uint32()
        check_ffi_error(lib.vec_fst_add_state(self.ptr, C.byref(s)), "Error during `add_state`")
        return s.value

    def add_tr(self, state: int, tr: Tr) -> "VectorFst":
        check_ffi_error(lib.vec_fst_add_tr(self.ptr, state, tr.ptr), "Error during `add_tr`")
        return self

    def set_start(self, state: int) -> "VectorFst":
        check_ffi_error(lib.vec_fst_set_start(self.ptr, state), "Error setting start state")
        return self

    def set_final(self, state: int, weight: Optional[float] = None) -> "VectorFst":
        if weight is None:
            weight = weight_one()
        check_ffi_error(lib.vec_fst_set_final(self.ptr, state, weight), "Error setting final state")
        return self

    def unset_final(self, state: int):
        check_ffi_error(lib.vec_fst_del_final_weight(self.ptr, state), "Error unsetting final state")

    def delete_states(self):
        check_ffi_error(lib.vec_fst_delete_states(self.ptr), "Error deleting states")

    # ---- inspection
    def start(self) -> Optional[int]:
        slot = C.c_uint32(0xFFFFFFFF)  # the callee leaves the slot untouched when there is no start state
        check_ffi_error(lib.fst_start(self.ptr, C.byref(slot)), "Error getting start state")
        return None if slot.value == 0xFFFFFFFF else slot.value

    def final(self, state: int) -> Optional[float]:
        is_final = C.c_size_t()
        check_ffi_error(lib.fst_is_final(self.ptr, state, C.byref(is_final)), "Error checking if the state is final")
        if not is_final.value:
            return None
        w = C.c_float()
        check_ffi_error(lib.fst_final_weight(self.ptr, state, C.byref(w)), "Error getting final weight")
        return w.value

    def is_final(self, state: int) -> bool:
        return self.final(state) is not None

    def is_start(self, state: int) -> bool:
        r = C.c_size_t()
        check_ffi_error(lib.fst_is_start(self.ptr, state, C.byref(r)), "Error checking if the state is start")
        return bool(r.value)

    def num_states(self) -> int:
        n = C.c_size_t()
        check_ffi_error(lib.vec_fst_num_states(self.ptr, C.byref(n)), "Error getting number of states")
        return n.value

    def num_trs(self, state: int) -> int:
        n = C.c_size_t()
        check_ffi_error(lib.fst_num_trs(self.ptr, state, C.byref(n)), "Error getting number of trs")
        return n.value

    def num_trs_total(self) -> int:
        n = C.c_uint64()
        check_ffi_error(lib.b200_fst_num_trs_total(self.ptr, C.byref(n)), "Error getting number of trs")
        return n.value

    def trs(self, state: int) -> TrsIterator:
        return TrsIterator(self, state)

    def states(self):
        return iter(range(self.num_states()))

    @property
    def properties(self) -> int:
        p = C.c_uint64()
        check_ffi_error(lib.b200_fst_properties(self.ptr, C.byref(p)), "Error getting properties")
        return p.value

    @properties.setter
    def properties(self, p: int):
        check_ffi_error(lib.b200_fst_set_properties(self.ptr, p), "Error setting properties")

    def compute_properties(self) -> int:
        """compute_and_update_properties_all: fills in every unknown property bit from the content (b200 addition)."""
        p = C.c_uint64()
        check_ffi_error(lib.b200_fst_compute_properties(self.ptr, C.byref(p)), "Error computing properties")
        return p.value

    # ---- algorithms (rustfst-python/rustfst/fst/vector_fst.py:419-436, 621-638, tr_sort, connect)
    def compose(self, other: "VectorFst", config=None) -> "VectorFst":
        from .algorithms import compose, compose_with_config
        return compose_with_config(self, other, config) if config else compose(self, other)

    def shortest_path(self, config=None) -> "VectorFst":
        from .algorithms import shortestpath, shortestpath_with_config
        return shortestpath_with_config(self, config) if config else shortestpath(self)

    def tr_sort(self, ilabel_cmp: bool = True):
        check_ffi_error(lib.fst_tr_sort(self.ptr, bool(ilabel_cmp)), "Error during tr_sort")

    def isomorphic(self, other: "VectorFst") -> bool:  # vector_fst.py:718-728
        r = C.c_size_t()
        check_ffi_error(lib.fst_isomorphic(self.ptr, other.ptr, C.byref(r)), "Error during isomorphic")
        return bool(r.value)

    def top_sort(self) -> "VectorFst":  # vector_fst.py (top_sort) / algorithms/top_sort.py
        check_ffi_error(lib.fst_top_sort(self.ptr), "Error during top_sort")
        return self

    def reverse(self) -> "VectorFst":  # vector_fst.py:599 / algorithms/reverse.py:11-31
        out = C.c_void_p()
        check_ffi_error(lib.fst_reverse(self.ptr, C.byref(out)), "Error during reverse")
        return VectorFst(ptr=out)

    def connect(self) -> "VectorFst":
        check_ffi_error(lib.fst_connect(self.ptr), "Error during connect")
        return self

    # ---- comparison / copies / text
    def equals(self, other: "VectorFst") -> bool:
        r = C.c_size_t()
        check_ffi_error(lib.vec_fst_equals(self.ptr, other.ptr, C.byref(r)), "Error checking equality")
        return bool(r.value)

    def __eq__(self, other):
        return self.equals(other)

    def copy(self) -> "VectorFst":
        p = C.c_void_p()
        check_ffi_error(lib.vec_fst_copy(self.ptr, C.byref(p)), "Error copying fst")
        return VectorFst(p)

    def __str__(self):
        s = C.c_char_p()
        check_ffi_error(lib.vec_fst_display(self.ptr, C.byref(s)), "Error displaying fst")
        out = C.string_at(s).decode("utf8")
        lib.rustfst_destroy_string(s)
        return out

    # ---- I/O (OpenFst binary "vector" format)
    @classmethod
    def read(cls, path) -> "VectorFst":
        p = C.c_void_p()
        check_ffi_error(lib.vec_fst_from_path(C.byref(p), str(path).encode("utf-8")), f"Read failed. file: {path}")
        return cls(p)

    def write(self, path):
        check_ffi_error(lib.vec_fst_write_file(self.ptr, str(path).encode("utf-8")), f"Write failed. file: {path}")

    @classmethod
    def from_bytes(cls, data: bytes) -> "VectorFst":
        buf = C.create_string_buffer(data, len(data))
        arr = CArrayU8(C.cast(buf, C.c_void_p), len(data))
        p = C.c_void_p()
        check_ffi_error(lib.vec_fst_from_bytes(C.byref(arr), C.byref(p)), "`from_bytes` failed")
        return cls(p)

    def to_bytes(self) -> bytes:
        arr = C.POINTER(CArrayU8)()
        check_ffi_error(lib.vec_fst_to_bytes(self.ptr, C.byref(arr)), "`to_bytes` failed")
        out = C.string_at(arr.contents.data_ptr, arr.contents.size)
        lib.b200_bytes_destroy(arr)
        return out

    # ---- bulk CSR (b200_ additions)
    @classmethod
    def from_csr(cls, offsets, arcs, finals, start, properties) -> "VectorFst":
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        arcs = np.ascontiguousarray(arcs, dtype=TR_DTYPE)
        finals = np.ascontiguousarray(finals, dtype=np.float32)
        assert len(offsets) == len(finals) + 1 and int(offsets[-1]) == len(arcs)
        p = C.c_void_p()
        check_ffi_error(lib.b200_fst_from_csr(len(finals), offsets.ctypes.data, arcs.ctypes.data, finals.ctypes.data,
                                              -1 if start is None else int(start), int(properties), C.byref(p)),
                        "`from_csr` failed")
        return cls(p)

    def to_csr(self):
        n, a = self.num_states(), self.num_trs_total()
        offsets = np.zeros(n + 1, dtype=np.uint32)
        arcs = np.zeros(a, dtype=TR_DTYPE)
        finals = np.zeros(n, dtype=np.float32)
        start = C.c_int64()
        check_ffi_error(lib.b200_fst_to_csr(self.ptr, offsets.ctypes.data, arcs.ctypes.data, finals.ctypes.data,
                                            C.byref(start)), "`to_csr` failed")
        return offsets, arcs, finals, (None if start.value < 0 else start.value)


class ConstFst:
    """rustfst-python/rustfst/fst/const_fst.py:16-175 — immutable FST in the OpenFst "const" layout (read / write /
    compare / copy / print / from_vector_fst; the generic read accessors of `Fst`).  Algorithms take VectorFst handles,
    as in the reference; `draw` is not part of this build."""

    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            lib.fst_destroy(self.ptr)
        except Exception:
            pass

    @classmethod
    def read(cls, path) -> "ConstFst":  # const_fst.py:90-107
        p = C.c_void_p()
        check_ffi_error(lib.const_fst_from_path(C.byref(p), str(path).encode("utf-8")), f"Read failed. file: {path}")
        return cls(p)

    def write(self, path):  # const_fst.py:127-139
        check_ffi_error(lib.const_fst_write_file(self.ptr, str(path).encode("utf-8")), f"Write failed. file: {path}")

    def equals(self, other: "ConstFst") -> bool:  # const_fst.py:141-154
        r = C.c_size_t()
        check_ffi_error(lib.const_fst_equals(self.ptr, other.ptr, C.byref(r)), "Error checking equality")
        return bool(r.value)

    def __eq__(self, other):
        return isinstance(other, ConstFst) and self.equals(other)

    @classmethod
    def from_vector_fst(cls, fst: "VectorFst") -> "ConstFst":  # const_fst.py:109-125
        p = C.c_void_p()
        check_ffi_error(lib.const_fst_from_vec_fst(fst.ptr, C.byref(p)), "Error converting VectorFst to ConstFst")
        return cls(p)

    def copy(self) -> "ConstFst":  # const_fst.py:156-166
        p = C.c_void_p()
        check_ffi_error(lib.const_fst_copy(self.ptr, C.byref(p)), "Error copying fst")
        return ConstFst(p)

    def __str__(self):  # const_fst.py:168-175
        s = C.c_char_p()
        check_ffi_error(lib.const_fst_display(self.ptr, C.byref(s)), "Error displaying ConstFst")
        out = C.string_at(s).decode("utf8")
        lib.rustfst_destroy_string(s)
        return out

    # generic accessors of the reference's `Fst` base class (fst.py): they work on any handle kind
    start = VectorFst.start
    final = VectorFst.final
    is_final = VectorFst.is_final
    is_start = VectorFst.is_start
    num_trs = VectorFst.num_trs
    trs = VectorFst.trs
    num_trs_total = VectorFst.num_trs_total
    properties = VectorFst.properties

    def num_states(self) -> int:
        n = C.c_uint64()
        check_ffi_error(lib.b200_fst_num_states(self.ptr, C.byref(n)), "Error getting number of states")
        return n.value

    to_csr = VectorFst.to_csr
