"""ctypes loader for librustfst_b200.so — the host-side mirror of rustfst-python/rustfst/ffi_utils.py:16-52.

The library is the product: if it is missing the import fails loudly (build it with `python rustfst_b200/build.py`).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200_LIB") or os.path.join(HERE, "librustfst_b200.so")  # B200_LIB: experimental builds

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"Could not find compiled library {LIB_PATH}; build it with `python rustfst_b200/build.py` "
        "(nvcc, sm_100a). There is no pure-Python or CPU fallback.")

lib = C.CDLL(LIB_PATH)


class CTr(C.Structure):
    _fields_ = [("ilabel", C.c_uint32), ("olabel", C.c_uint32), ("weight", C.c_float), ("nextstate", C.c_uint32)]


class CArrayU8(C.Structure):
    _fields_ = [("data_ptr", C.c_void_p), ("size", C.c_size_t)]


class CIntArray(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint32)), ("size", C.c_size_t)]


class ComposeStats(C.Structure):
    _fields_ = [("states_expanded", C.c_uint64), ("arcs_iterated", C.c_uint64), ("arcs_emitted", C.c_uint64),
                ("waves", C.c_uint64), ("states_out", C.c_uint64), ("arcs_out", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("emit_launches", C.c_uint64),
                ("ms_expand", C.c_float), ("ms_connect", C.c_float), ("ms_emit_kernel", C.c_float),
                ("ms_h2d", C.c_float), ("ms_d2h", C.c_float),
                ("ms_phase_match", C.c_float), ("ms_phase_emit", C.c_float), ("ms_phase_rank", C.c_float),
                ("ms_phase_resolve", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class SsspStats(C.Structure):
    _fields_ = [("arcs_relaxed", C.c_uint64), ("states_settled", C.c_uint64), ("waves", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("relax_launches", C.c_uint64),
                ("path", C.c_int32), ("queue_kind", C.c_int32),
                ("ms_device", C.c_float), ("ms_relax_kernel", C.c_float), ("ms_h2d", C.c_float),
                ("ms_queue_plan_host", C.c_double), ("ms_order_device", C.c_float), ("order_on_device", C.c_int32),
                ("sweep", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def _sig(name, *argtypes):
    fn = getattr(lib, name)
    fn.argtypes = list(argtypes)
    fn.restype = C.c_int
    return fn


_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_sig("rustfst_ffi_get_last_error", C.POINTER(C.c_char_p))
_sig("rustfst_destroy_string", C.c_char_p)
_sig("fst_compose", _P, _P, _PP)
_sig("fst_compose_with_config", _P, _P, _P, _PP)
_sig("fst_compose_config_new", C.c_size_t, C.c_bool, _P, _P, _PP)
_sig("fst_compose_config_destroy", _P)
_sig("fst_matcher_config_new", C.c_size_t, C.c_size_t, CIntArray, _PP)
_sig("fst_matcher_config_destroy", _P)
_sig("fst_shortest_path", _P, _PP)
_sig("fst_shortest_path_with_config", _P, _P, _PP)
_sig("fst_shortest_path_config_new", C.c_float, C.c_size_t, C.c_bool, _PP)
_sig("b200_shortest_path_config_destroy", _P)
_sig("fst_connect", _P)
_sig("fst_reverse", _P, _PP)
_sig("fst_top_sort", _P)
_sig("fst_isomorphic", _P, _P, C.POINTER(C.c_size_t))
_sig("fst_tr_sort", _P, C.c_bool)
_sig("fst_start", _P, C.POINTER(C.c_uint32))
_sig("fst_final_weight", _P, C.c_uint32, C.POINTER(C.c_float))
_sig("fst_num_trs", _P, C.c_uint32, C.POINTER(C.c_size_t))
_sig("fst_get_trs", _P, C.c_uint32, _PP)
_sig("fst_is_final", _P, C.c_uint32, C.POINTER(C.c_size_t))
_sig("fst_is_start", _P, C.c_uint32, C.POINTER(C.c_size_t))
_sig("fst_input_symbols", _P, _PP)
_sig("fst_output_symbols", _P, _PP)
_sig("fst_weight_one", C.POINTER(C.c_float))
_sig("fst_weight_zero", C.POINTER(C.c_float))
_sig("fst_destroy", _P)
_sig("vec_fst_new", _PP)
_sig("vec_fst_set_start", _P, C.c_uint32)
_sig("vec_fst_set_final", _P, C.c_uint32, C.c_float)
_sig("vec_fst_add_state", _P, C.POINTER(C.c_uint32))
_sig("vec_fst_delete_states", _P)
_sig("vec_fst_add_tr", _P, C.c_uint32, _P)
_sig("vec_fst_del_final_weight", _P, C.c_uint32)
_sig("vec_fst_from_path", _PP, C.c_char_p)
_sig("vec_fst_write_file", _P, C.c_char_p)
_sig("vec_fst_num_states", _P, C.POINTER(C.c_size_t))
_sig("vec_fst_equals", _P, _P, C.POINTER(C.c_size_t))
_sig("vec_fst_copy", _P, _PP)
_sig("vec_fst_display", _P, C.POINTER(C.c_char_p))
_sig("vec_fst_to_bytes", _P, C.POINTER(C.POINTER(CArrayU8)))
_sig("vec_fst_from_bytes", C.POINTER(CArrayU8), _PP)
_sig("b200_bytes_destroy", C.POINTER(CArrayU8))
_sig("tr_new", C.c_uint32, C.c_uint32, C.c_float, C.c_uint32, _PP)
_sig("tr_ilabel", _P, C.POINTER(C.c_uint32))
_sig("tr_set_ilabel", _P, C.c_size_t)
_sig("tr_olabel", _P, C.POINTER(C.c_uint32))
_sig("tr_set_olabel", _P, C.c_size_t)
_sig("tr_weight", _P, C.POINTER(C.c_float))
_sig("tr_set_weight", _P, C.c_float)
_sig("tr_next_state", _P, C.POINTER(C.c_uint32))
_sig("tr_set_next_state", _P, C.c_size_t)
_sig("tr_delete", _P)
_sig("trs_vec_new", _PP)
_sig("trs_vec_remove", _P, C.c_size_t, _PP)
_sig("trs_vec_push", _P, _P)
_sig("trs_vec_shallow_clone", _P, _PP)
_sig("trs_vec_len", _P, C.POINTER(C.c_size_t))
_sig("trs_vec_display", _P, C.POINTER(C.c_char_p))
_sig("trs_vec_delete", _P)
_sig("trs_iterator_new", _P, C.c_uint32, _PP)
_sig("trs_iterator_next", _P, _PP)
_sig("trs_iterator_done", _P, C.POINTER(C.c_size_t))
_sig("trs_iterator_reset", _P)
_sig("trs_iterator_destroy", _P)
_sig("mut_trs_iterator_new", _P, C.c_uint32, _PP)
_sig("mut_trs_iterator_next", _P)
_sig("mut_trs_iterator_value", _P, _PP)
_sig("mut_trs_iterator_set_value", _P, _P)
_sig("mut_trs_iterator_done", _P, C.POINTER(C.c_size_t))
_sig("mut_trs_iterator_reset", _P)
_sig("mut_trs_iterator_destroy", _P)
_sig("state_iterator_new", _P, _PP)
_sig("state_iterator_next", _P, C.POINTER(C.c_uint32))
_sig("state_iterator_done", _P, C.POINTER(C.c_size_t))
_sig("state_iterator_destroy", _P)
_sig("b200_fst_properties", _P, C.POINTER(C.c_uint64))
_sig("b200_fst_set_properties", _P, C.c_uint64)
_sig("b200_fst_from_csr", C.c_uint64, _P, _P, _P, C.c_int64, C.c_uint64, _PP)
_sig("b200_fst_num_states", _P, C.POINTER(C.c_uint64))
_sig("const_fst_from_path", _PP, C.c_char_p)
_sig("const_fst_write_file", _P, C.c_char_p)
_sig("const_fst_equals", _P, _P, C.POINTER(C.c_size_t))
_sig("const_fst_copy", _P, _PP)
_sig("const_fst_display", _P, C.POINTER(C.c_char_p))
_sig("const_fst_from_vec_fst", _P, _PP)
_sig("b200_fst_compute_properties", _P, C.POINTER(C.c_uint64))
_sig("b200_fst_num_trs_total", _P, C.POINTER(C.c_uint64))
_sig("b200_fst_to_csr", _P, _P, _P, _P, C.POINTER(C.c_int64))
_sig("b200_compose_with_stats", _P, _P, _P, _PP, C.POINTER(ComposeStats))
_sig("b200_shortest_path_with_stats", _P, _P, _PP, C.POINTER(SsspStats), C.c_bool)
_sig("b200_device_fst_upload", _P, _PP)
_sig("b200_device_fst_download", _P, _PP)
_sig("b200_device_fst_info", _P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64))
_sig("b200_device_fst_destroy", _P)
_sig("b200_device_compose", _P, _P, _P, _PP, C.POINTER(ComposeStats))
_sig("b200_device_shortest_path", _P, _P, _PP, C.POINTER(SsspStats), C.c_bool)
_sig("b200_device_shortest_path_with_config", _P, _P, _P, _PP, C.POINTER(SsspStats), C.c_bool)
_sig("b200_compose_batch", C.POINTER(C.c_void_p), C.c_size_t, _P, _P, C.POINTER(C.c_void_p), C.POINTER(ComposeStats))
_sig("b200_shortest_path_queue_plan", _P, C.POINTER(C.c_int32), _P, _P, C.POINTER(C.c_uint32))
_sig("b200_device_isomorphic", _P, _P, C.POINTER(C.c_int32))
_sig("b200_compose_batch_packed", _P, C.c_size_t, _P, _P, _P, _PP, C.POINTER(ComposeStats))
_sig("b200_packed_batch_info", _P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64))
_sig("b200_packed_batch_get", _P, C.c_size_t, _PP)
_sig("b200_packed_batch_serialize", _P, _P, C.c_size_t)
_sig("b200_packed_batch_deserialize", _P, C.c_size_t, _PP)
_sig("b200_packed_batch_destroy", _P)
_sig("b200_dag_top_order_device", _P, _P, C.POINTER(C.c_int32), C.POINTER(C.c_float))
_sig("b200_set_device", C.c_int)
_sig("b200_device_count", C.POINTER(C.c_int))
_sig("b200_device_synchronize")
lib.b200_version.restype = C.c_char_p


def check_ffi_error(exit_code, error_context_msg):
    """rustfst-python/rustfst/ffi_utils.py:45-52"""
    if exit_code != 0:
        ptr = C.c_char_p()
        if lib.rustfst_ffi_get_last_error(C.byref(ptr)) == 0:
            msg = C.string_at(ptr).decode("utf8")
            lib.rustfst_destroy_string(ptr)
        else:
            msg = "see stderr"
        raise ValueError(f"{error_context_msg}: {msg}")


def device_count():
    n = C.c_int()
    lib.b200_device_count(C.byref(n))
    return n.value
