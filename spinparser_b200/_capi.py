"""ctypes binding of ``libpffrg.so`` (the C ABI declared in ``include/pffrg.h``).

This module is plumbing only: it mirrors the header one to one and raises :class:`PffrgError` with the library's
own error text on any non-zero status. There is no fallback path: if the CUDA library cannot be loaded, importing
this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpffrg.so")

ABI_VERSION = 3
CORE_IDS = {"SU2": 0, "XYZ": 1, "TRI": 2}
F32, F64 = 0, 1
UNIQUE_ID_BYTES = 128

# every symbol include/pffrg.h declares (checked by tests/test_capi_symbols.py against the header text)
SYMBOLS = [
    "pffrg_abi_version", "pffrg_last_error", "pffrg_device_count", "pffrg_create", "pffrg_destroy",
    "pffrg_num_vertex_arrays", "pffrg_vertex_array_length", "pffrg_num_items", "pffrg_comm_unique_id",
    "pffrg_comm_init", "pffrg_item_range", "pffrg_set_state", "pffrg_set_initial_condition", "pffrg_get_state", "pffrg_get_flow",
    "pffrg_compute_step", "pffrg_finalize_step", "pffrg_synchronize", "pffrg_num_channels", "pffrg_measure_correlation", "pffrg_set_item_range", "pffrg_get_stats",
    "pffrg_stream", "pffrg_fp64_peak", "pffrg_dmma_peak", "pffrg_host_alloc", "pffrg_host_free", "pffrg_host_register", "pffrg_host_unregister", "pffrg_jit_compile_check", "pffrg_tri_terms", "pffrg_gram_tables", "pffrg_trigram_tables", "pffrg_site_order", "pffrg_upload_slice", "pffrg_set_state_sharded", "pffrg_get_state_slice", "pffrg_plan_partition", "pffrg_plan_partition_feedback",
]


class PffrgError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libpffrg error {code}: {message}")
        self.code = code


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class Desc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("core", C.c_int32),
        ("n_frequencies", C.c_int32), ("frequencies", _dp),
        ("n_sites", C.c_int32),
        ("sites_rid", _ip), ("sites_perm", _ip), ("inverted_rid", _ip), ("inverted_perm", _ip),
        ("overlap_offsets", _ip), ("overlap_rid1", _ip), ("overlap_rid2", _ip), ("overlap_perm1", _ip), ("overlap_perm2", _ip),
        ("n_range", C.c_int32), ("range_fwd_rid", _ip), ("range_inv_rid", _ip),
        ("spin_length", C.c_double), ("device", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("ms_v2_flow", C.c_double), ("ms_node_table", C.c_double), ("ms_v4_flow", C.c_double),
        ("ms_finalize", C.c_double), ("ms_exchange", C.c_double),
        ("kernel_evals", C.c_int64), ("kernel_evals_t", C.c_int64), ("items", C.c_int64),
        ("alg_bytes", C.c_double), ("alg_flops", C.c_double), ("launches", C.c_int32), ("jit_rpa", C.c_int32), ("jit_compile_ms", C.c_double),
        ("threads", C.c_int32), ("smem_bytes", C.c_int32), ("node_batch", C.c_int32), ("rpa_batch", C.c_int32), ("rpa_warps", C.c_int32), ("min_blocks", C.c_int32), ("autotuned_shapes", C.c_int32), ("sub_ctas", C.c_int32),
        ("gram_rows", C.c_int32), ("rpa_terms_merged", C.c_int32), ("exec_flops", C.c_double),
        ("gather_threads", C.c_int32), ("producer_warps", C.c_int32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C spinparser_b200/csrc` (or __graft_entry__.build()). "
            "spinparser_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.pffrg_abi_version.restype = C.c_int
    lib.pffrg_last_error.restype = C.c_char_p
    lib.pffrg_device_count.restype = C.c_int
    lib.pffrg_create.argtypes = [C.POINTER(Desc), C.POINTER(vp)]
    lib.pffrg_destroy.argtypes = [vp]
    lib.pffrg_num_vertex_arrays.argtypes = [vp]
    lib.pffrg_vertex_array_length.argtypes = [vp]
    lib.pffrg_vertex_array_length.restype = C.c_int64
    lib.pffrg_num_items.argtypes = [vp]
    lib.pffrg_num_items.restype = C.c_int64
    lib.pffrg_comm_unique_id.argtypes = [vp]
    lib.pffrg_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.pffrg_item_range.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.pffrg_set_item_range.argtypes = [vp, C.c_int64, C.c_int64]
    lib.pffrg_set_state.argtypes = [vp, C.c_double, vp, C.POINTER(vp), C.c_int]
    lib.pffrg_set_initial_condition.argtypes = [vp, C.c_double, _dp]
    lib.pffrg_get_state.argtypes = [vp, _dp, vp, C.POINTER(vp), C.c_int]
    lib.pffrg_get_flow.argtypes = [vp, vp, C.POINTER(vp), C.c_int]
    lib.pffrg_compute_step.argtypes = [vp, C.POINTER(C.c_int)]
    lib.pffrg_finalize_step.argtypes = [vp, C.c_double]
    lib.pffrg_synchronize.argtypes = [vp]
    lib.pffrg_num_channels.argtypes = [vp]
    lib.pffrg_measure_correlation.argtypes = [vp, _dp]
    lib.pffrg_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.pffrg_stream.argtypes = [vp]
    lib.pffrg_stream.restype = vp
    lib.pffrg_fp64_peak.argtypes = [C.c_int]
    lib.pffrg_fp64_peak.restype = C.c_double
    lib.pffrg_dmma_peak.argtypes = [C.c_int]
    lib.pffrg_dmma_peak.restype = C.c_double
    lib.pffrg_host_alloc.argtypes = [C.c_size_t]
    lib.pffrg_host_alloc.restype = vp
    lib.pffrg_host_free.argtypes = [vp]
    lib.pffrg_host_register.argtypes = [vp, C.c_size_t]
    lib.pffrg_host_unregister.argtypes = [vp]
    lib.pffrg_jit_compile_check.argtypes = [C.POINTER(Desc), C.POINTER(C.c_int64)]
    lib.pffrg_tri_terms.argtypes = [C.c_int, _ip, C.c_int]
    lib.pffrg_upload_slice.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.pffrg_set_state_sharded.argtypes = [vp, C.c_double, vp, C.POINTER(vp), C.c_int]
    lib.pffrg_get_state_slice.argtypes = [vp, C.POINTER(C.c_double), vp, C.POINTER(vp), C.c_int, C.c_int64, C.c_int64]
    lib.pffrg_trigram_tables.argtypes = [C.POINTER(Desc), C.c_int, C.c_int, C.POINTER(C.c_uint16), C.c_int, C.POINTER(C.c_uint32), C.c_int, _ip, C.c_int, _ip]
    lib.pffrg_site_order.argtypes = [C.POINTER(Desc), _ip]
    lib.pffrg_gram_tables.argtypes = [C.POINTER(Desc), C.c_int, C.c_int, C.POINTER(C.c_uint32), C.c_int, _ip, _dp]
    lib.pffrg_plan_partition.argtypes = [C.c_int, C.c_int, _dp, C.c_int, C.c_int64, C.c_double, C.c_int, C.POINTER(C.c_int64)]
    lib.pffrg_plan_partition_feedback.argtypes = [C.c_int, C.c_int, _dp, C.c_int, C.c_int64, C.c_double, C.c_int, C.POINTER(C.c_int64), _dp, C.POINTER(C.c_int64)]
    if lib.pffrg_abi_version() != ABI_VERSION:
        raise ImportError(f"libpffrg ABI {lib.pffrg_abi_version()} != binding ABI {ABI_VERSION}")
    return lib


lib = _load()


def check(code: int) -> int:
    if code < 0:
        raise PffrgError(code, lib.pffrg_last_error().decode())
    return code
