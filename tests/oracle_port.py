"""ctypes binding of oracle/liboracle.so (the plain-C restatement of the reference hot path).

TEST INFRASTRUCTURE: imported only by tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, List, Optional, Sequence

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
CORES = {"SU2": 0, "XYZ": 1, "TRI": 2}
N_ARRAYS = {"SU2": 2, "XYZ": 4, "TRI": 1}
N_CHANNELS = {"SU2": 2, "XYZ": 4, "TRI": 16}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class _Problem(C.Structure):
    _fields_ = [
        ("core", C.c_int), ("nw", C.c_int), ("mesh", _dp), ("L", C.c_int),
        ("sites_rid", _ip), ("sites_perm", _ip), ("inv_rid", _ip), ("inv_perm", _ip),
        ("ov_off", _ip), ("ov_rid1", _ip), ("ov_rid2", _ip), ("ov_perm1", _ip), ("ov_perm2", _ip),
        ("nrange", C.c_int), ("rng_fwd_rid", _ip), ("rng_inv_rid", _ip),
        ("spin_length", C.c_double),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", ORACLE_DIR, "port"], check=True, capture_output=True)
        _lib = C.CDLL(path)
        _lib.pfo_v2_flow.argtypes = [C.POINTER(_Problem), C.c_double, _dp, C.POINTER(_dp), _dp]
        _lib.pfo_v4_flow.argtypes = [C.POINTER(_Problem), C.c_double, _dp, _dp, C.POINTER(_dp), _ip, C.c_int, C.POINTER(_dp)]
        _lib.pfo_euler.argtypes = [_dp, _dp, C.c_long, C.c_double, C.c_double]
        _lib.pfo_node_count.argtypes = [C.POINTER(_Problem), C.c_double, C.c_double]
        _lib.pfo_node_count.restype = C.c_int
        for name in ("pfo_mesh_lesser", "pfo_mesh_greater", "pfo_mesh_offset"):
            getattr(_lib, name).argtypes = [C.c_int, _dp, C.c_double]
            getattr(_lib, name).restype = C.c_int
        _lib.pfo_mesh_interpolate.argtypes = [C.c_int, _dp, C.c_double, _ip, _ip, _dp]
        _lib.pfo_mesh_value.argtypes = [C.c_int, _dp, C.c_int]
        _lib.pfo_mesh_value.restype = C.c_double
        fn = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)
        _lib.scalar_fn = fn
        _lib.pfo_integrate_left.argtypes = [C.c_int, _dp, C.c_double, C.c_int, fn, C.c_void_p]
        _lib.pfo_integrate_right.argtypes = [C.c_int, _dp, C.c_int, C.c_double, fn, C.c_void_p]
        _lib.pfo_integrate_both.argtypes = [C.c_int, _dp, C.c_double, C.c_double, fn, C.c_void_p]
        for name in ("pfo_integrate_left", "pfo_integrate_right", "pfo_integrate_both"):
            getattr(_lib, name).restype = C.c_double
    return _lib


def _d(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def _i(a: np.ndarray):
    return a.ctypes.data_as(_ip)


class OraclePort:
    """The reference algorithm on one problem (lattice tables + mesh), reference memory layout throughout."""

    def __init__(self, tables: Dict[str, np.ndarray]):
        t = tables
        self.core = t["core"] if isinstance(t["core"], str) else bytes(t["core"].astype(np.uint8)).decode()
        self.mesh = np.ascontiguousarray(t["frequency"], dtype=np.float64)
        self.nw = len(self.mesh)
        self.L = int(t["lattice/size"])
        self.nf = self.nw * self.nw * (self.nw + 1) // 2
        i32 = lambda k: np.ascontiguousarray(t[k], dtype=np.int32)
        self._keep = dict(
            sites_rid=i32("lattice/sites_rid"), sites_perm=i32("lattice/sites_perm"),
            inv_rid=i32("lattice/invertedSites_rid"), inv_perm=i32("lattice/invertedSites_perm"),
            ov_off=i32("lattice/overlap_offsets"), ov_rid1=i32("lattice/overlap_rid1"), ov_rid2=i32("lattice/overlap_rid2"),
            ov_perm1=i32("lattice/overlap_perm1"), ov_perm2=i32("lattice/overlap_perm2"),
            rng_fwd_rid=i32("lattice/range0_fwd_rid"), rng_inv_rid=i32("lattice/range0_inv_rid"),
        )
        k = self._keep
        self.spin_length = float(t["spinLength"]) if "spinLength" in t else 0.5
        self._p = _Problem(
            CORES[self.core], self.nw, _d(self.mesh), self.L,
            _i(k["sites_rid"]), _i(k["sites_perm"]), _i(k["inv_rid"]), _i(k["inv_perm"]),
            _i(k["ov_off"]), _i(k["ov_rid1"]), _i(k["ov_rid2"]), _i(k["ov_perm1"]), _i(k["ov_perm2"]),
            len(k["rng_fwd_rid"]), _i(k["rng_fwd_rid"]), _i(k["rng_inv_rid"]), self.spin_length,
        )

    @property
    def n_arrays(self) -> int:
        return N_ARRAYS[self.core]

    @property
    def array_len(self) -> int:
        return self.nf * self.L * (16 if self.core == "TRI" else 1)

    def _ptrs(self, arrays: Sequence[np.ndarray]):
        assert len(arrays) == self.n_arrays
        for a in arrays:
            assert a.dtype == np.float64 and a.flags.c_contiguous and a.size == self.array_len
        return (_dp * len(arrays))(*[_d(a) for a in arrays])

    def v2_flow(self, cutoff: float, v2: np.ndarray, v4: Sequence[np.ndarray]) -> np.ndarray:
        out = np.zeros(self.nw)
        lib().pfo_v2_flow(C.byref(self._p), cutoff, _d(v2), self._ptrs(v4), _d(out))
        return out

    def v4_flow(self, cutoff: float, v2: np.ndarray, v2flow: np.ndarray, v4: Sequence[np.ndarray],
                items: Optional[np.ndarray] = None) -> List[np.ndarray]:
        """Flow of the listed work items (all if None); untouched entries of the returned arrays are zero."""
        out = [np.zeros(self.array_len) for _ in range(self.n_arrays)]
        if items is None:
            ip, n = None, self.nf
        else:
            items = np.ascontiguousarray(items, dtype=np.int32)
            ip, n = _i(items), len(items)
        lib().pfo_v4_flow(C.byref(self._p), cutoff, _d(v2), _d(v2flow), self._ptrs(v4), ip, n, self._ptrs(out))
        return out

    def node_count(self, cutoff: float, x: float) -> int:
        return lib().pfo_node_count(C.byref(self._p), cutoff, x)
