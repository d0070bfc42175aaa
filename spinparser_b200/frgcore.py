"""Host-side mirror of SpinParser's flow-core plugin surface on top of ``libpffrg``.

Names, argument meaning and error behaviour follow the reference (file:line relative to the SpinParser tree):

* :class:`FrgCoreFactory` -- ``FrgCoreFactory::newFrgCore(identifier, ...)``, ``src/FrgCoreFactory.cpp:25-51``: the string
  identifiers ``"SU2" | "XYZ" | "TRI"``; any other identifier raises (the reference throws
  ``Exception::Type::InitializationError``).
* :class:`FrgCore` -- ``src/FrgCore.hpp:29-141``: ``computeStep()``, ``finalizeStep(newCutoff)``, ``flowingFunctional()``,
  ``flow()``.
* :class:`EffectiveAction` -- ``src/EffectiveAction.hpp:40-59`` + ``src/SU2/SU2EffectiveAction.hpp:19-233``: ``cutoff``, the
  vertex arrays in the reference's memory layout, ``isDiverged()``.
* :class:`ProblemTables` -- what the hot path reads from ``FrgCommon::lattice()`` / ``FrgCommon::frequency()``
  (``src/FrgCommon.hpp:26-49``), produced by the reference's own ``LatticeModelFactory`` on the host.

All numerics run on the GPU inside ``libpffrg.so``; this file only moves buffers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Mapping, Optional, Sequence

import numpy as np

from . import _capi
from ._capi import PffrgError, check, lib

N_ARRAYS = {"SU2": 2, "XYZ": 4, "TRI": 1}
N_CHANNELS = {"SU2": 2, "XYZ": 4, "TRI": 16}


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


@dataclass
class ProblemTables:
    """Frequency mesh and symmetry-reduced lattice tables (inputs of the hot path, uploaded once)."""

    frequencies: np.ndarray            # [Nw] FrequencyDiscretization::_data (positive half)
    sites_rid: np.ndarray              # [L]   Lattice::getSites()
    sites_perm: np.ndarray             # [L,3]
    inverted_rid: np.ndarray           # [L]   Lattice::getInvertedSites()
    inverted_perm: np.ndarray          # [L,3]
    overlap_offsets: np.ndarray        # [L+1] CSR of Lattice::getOverlap(rid)
    overlap_rid1: np.ndarray
    overlap_rid2: np.ndarray
    overlap_perm1: np.ndarray          # [n,3]
    overlap_perm2: np.ndarray          # [n,3]
    range_fwd_rid: np.ndarray          # [n_range] symmetryTransform(zero, j)
    range_inv_rid: np.ndarray          # [n_range] symmetryTransform(j, zero)
    extra: Dict[str, np.ndarray] = field(default_factory=dict)

    @classmethod
    def from_pfd(cls, d: Mapping[str, np.ndarray]) -> "ProblemTables":
        """Build from a PFD dump written by the host side (keys ``frequency`` and ``lattice/*``)."""
        return cls(
            frequencies=np.ascontiguousarray(d["frequency"], dtype=np.float64),
            sites_rid=_i32(d["lattice/sites_rid"]), sites_perm=_i32(d["lattice/sites_perm"]),
            inverted_rid=_i32(d["lattice/invertedSites_rid"]), inverted_perm=_i32(d["lattice/invertedSites_perm"]),
            overlap_offsets=_i32(d["lattice/overlap_offsets"]),
            overlap_rid1=_i32(d["lattice/overlap_rid1"]), overlap_rid2=_i32(d["lattice/overlap_rid2"]),
            overlap_perm1=_i32(d["lattice/overlap_perm1"]), overlap_perm2=_i32(d["lattice/overlap_perm2"]),
            range_fwd_rid=_i32(d["lattice/range0_fwd_rid"]), range_inv_rid=_i32(d["lattice/range0_inv_rid"]),
        )

    @property
    def n_frequencies(self) -> int:
        return int(self.frequencies.shape[0])

    @property
    def n_sites(self) -> int:
        return int(self.sites_rid.shape[0])

    @property
    def n_items(self) -> int:
        nw = self.n_frequencies
        return nw * nw * (nw + 1) // 2


def make_descriptor(identifier: str, t: ProblemTables, spin_length: float = 0.5, device: int = 0) -> "_capi.Desc":
    """``pffrg_desc`` borrowing the numpy buffers of ``t`` (which must outlive the call that consumes the descriptor)."""
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    return _capi.Desc(
        _capi.ABI_VERSION, _capi.CORE_IDS[identifier],
        t.n_frequencies, t.frequencies.ctypes.data_as(C.POINTER(C.c_double)),
        t.n_sites, ip(t.sites_rid), ip(t.sites_perm), ip(t.inverted_rid), ip(t.inverted_perm),
        ip(t.overlap_offsets), ip(t.overlap_rid1), ip(t.overlap_rid2), ip(t.overlap_perm1), ip(t.overlap_perm2),
        len(t.range_fwd_rid), ip(t.range_fwd_rid), ip(t.range_inv_rid),
        float(spin_length), int(device),
    )


def jit_compile_check(identifier: str, tables: ProblemTables) -> int:
    """Generate + compile the lattice-specialised kernel without a GPU; returns the cubin size in bytes."""
    size = C.c_int64(0)
    check(lib.pffrg_jit_compile_check(C.byref(make_descriptor(identifier, tables)), C.byref(size)))
    return size.value


class EffectiveAction:
    """Host copy of a vertex set in the reference's memory layout (float64 or float32)."""

    def __init__(self, core: str, n_frequencies: int, n_sites: int, dtype=np.float64):
        self.core = core
        self.cutoff = 0.0
        nf = n_frequencies * n_frequencies * (n_frequencies + 1) // 2
        length = nf * n_sites * (16 if core == "TRI" else 1)
        self.v2 = np.zeros(n_frequencies, dtype=dtype)
        self.v4 = [np.zeros(length, dtype=dtype) for _ in range(N_ARRAYS[core])]

    def isDiverged(self) -> bool:
        """NaN scan, ``SU2EffectiveAction::isDiverged`` (src/SU2/SU2EffectiveAction.hpp:212-230)."""
        return bool(np.isnan(self.v2).any() or any(np.isnan(a).any() for a in self.v4))


class FrgCore:
    """One flow core bound to one GPU (``FrgCore``, src/FrgCore.hpp:29-141)."""

    def __init__(self, identifier: str, tables: ProblemTables, options: Optional[Mapping[str, str]] = None, device: int = 0):
        if identifier not in _capi.CORE_IDS:
            raise PffrgError(-1, f"FRG core identifier '{identifier}' is invalid.")
        self.identifier = identifier
        self.tables = tables
        # core options as in src/SU2/SU2FrgCore.cpp:20-29 and src/XYZ/XYZFrgCore.cpp:22-27
        self.spinLength = 0.5
        self.normalization = None
        for key, value in (options or {}).items():
            if key == "spin" and identifier == "SU2":
                self.spinLength = float(value)
            elif key == "normalization":
                self.normalization = float(value)
            else:
                raise PffrgError(-1, f"Unknown spin model option '{key}'.")
        if self.normalization is None:
            self.normalization = 2.0 * self.spinLength if identifier == "SU2" else 1.0

        desc = make_descriptor(identifier, tables, self.spinLength, device)
        handle = C.c_void_p()
        check(lib.pffrg_create(C.byref(desc), C.byref(handle)))
        self._h = handle
        self.n_arrays = lib.pffrg_num_vertex_arrays(self._h)
        self.array_length = int(lib.pffrg_vertex_array_length(self._h))
        self.n_items = int(lib.pffrg_num_items(self._h))
        self.diverged = False
        self._pinned: List[int] = []

    # ---- lifetime ------------------------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            lib.pffrg_destroy(self._h)
            self._h = None
            for ptr in self._pinned:
                lib.pffrg_host_free(ptr)
            self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- multi-GPU ---------------------------------------------------------------------------------------------------
    @staticmethod
    def uniqueId() -> bytes:
        buf = C.create_string_buffer(_capi.UNIQUE_ID_BYTES)
        check(lib.pffrg_comm_unique_id(buf))
        return buf.raw

    def initCommunicator(self, unique_id: bytes, rank: int, n_ranks: int) -> None:
        buf = C.create_string_buffer(unique_id, _capi.UNIQUE_ID_BYTES)
        check(lib.pffrg_comm_init(self._h, buf, rank, n_ranks))

    def itemRange(self):
        b, e = C.c_int64(), C.c_int64()
        check(lib.pffrg_item_range(self._h, C.byref(b), C.byref(e)))
        return b.value, e.value

    def setItemRange(self, begin: int, end: int) -> None:
        check(lib.pffrg_set_item_range(self._h, begin, end))

    # ---- state -----------------------------------------------------------------------------------------------------
    def _ptrs(self, arrays: Sequence[np.ndarray]):
        if len(arrays) != self.n_arrays:
            raise PffrgError(-1, f"expected {self.n_arrays} vertex arrays, got {len(arrays)}")
        for a in arrays:
            if a.size != self.array_length or not a.flags.c_contiguous:
                raise PffrgError(-1, "vertex array has the wrong size or is not contiguous")
        return (C.c_void_p * self.n_arrays)(*[a.ctypes.data for a in arrays])

    @staticmethod
    def _dtype_code(dtype) -> int:
        dtype = np.dtype(dtype)
        if dtype == np.float64:
            return _capi.F64
        if dtype == np.float32:
            return _capi.F32
        raise PffrgError(-1, f"unsupported dtype {dtype}")

    def setState(self, cutoff: float, v2: np.ndarray, v4: Sequence[np.ndarray], sharded: bool = False) -> None:
        """Upload a state. ``sharded=True`` (multi-GPU runs in which every rank holds the same host state, like the MPI ranks of the
        reference): every rank uploads only its share of the rows and the shares are distributed over NVLink
        (``pffrg_set_state_sharded``; collective)."""
        code = self._dtype_code(v2.dtype)
        if any(a.dtype != v2.dtype for a in v4) or v2.size != self.tables.n_frequencies:
            raise PffrgError(-1, "state arrays must share one dtype and match the mesh size")
        fn = lib.pffrg_set_state_sharded if sharded else lib.pffrg_set_state
        check(fn(self._h, float(cutoff), v2.ctypes.data, self._ptrs(v4), code))

    def uploadSlice(self):
        """Work items [begin, end) this rank uploads in ``setState(..., sharded=True)`` (an even split)."""
        b, e = C.c_int64(), C.c_int64()
        check(lib.pffrg_upload_slice(self._h, C.byref(b), C.byref(e)))
        return b.value, e.value

    def setInitialCondition(self, bare_couplings: Sequence[np.ndarray], cutoff: float) -> None:
        """Initial condition of ``SU2EffectiveAction`` (src/SU2/SU2EffectiveAction.hpp:38-60) and the XYZ/TRI equivalents:
        every frequency entry of channel c at representative r is ``bare_couplings[c][r]`` (already divided by the
        normalization and, for XYZ/TRI, multiplied by 1/4); the self energy starts at zero. Built on the device
        (``pffrg_set_initial_condition``)."""
        L, n_ch = self.tables.n_sites, N_CHANNELS[self.identifier]
        # one array per channel, or (TRI) the first row of the reference's single array, [mu][nu][rid]
        flat = np.concatenate([np.ravel(np.asarray(v, dtype=np.float64)) for v in bare_couplings])
        if flat.size != n_ch * L:
            raise PffrgError(-1, f"expected {n_ch} x {L} bare couplings, got {flat.size}")
        bare = np.ascontiguousarray(flat.reshape(n_ch, L))
        check(lib.pffrg_set_initial_condition(self._h, float(cutoff), bare.ctypes.data_as(C.POINTER(C.c_double))))

    def pinnedEffectiveAction(self, dtype=np.float64) -> EffectiveAction:
        """An :class:`EffectiveAction` whose arrays live in page-locked host memory (``pffrg_host_alloc``), for full-speed
        transfers in ``setState`` / ``flowingFunctional(into=...)``. The buffers are released with the core."""
        ea = EffectiveAction.__new__(EffectiveAction)
        ea.core, ea.cutoff = self.identifier, 0.0
        dtype = np.dtype(dtype)

        def pinned(n):
            ptr = lib.pffrg_host_alloc(n * dtype.itemsize)
            if not ptr:
                raise PffrgError(-2, lib.pffrg_last_error().decode())
            self._pinned.append(ptr)
            ctype = C.c_double if dtype == np.float64 else C.c_float
            a = np.ctypeslib.as_array((ctype * n).from_address(ptr))
            a[:] = 0
            return a

        ea.v2 = pinned(self.tables.n_frequencies)
        ea.v4 = [pinned(self.array_length) for _ in range(self.n_arrays)]
        return ea

    def flowingFunctional(self, dtype=np.float64, into: Optional[EffectiveAction] = None, items: Optional[Sequence[int]] = None) -> EffectiveAction:
        """Download the current state (``FrgCore::flowingFunctional``, src/FrgCore.hpp:93-96). ``items=(begin, end)`` with ``into``:
        only the rows of those work items are transferred (``pffrg_get_state_slice``); the other rows of ``into`` are left as they are."""
        if into is not None:
            cutoff = C.c_double()
            if items is not None:
                check(lib.pffrg_get_state_slice(self._h, C.byref(cutoff), into.v2.ctypes.data, self._ptrs(into.v4), self._dtype_code(into.v2.dtype), int(items[0]), int(items[1])))
            else:
                check(lib.pffrg_get_state(self._h, C.byref(cutoff), into.v2.ctypes.data, self._ptrs(into.v4), self._dtype_code(into.v2.dtype)))
            into.cutoff = cutoff.value
            return into
        ea = EffectiveAction(self.identifier, self.tables.n_frequencies, self.tables.n_sites, dtype)
        cutoff = C.c_double()
        check(lib.pffrg_get_state(self._h, C.byref(cutoff), ea.v2.ctypes.data, self._ptrs(ea.v4), self._dtype_code(dtype)))
        ea.cutoff = cutoff.value
        return ea

    def flow(self, dtype=np.float64) -> EffectiveAction:
        """Download the flow of the last ``computeStep`` (``FrgCore::flow``, src/FrgCore.hpp:103-106)."""
        ea = EffectiveAction(self.identifier, self.tables.n_frequencies, self.tables.n_sites, dtype)
        check(lib.pffrg_get_flow(self._h, ea.v2.ctypes.data, self._ptrs(ea.v4), self._dtype_code(dtype)))
        return ea

    # ---- the two virtuals of the reference interface ---------------------------------------------------------------------
    def computeStep(self) -> bool:
        """``FrgCore::computeStep`` (src/SU2/SU2FrgCore.cpp:89-109). Returns the divergence flag the reference obtains from
        ``_flow->isDiverged()`` right afterwards (src/SpinParser.cpp:151)."""
        flag = C.c_int(0)
        check(lib.pffrg_compute_step(self._h, C.byref(flag)))
        self.diverged = bool(flag.value)
        return self.diverged

    def finalizeStep(self, newCutoff: float) -> None:
        """``FrgCore::finalizeStep`` (src/SU2/SU2FrgCore.cpp:111-137)."""
        check(lib.pffrg_finalize_step(self._h, float(newCutoff)))

    def measureCorrelation(self) -> np.ndarray:
        """Static correlations ``chi[c, rid]`` of the current state, computed on the device (replaces the work item of
        ``SU2MeasurementCorrelation::_calculateCorrelation``, src/SU2/SU2MeasurementCorrelation.cpp:77-160, and the XYZ/TRI
        equivalents). Use :func:`correlation_datasets` to arrange them like the reference's ``.obs`` datasets."""
        chi = np.zeros((N_CHANNELS[self.identifier], self.tables.n_sites), dtype=np.float64)
        check(lib.pffrg_measure_correlation(self._h, chi.ctypes.data_as(C.POINTER(C.c_double))))
        return chi

    def synchronize(self) -> None:
        check(lib.pffrg_synchronize(self._h))

    def stats(self) -> Dict[str, float]:
        s = _capi.Stats()
        check(lib.pffrg_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def shapeEnvironment(self) -> Dict[str, str]:
        """Environment settings that make ``pffrg_create`` arrive at this core's launch shape without autotuning: the ranks of a
        sharded run must all use ONE shape (rank 0 tunes, the others adopt its choice), otherwise their summation orders differ in
        the last bit and the sharded state is no longer bit-identical to the single-GPU one."""
        st = self.stats()
        env = {"PFFRG_AUTOTUNE": "0"}
        if st["jit_rpa"] and st["autotuned_shapes"] > 1:
            subs = max(1, st["sub_ctas"])
            env.update({"PFFRG_THREADS": str(st["threads"] // subs), "PFFRG_JIT_NBT": str(st["rpa_batch"] // subs), "PFFRG_JIT_NB": str(st["node_batch"]),
                        "PFFRG_JIT_MINBLOCKS": str(st["min_blocks"])})
            if subs > 1:
                env["PFFRG_SUBCTAS"] = str(subs)
        return env

    @property
    def stream(self) -> int:
        return int(lib.pffrg_stream(self._h) or 0)


class FrgCoreFactory:
    """``FrgCoreFactory::newFrgCore`` (src/FrgCoreFactory.cpp:25-51)."""

    @staticmethod
    def newFrgCore(identifier: str, tables: ProblemTables, options: Optional[Mapping[str, str]] = None, device: int = 0) -> FrgCore:
        return FrgCore(identifier, tables, options, device)


def correlation_datasets(identifier: str, chi: np.ndarray, range_rid: Sequence[np.ndarray], range_perm: Sequence[np.ndarray]) -> Dict[str, np.ndarray]:
    """Arrange ``chi[c, rid]`` like the datasets the reference writes (row b = basis site b, column k = k-th site in range of b):
    ``range_rid[b][k]``, ``range_perm[b][k]`` = ``Lattice::symmetryTransform(b, j_k, x, y, z)`` and the transformed spin components
    (src/SU2/SU2MeasurementCorrelation.cpp:161-176, src/XYZ/XYZMeasurementCorrelation.cpp:206-225, src/TRI/TRIMeasurementCorrelation.cpp:297-330)."""
    rid = np.stack([np.asarray(r) for r in range_rid])
    perm = np.stack([np.asarray(p).reshape(-1, 3) for p in range_perm])
    if identifier == "SU2":
        return {"SU2CorZZ": chi[0][rid], "SU2CorDD": chi[1][rid]}
    if identifier == "XYZ":
        out = {f"XYZCor{n}{n}": chi[perm[..., k], rid] for k, n in enumerate("XYZ")}
        out["XYZCorDD"] = chi[3][rid]
        return out
    out = {"TRICorDD": chi[15][rid]}
    for a, na in enumerate("XYZ"):
        for b, nb in enumerate("XYZ"):
            out[f"TRICor{na}{nb}"] = chi[4 * perm[..., a] + perm[..., b], rid]
    return out


def plan_partition(identifier: str, tables: ProblemTables, cutoff: float, n_ranks: int) -> List[int]:
    """Work-item boundaries of an ``n_ranks``-GPU run at ``cutoff`` (``pffrg_plan_partition``; pure host logic, no GPU)."""
    bounds = (C.c_int64 * (n_ranks + 1))()
    check(lib.pffrg_plan_partition(_capi.CORE_IDS[identifier], tables.n_frequencies, tables.frequencies.ctypes.data_as(C.POINTER(C.c_double)),
                                   tables.n_sites, int(tables.overlap_offsets[-1]), float(cutoff), int(n_ranks), bounds))
    return list(bounds)


def plan_partition_feedback(identifier: str, tables: ProblemTables, cutoff: float, prev_bounds: Sequence[int], prev_ms: Sequence[float]) -> List[int]:
    """Boundaries of the next step given the previous step's boundaries and measured flow-kernel times per rank
    (``pffrg_plan_partition_feedback``; pure host logic, no GPU)."""
    n_ranks = len(prev_ms)
    bounds = (C.c_int64 * (n_ranks + 1))()
    check(lib.pffrg_plan_partition_feedback(_capi.CORE_IDS[identifier], tables.n_frequencies, tables.frequencies.ctypes.data_as(C.POINTER(C.c_double)),
                                            tables.n_sites, int(tables.overlap_offsets[-1]), float(cutoff), n_ranks,
                                            (C.c_int64 * (n_ranks + 1))(*[int(b) for b in prev_bounds]), (C.c_double * n_ranks)(*[float(t) for t in prev_ms]), bounds))
    return list(bounds)


def device_count() -> int:
    return int(lib.pffrg_device_count())
