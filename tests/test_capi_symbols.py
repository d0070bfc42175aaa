"""The C-ABI library loads and exports every symbol include/pffrg.h declares; without a GPU it refuses to compute.

No compute calls are made here (this file runs on the CPU-only build container as well as on the GPU box).
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pffrg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pffrg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from spinparser_b200 import _capi
    declared = _declared_symbols()
    assert len(declared) >= 20
    missing = [name for name in declared if not hasattr(_capi.lib, name)]
    assert not missing, f"libpffrg.so lacks {missing}"
    assert sorted(_capi.SYMBOLS) == declared, "spinparser_b200/_capi.py and include/pffrg.h disagree on the exported surface"
    assert _capi.lib.pffrg_abi_version() == _capi.ABI_VERSION


def test_descriptor_layout_matches_header():
    """Field order of the ctypes mirror == field order of `struct pffrg_desc` / `struct pffrg_stats` in the header."""
    from spinparser_b200 import _capi
    text = open(os.path.join(ROOT, "include", "pffrg.h")).read()
    for struct, mirror in (("pffrg_desc", _capi.Desc), ("pffrg_stats", _capi.Stats)):
        body = re.search(r"typedef struct %s\s*\{(.*?)\}\s*%s;" % (struct, struct), text, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = [re.search(r"(\w+)\s*$", decl.strip()).group(1) for decl in body.split(";") if decl.strip()]
        assert fields == [name for name, _ in mirror._fields_], struct


def test_invalid_arguments_are_rejected_without_touching_a_gpu():
    from spinparser_b200 import _capi
    lib = _capi.lib
    handle = C.c_void_p()
    assert lib.pffrg_create(None, C.byref(handle)) == -1  # PFFRG_ERR_ARGUMENT
    assert b"null" in lib.pffrg_last_error()
    desc = _capi.Desc()
    desc.abi_version = 999
    assert lib.pffrg_create(C.byref(desc), C.byref(handle)) == -1
    assert b"ABI version" in lib.pffrg_last_error()
    assert lib.pffrg_compute_step(None, None) == -1
    assert lib.pffrg_destroy(None) == 0


def test_unknown_core_identifier_raises_like_the_reference_factory():
    # FrgCoreFactory::newFrgCore throws InitializationError for unknown identifiers (src/FrgCoreFactory.cpp:48-50)
    from spinparser_b200 import FrgCoreFactory, PffrgError, ProblemTables
    with pytest.raises(PffrgError, match="invalid"):
        FrgCoreFactory.newFrgCore("SU3", ProblemTables.from_pfd(golden("su2_square_r3_nw10")))


def test_no_cpu_fallback():
    """On a machine without a GPU creating a core fails loudly (PFFRG_ERR_CUDA); on the GPU box this test is a no-op."""
    from spinparser_b200 import FrgCoreFactory, PffrgError, ProblemTables
    from spinparser_b200.frgcore import device_count
    if device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(PffrgError) as err:
        FrgCoreFactory.newFrgCore("SU2", ProblemTables.from_pfd(golden("su2_square_r3_nw10")), {"spin": "0.5"})
    assert err.value.code == -2 and "no CPU fallback" in str(err.value)


@pytest.mark.parametrize("case,core", [("su2_square_r3_nw10", "SU2"), ("xyz_honeycomb_kitaev_r3_nw10", "XYZ")])
def test_lattice_specialised_kernel_compiles_without_a_gpu(case, core):
    """NVRTC generates sm_100a code for the run-time specialised flow kernel on the build machine (no device needed)."""
    from spinparser_b200 import ProblemTables
    from spinparser_b200.frgcore import jit_compile_check
    assert jit_compile_check(core, ProblemTables.from_pfd(golden(case))) > 10000


def test_pfd_round_trip(tmp_path):
    from spinparser_b200 import read_pfd, write_pfd
    arrays = {"a/b": np.arange(6, dtype=np.float64).reshape(2, 3), "i": np.array([1, -2], dtype=np.int32), "s": np.float32(2.5)}
    write_pfd(str(tmp_path / "x.pfd"), arrays)
    back = read_pfd(str(tmp_path / "x.pfd"))
    assert list(back) == list(arrays)
    for k in arrays:
        assert back[k].dtype == np.asarray(arrays[k]).dtype and np.array_equal(back[k], arrays[k])


def test_unknown_core_options_raise_like_the_reference_constructors():
    # "Unknown spin model option" -- src/SU2/SU2FrgCore.cpp:27, src/XYZ/XYZFrgCore.cpp:25 (XYZ/TRI do not know `spin`)
    from spinparser_b200 import FrgCoreFactory, PffrgError, ProblemTables
    tables = ProblemTables.from_pfd(golden("su2_square_r3_nw10"))
    with pytest.raises(PffrgError, match="Unknown spin model option 'foo'"):
        FrgCoreFactory.newFrgCore("SU2", tables, {"foo": "1"})
    with pytest.raises(PffrgError, match="Unknown spin model option 'spin'"):
        FrgCoreFactory.newFrgCore("XYZ", tables, {"spin": "0.5"})
