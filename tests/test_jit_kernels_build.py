"""Device-less build check of the run-time compiled flow kernels of the benchmark workloads (the ahead-of-time library is checked by
__graft_entry__.build()): pffrg_jit_compile_check arrives at the same kernel as pffrg_create would on a B200 and compiles it with
NVRTC for sm_100a. For the warp-specialised kernel the SASS must show what DESIGN.md claims: the register file re-partitioned between
the warp groups (USETMAXREG), the Gram update on the FP64 tensor cores (DMMA.8x8x4), arrive / sync hand-overs on named barriers."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT

WORKLOADS = ["pyrochlore_r8_su2_nw64", "cubic_r7_su2_nw64", "pyrochlore_r10_su2_nw64", "honeycomb_kitaev_r10_xyz_nw64"]


@pytest.mark.parametrize("workload", WORKLOADS)
def test_default_kernel_of_the_workload_compiles_for_sm_100a(workload, monkeypatch, tmp_path):
    from spinparser_b200 import ProblemTables, read_pfd
    from spinparser_b200.frgcore import jit_compile_check
    for k in ("PFFRG_RPA", "PFFRG_SPLIT", "PFFRG_JIT_NBT", "PFFRG_JIT_NB", "PFFRG_THREADS", "PFFRG_PRODUCER", "PFFRG_PERSISTENT"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("PFFRG_JIT_DUMP", str(tmp_path / "kernel"))
    d = read_pfd(os.path.join(ROOT, "bench_data", workload + ".tables.pfd"))
    assert jit_compile_check(bytes(d["core"]).decode(), ProblemTables.from_pfd(d)) > 100000
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump is not installed")
    sass = subprocess.run([cuobjdump, "-sass", "-fun", "pffrg_v4flow_jit", str(tmp_path / "kernel.cubin")], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in subprocess.run([cuobjdump, "-lelf", str(tmp_path / "kernel.cubin")], capture_output=True, text=True).stdout or "SM100" in sass or "sm_100" in sass
    assert sass.count("USETMAXREG") in (2, 3)     # RPA and producer warp groups; gather warp groups unless they keep the launch allocation
    assert sass.count("DMMA.8x8x4") >= 8          # block update of the Gram matrix on the FP64 tensor cores (two channels per tile)
    assert "BAR.ARV" in sass and "BAR.SYNC" in sass  # producer / consumer hand-overs on named barriers
    assert "LDG.E.128" in sass                    # 16-byte row gathers
