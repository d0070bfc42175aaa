"""Parity of the CUDA path AT THE BENCHMARKED SIZES (Nw = 64 compile-time constants, the launch shapes bench.py runs, the bucketed
mesh search on a 64-point exponential mesh, L up to 103): every bench_data workload is brought to a physical state by running the
flow on the GPU from the bare couplings, the state is downloaded, and the GPU's vertex flow of the next step is compared on a
strided sample of work items (stride coprime to Nw, so every t index occurs) with the FP64 build of the UNMODIFIED reference
(oracle/_ref/oracle64 --mode time: the per-item calculators of src/SU2/SU2FrgCore.cpp:139-431 and the XYZ / TRI equivalents) started
from that same state. Where the reference binary is absent (it is a build output that travels with the repo snapshot) the plain-C
restatement oracle/liboracle.so, pinned bit-for-bit against it on the small fixtures, stands in.

Bar: |d_i| <= 1e-10 |x_i| + 1e-12 max|x| per channel array (conftest.assert_parity; north star: 1e-10 relative per entry).
"""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT, assert_parity

pytestmark = pytest.mark.gpu

ORACLE64 = os.path.join(ROOT, "oracle", "_ref", "oracle64")
STRIDE = 131  # coprime to Nw = 64 and 32: the sample visits every t index and every part of the (s, u) triangle

# workload -> (cutoff step the physical state is taken at, further cutoff indices evaluated ON that state, kernel variants)
# The TRI workload costs ~1 s per step at step 211, so its physical state is taken at step 120 and the benchmark cutoff (index 211,
# ~62 quadrature nodes per item) is evaluated on that state as well.
WORKLOADS = {
    "square_r4_su2_nw32": (100, [], [{}, {"PFFRG_AUTOTUNE": "1"}, {"PFFRG_RPA": "gram"}]),
    # default: the warp-specialised Gram kernel with persistent CTAs; then the lattice-specialised straight-line code (first shape and autotuned), the unsplit Gram kernel
    "cubic_r7_su2_nw64": (211, [], [{}, {"PFFRG_RPA": "code"}, {"PFFRG_RPA": "code", "PFFRG_AUTOTUNE": "1"}, {"PFFRG_RPA": "gram"}, {"PFFRG_PERSISTENT": "0"}]),
    "honeycomb_kitaev_r7_xyz_nw64": (211, [], [{}, {"PFFRG_AUTOTUNE": "1"}, {"PFFRG_RPA": "gram"}]),
    # large-range XYZ: the Gram form in the warp-specialised kernel by default (3 323 merged overlap terms), then the straight-line code
    "honeycomb_kitaev_r10_xyz_nw64": (211, [], [{}, {"PFFRG_RPA": "code"}]),
    # default: the warp-specialised Gram kernel (gather / RPA / producer warp groups); then with several RPA rounds and small batches,
    # the unsplit Gram kernel without and with a producer warp
    "pyrochlore_r8_su2_nw64": (211, [], [{}, {"PFFRG_JIT_NBT": "16", "PFFRG_JIT_NB": "8"}, {"PFFRG_SPLIT": "0"}, {"PFFRG_SPLIT": "0", "PFFRG_PRODUCER": "1"}]),
    "pyrochlore_r10_su2_nw64": (211, [], [{}, {"PFFRG_SPLIT": "0"}]),
    "kagome_dm_r7_tri_nw64": (120, [211], [{}, {"PFFRG_RPA": "gram"}]),
}


def _tables(workload):
    from spinparser_b200 import read_pfd
    return read_pfd(os.path.join(ROOT, "bench_data", workload + ".tables.pfd"))


def _core(d, env, monkeypatch):
    from spinparser_b200 import FrgCoreFactory, ProblemTables
    for k in ("PFFRG_AUTOTUNE", "PFFRG_RPA", "PFFRG_JIT_NBT", "PFFRG_JIT_NB", "PFFRG_PRODUCER", "PFFRG_SPLIT", "PFFRG_PERSISTENT"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    name = bytes(d["core"]).decode()
    opts = {"spin": str(float(d["spinLength"]))} if name == "SU2" else {}
    return name, FrgCoreFactory.newFrgCore(name, ProblemTables.from_pfd(d), opts)


def reference_flow_rows(workload, d, step, v2, v4, items):
    """Flow of the listed work items at cutoff index `step` from the given state: (v2 flow, [rows per channel array], what ran)."""
    from spinparser_b200.pfd import read_pfd, write_pfd
    core = bytes(d["core"]).decode()
    nw, L = len(d["frequency"]), int(d["lattice/size"])
    per = L * (16 if core == "TRI" else 1)
    if os.path.exists(ORACLE64):
        with tempfile.TemporaryDirectory() as tmp:
            state, out = os.path.join(tmp, "state.pfd"), os.path.join(tmp, "out.pfd")
            write_pfd(state, {"v2": np.asarray(v2, dtype=np.float64), **{f"v4_{c}": np.asarray(a, dtype=np.float64) for c, a in enumerate(v4)}})
            cmd = [ORACLE64, "-r", os.path.join(ROOT, "oracle", "res"), os.path.join(ROOT, "bench_data", "tasks", workload + ".xml"), "--out", out,
                   "--mode", "time", "--time-compact", "--load-state", state, "--start-step", str(step), "--time-stride", str(STRIDE), "--time-offset", str(int(items[0])),
                   "--time-repeat", "1", "--time-warmup", "0", "--no-lattice", "--threads", str(os.cpu_count() or 1)]
            subprocess.run(cmd, check=True, capture_output=True, text=True)
            r = read_pfd(out)
        assert np.array_equal(r["time/itemIds"], items)
        return r["time/flow/v2"], [r[f"time/flowItems/v4_{c}"].reshape(len(items), per) for c in range(len(v4))], "oracle64 (unmodified reference, FP64 build)"
    from oracle_port import OraclePort
    port = OraclePort(d)
    cutoff = float(d["cutoff"][step])
    f2 = port.v2_flow(cutoff, v2, v4)
    full = port.v4_flow(cutoff, v2, f2, v4, items)
    return f2, [a.reshape(-1, per)[items] for a in full], "oracle port (plain-C restatement)"


@pytest.mark.parametrize("workload", list(WORKLOADS))
def test_flow_at_benchmark_size_matches_reference(workload, monkeypatch):
    state_step, extra_steps, variants = WORKLOADS[workload]
    d = _tables(workload)
    cutoffs = [float(x) for x in d["cutoff"]]
    nw, L = len(d["frequency"]), int(d["lattice/size"])
    nf = nw * nw * (nw + 1) // 2
    items = np.arange(5, nf, STRIDE, dtype=np.int32)
    assert len(set(int(i) % nw for i in items)) == nw  # every t index is sampled

    # physical state: the flow from the bare couplings on the GPU (default kernel variant)
    name, core = _core(d, variants[0], monkeypatch)
    per = L * (16 if name == "TRI" else 1)
    core.setInitialCondition(list(d["bare"]), cutoffs[0])
    for step in range(state_step):
        assert not core.computeStep(), f"diverged at step {step}"
        core.finalizeStep(cutoffs[step + 1])
    state = core.flowingFunctional()
    assert state.cutoff == cutoffs[state_step] and not state.isDiverged()
    report = {}
    current = variants[0]
    for step in [state_step] + extra_steps:
        want2, want4, what = reference_flow_rows(workload, d, step, state.v2, state.v4, items)
        for k, env in enumerate(variants):
            if k > 0 or step != state_step:
                if env != current:
                    core.close()
                    name, core = _core(d, env, monkeypatch)
                    current = env
                core.setState(cutoffs[step], state.v2, state.v4)
            assert not core.computeStep()
            flow = core.flow()
            tag = f"{workload} step {step} {env or 'default'} vs {what}"
            assert_parity(flow.v2, want2, tag + ": v2 flow")
            worst = 0.0
            for c in range(core.n_arrays):
                got = flow.v4[c].reshape(-1, per)[items]
                assert np.abs(want4[c]).max() > 0
                assert_parity(got, want4[c], tag + f": v4 flow channel array {c}")
                worst = max(worst, float(np.abs(got - want4[c]).max() / np.abs(want4[c]).max()))
            st = core.stats()
            report[f"step{step}/{json.dumps(env, sort_keys=True)}"] = {"max_normwise": worst, "threads": st["threads"], "rpa_batch": st["rpa_batch"], "node_batch": st["node_batch"], "ms_v4_flow": st["ms_v4_flow"]}
    core.close()
    print(f"\n[parity at size] {workload}: {len(items)} of {nf} items, {json.dumps(report)}")
