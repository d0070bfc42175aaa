"""The reference's own golden vectors, test/scripted/assets/test_reference{1,2,3}.ref (hyperkagome Heisenberg on all three cores,
hyperhoneycomb Kitaev on XYZ / TRI, honeycomb Kitaev-Gamma on TRI; 34 cutoffs each; tolerance 1e-5, test/scripted/assets/test_eval.py:14).

* CPU, where the reference tree is mounted: the as-shipped FP32 build of the UNMODIFIED reference (oracle/_ref/oracle32) reproduces all 544
  golden datasets BIT FOR BIT -- this pins the oracle --, and the committed fixture tests/golden/reference_goldens.pfd equals the HDF5
  files parsed by the dependency-free reader tests/hdf5_v0.py.
* GPU: the complete flow on the device (initial condition, 33 Euler steps, correlation measurement at every cutoff through the C ABI)
  from the fixture's problem tables lands within the reference's own tolerance of the golden datasets, and the cross-core identities of
  test_reference1.sh:70-78 / test_reference2.sh hold.
"""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

REFERENCE = "/root/reference"
FIXTURE = os.path.join(GOLDEN, "reference_goldens.pfd")
RUNS = [("ref1", "SU2"), ("ref1", "XYZ"), ("ref1", "TRI"), ("ref2", "XYZ"), ("ref2", "TRI"), ("ref3", "TRI")]
WRITER = {"ref1": "SU2", "ref2": "XYZ", "ref3": "TRI"}  # the core each .ref file was written by
TOLERANCE = 1e-5


def _fixture():
    from spinparser_b200.pfd import read_pfd
    return read_pfd(FIXTURE)


def _golden_datasets(fx, run):
    pre = f"{run}/ref/"
    return {k[len(pre):]: v for k, v in fx.items() if k.startswith(pre) and k.endswith("/data") and "/data/measurement_" in k}


needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "test", "scripted", "assets")), reason="the reference tree is not mounted here")


@needs_reference
def test_fixture_equals_the_reference_files():
    sys.path.insert(0, os.path.join(GOLDEN))
    from hdf5_v0 import read_hdf5
    import make_reference_goldens as gen
    fx = _fixture()
    total = 0
    for run, (ref_file, *_rest) in gen.TASKS.items():
        parsed = read_hdf5(os.path.join(REFERENCE, "test", "scripted", "assets", ref_file))
        for path, values in parsed.items():
            assert np.array_equal(fx[f"{run}/ref{path}"], values), path
        total += sum(1 for p in parsed if p.endswith("/data") and "/data/measurement_" in p)
    assert total == 544


@needs_reference
@pytest.mark.parametrize("run", ["ref1", "ref2", "ref3"])
def test_fp32_reference_build_reproduces_the_goldens_bit_for_bit(run):
    sys.path.insert(0, os.path.join(GOLDEN))
    import make_reference_goldens as gen
    ref_file, lattice, model, couplings, cores = gen.TASKS[run]
    # oracle32_pin: the FP32 sources as shipped, compiled for the baseline x86-64 target (no fused multiply-adds) -> bit for bit;
    # oracle32 (x86-64-v3, the binary the CPU baseline is timed with) contracts multiply-adds and differs in the last bits
    fx = _fixture()
    want = _golden_datasets(fx, run)
    assert len(want) in (68, 136, 340)
    for binary, exact in (("oracle32_pin", True), ("oracle32", False)):
        if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", binary)):
            pytest.skip(f"oracle/_ref/{binary} has not been built")
        got = gen.run_oracle(binary, gen.task_xml(lattice, model, couplings, cores[0]), ("--no-lattice",))
        for path, values in want.items():
            mine = np.asarray(got["h5/obs/" + path], dtype=np.float32)
            if exact:
                assert np.array_equal(mine, values), (binary, path)
            else:
                assert float(np.abs(mine - values).max()) < 1e-6, (binary, path)
            cutoff = path[:-len("/data")] + "@cutoff"
            assert np.array_equal(np.asarray(got["h5/obs/" + cutoff], dtype=np.float32).ravel(), fx[f"{run}/ref/{cutoff}"].ravel()), cutoff


def _gpu_flow(fx, run, core_name):
    """Full flow on the device from the fixture's tables; returns {dataset path: values} like the reference's .obs file."""
    from spinparser_b200 import FrgCoreFactory, ProblemTables
    from spinparser_b200.frgcore import correlation_datasets
    pre = f"{run}/{core_name}/"
    d = {k[len(pre):]: v for k, v in fx.items() if k.startswith(pre)}
    opts = {"spin": str(float(d["spinLength"]))} if core_name == "SU2" else {}
    core = FrgCoreFactory.newFrgCore(core_name, ProblemTables.from_pfd(d), opts)
    cutoffs = [float(x) for x in d["cutoff"]]
    n_basis = int(d["lattice/nBasis"])
    rid = [d[f"lattice/range{b}_fwd_rid"] for b in range(n_basis)]
    perm = [d[f"lattice/range{b}_fwd_perm"] for b in range(n_basis)]
    core.setInitialCondition(list(d["bare"]), cutoffs[0])
    out = {}
    for k, cutoff in enumerate(cutoffs):
        for name, values in correlation_datasets(core_name, core.measureCorrelation(), rid, perm).items():
            out[f"{name}/data/measurement_{k}/data"] = values
        if k + 1 < len(cutoffs):
            assert not core.computeStep()
            core.finalizeStep(cutoffs[k + 1])
    core.close()
    return out


@pytest.mark.gpu
def test_gpu_flow_reproduces_the_reference_goldens():
    fx = _fixture()
    flows = {(run, core): _gpu_flow(fx, run, core) for run, core in RUNS}
    worst = 0.0
    for run, writer in WRITER.items():
        want = _golden_datasets(fx, run)
        got = flows[(run, writer)]
        assert sorted(got) == sorted(want)
        for path, values in want.items():
            err = float(np.abs(got[path] - values).max())
            assert err < TOLERANCE, f"{run} {path}: deviation {err:.3e}"
            worst = max(worst, err)
    # cross-core identities (test/scripted/test_reference1.sh:70-78, test_reference2.sh): Heisenberg on SU2 == XYZ == TRI, Kitaev on XYZ == TRI
    def same(a, na, b, nb):
        keys = sorted(k for k in a if k.startswith(na + "/"))
        assert keys
        for k in keys:
            err = float(np.abs(a[k] - b[nb + k[len(na):]]).max())
            assert err < TOLERANCE, (k, err)
    su2, xyz, tri = flows[("ref1", "SU2")], flows[("ref1", "XYZ")], flows[("ref1", "TRI")]
    same(su2, "SU2CorDD", xyz, "XYZCorDD"); same(su2, "SU2CorDD", tri, "TRICorDD")
    for c in "XYZ":
        same(su2, "SU2CorZZ", xyz, f"XYZCor{c}{c}"); same(su2, "SU2CorZZ", tri, f"TRICor{c}{c}")
    xyz2, tri2 = flows[("ref2", "XYZ")], flows[("ref2", "TRI")]
    for c in ("XX", "YY", "ZZ", "DD"):
        same(xyz2, "XYZCor" + c, tri2, "TRICor" + c)
    print(f"\n[reference goldens] 544 datasets, worst absolute deviation of the GPU flow {worst:.2e} (reference tolerance {TOLERANCE:.0e})")
