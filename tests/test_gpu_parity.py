"""Parity of the CUDA path (through the C ABI) against dumps of the unmodified reference in FP64 (tests/golden/*.f64.pfd)."""
import numpy as np
import pytest

from conftest import assert_parity, dumped_steps, golden

pytestmark = pytest.mark.gpu

CASES = ["su2_square_r3_nw10", "su2_kagome_r4_nw8", "su2_kagome_r7_nw6", "xyz_honeycomb_kitaev_r3_nw10", "xyz_kagome_r4_nw8",
         "tri_honeycomb_kg_r3_nw8", "tri_kagome_dm_r3_nw6"]


def _core(d):
    from spinparser_b200 import FrgCoreFactory, ProblemTables
    name = bytes(d["core"]).decode()
    opts = {"spin": str(float(d["spinLength"]))} if name == "SU2" else {}
    return name, FrgCoreFactory.newFrgCore(name, ProblemTables.from_pfd(d), opts)


# kernel variants: the run-time compiled lattice-specialised kernel (default for SU2/XYZ), the precompiled generic kernels
# (PFFRG_JIT=0) and their smaller gather batches (PFFRG_NB; NB = 8 selects the TRI core's rpaTri8 phase, which large
# lattices use)
VARIANTS = [{}, {"PFFRG_JIT": "0"}, {"PFFRG_JIT": "0", "PFFRG_NB": "8"}, {"PFFRG_JIT_NBT": "32", "PFFRG_JIT_NB": "16"}, {"PFFRG_AUTOTUNE": "1"},
            {"PFFRG_THREADS": "128", "PFFRG_JIT_NBT": "32", "PFFRG_JIT_NB": "16", "PFFRG_JIT_MINBLOCKS": "4"},
            # several work items per CTA (sub-CTAs sharing one RPA phase): 2, 3 and 4 items, a partial last CTA (550 items), and with
            # 16 nodes per round several RPA rounds per item with sub-CTAs that run out of nodes at different times
            {"PFFRG_SUBCTAS": "2", "PFFRG_THREADS": "128", "PFFRG_JIT_NBT": "32", "PFFRG_JIT_NB": "16", "PFFRG_JIT_MINBLOCKS": "2"},
            {"PFFRG_SUBCTAS": "4", "PFFRG_THREADS": "128", "PFFRG_JIT_NBT": "32", "PFFRG_JIT_NB": "16"},
            {"PFFRG_SUBCTAS": "3", "PFFRG_THREADS": "64", "PFFRG_JIT_NBT": "16", "PFFRG_JIT_NB": "8"},
            # thread-block clusters whose CTAs rendezvous before every RPA phase (grid padded to whole clusters)
            {"PFFRG_CLUSTER": "2"}, {"PFFRG_CLUSTER": "4", "PFFRG_SUBCTAS": "2", "PFFRG_THREADS": "64", "PFFRG_JIT_NBT": "16", "PFFRG_JIT_NB": "8"},
            # t channel: buffers 2, 3 formed from the rows loaded for buffers 0, 1 (gatherTwo)
            {"PFFRG_MIRROR": "1"},
            # small CTAs on their own: fewer warps than the preferred number of node groups (the shape search must skip those)
            {"PFFRG_THREADS": "64"}, {"PFFRG_THREADS": "96"},
            # SU2: Gram form of the RPA phase (rpaGram) -- default shape, several RPA phases per item, other thread grids / block sizes
            {"PFFRG_RPA": "code"}, {"PFFRG_RPA": "code", "PFFRG_CLUSTER": "2"}, # (lattices above PFFRG_GRAM_MIN_TERMS run the Gram form by default)
            {"PFFRG_RPA": "gram"}, {"PFFRG_RPA": "gram", "PFFRG_JIT_NBT": "16", "PFFRG_JIT_NB": "8"}, {"PFFRG_RPA": "gram", "PFFRG_THREADS": "128"},
            {"PFFRG_RPA": "gram", "PFFRG_THREADS": "512", "PFFRG_GRAM_TM": "1"}, {"PFFRG_RPA": "gram", "PFFRG_THREADS": "96", "PFFRG_JIT_NBT": "8", "PFFRG_JIT_NB": "8"},
            {"PFFRG_RPA": "gram", "PFFRG_GRAM_TM": "1", "PFFRG_JIT_MINBLOCKS": "1"},
            # t-major CTA -> work item map (CTAs that run at the same time share the transfer frequency t)
            {"PFFRG_ORDER": "t", "PFFRG_CLUSTER": "1"}, {"PFFRG_ORDER": "t", "PFFRG_RPA": "gram"},
            # TRI: the table-driven RPA phase of the precompiled kernels instead of the Gram form; Gram form with 1 / 2 resident blocks
            {"PFFRG_FUSED_LOCALS": "1"}, {"PFFRG_FUSED_LOCALS": "1", "PFFRG_RPA": "gram"},
            # SU2 Gram kernel with a producer warp that builds the access buffers one batch ahead of the workers
            {"PFFRG_RPA": "gram", "PFFRG_PRODUCER": "1"}, {"PFFRG_RPA": "gram", "PFFRG_PRODUCER": "1", "PFFRG_JIT_NBT": "16", "PFFRG_JIT_NB": "8"},
            {"PFFRG_RPA": "gram", "PFFRG_PRODUCER": "1", "PFFRG_THREADS": "128", "PFFRG_JIT_MINBLOCKS": "2"}, {"PFFRG_RPA": "gram", "PFFRG_PRODUCER": "2"},
            # warp-specialised SU2 Gram kernel (gather / RPA / producer warp groups): default, several RPA rounds per item, one gather group + one producer warp
            {"PFFRG_RPA": "gram", "PFFRG_SPLIT": "1"}, {"PFFRG_RPA": "gram", "PFFRG_SPLIT": "1", "PFFRG_JIT_NBT": "8", "PFFRG_JIT_NB": "8"},
            {"PFFRG_RPA": "gram", "PFFRG_SPLIT": "1", "PFFRG_THREADS": "128", "PFFRG_PRODUCER": "1"}, {"PFFRG_RPA": "gram", "PFFRG_SPLIT": "1", "PFFRG_PERSISTENT": "0"},
            {"PFFRG_RPA": "table"}, {"PFFRG_RPA": "gram", "PFFRG_TRIGRAM_RESIDENT": "1"}, {"PFFRG_RPA": "gram", "PFFRG_TRIGRAM_RESIDENT": "2", "PFFRG_THREADS": "128"}]


# XYZ core: Gram form of the RPA phase (spin channels as virtual sites) in the warp-specialised kernel, PFFRG_RPA=gram
XYZ_GRAM_VARIANTS = [{"PFFRG_RPA": "gram"}, {"PFFRG_RPA": "gram", "PFFRG_JIT_NBT": "16", "PFFRG_JIT_NB": "8"}, {"PFFRG_RPA": "gram", "PFFRG_SPLIT": "1"},
                     {"PFFRG_RPA": "gram", "PFFRG_SPLIT": "1", "PFFRG_JIT_NBT": "8", "PFFRG_JIT_NB": "8"}, {"PFFRG_RPA": "gram", "PFFRG_SPLIT": "1", "PFFRG_PERSISTENT": "0"},
                     {"PFFRG_RPA": "gram", "PFFRG_SPLIT": "1", "PFFRG_THREADS": "128", "PFFRG_PRODUCER": "1"}]


@pytest.mark.parametrize("variant", VARIANTS, ids=lambda v: ",".join(f"{k}={x}" for k, x in v.items()) or "default")
@pytest.mark.parametrize("case", CASES)
def test_one_step_flow_matches_reference(case, variant, monkeypatch):
    if case.startswith("tri") and ("PFFRG_FUSED_LOCALS" in variant or "PFFRG_JIT_NBT" in variant or "PFFRG_AUTOTUNE" in variant or "PFFRG_SUBCTAS" in variant or "PFFRG_CLUSTER" in variant or "PFFRG_MIRROR" in variant):
        pytest.skip("the TRI core has no run-time compiled variant")
    if case.startswith("xyz") and "PFFRG_RPA" in variant and variant not in XYZ_GRAM_VARIANTS:
        pytest.skip("XYZ: the Gram form exists in the warp-specialised kernel only; SU2 shape knobs")
    if not case.startswith("su2") and ("PFFRG_PRODUCER" in variant or "PFFRG_SPLIT" in variant) and not (case.startswith("xyz") and variant in XYZ_GRAM_VARIANTS):
        pytest.skip("SU2 only")
    if case.startswith("tri") and variant.get("PFFRG_RPA") == "gram" and len(variant) > 1 and "PFFRG_TRIGRAM_RESIDENT" not in variant:
        pytest.skip("SU2 shape knobs")
    if not case.startswith("tri") and (variant.get("PFFRG_RPA") == "table" or "PFFRG_TRIGRAM_RESIDENT" in variant):
        pytest.skip("TRI only")
    for k, x in variant.items():
        monkeypatch.setenv(k, x)
    d = golden(case)
    name, core = _core(d)
    if variant.get("PFFRG_RPA") == "gram":
        assert core.stats()["jit_rpa"] == 1 and core.stats()["gram_rows"] > 0
    if variant.get("PFFRG_RPA") in ("table", "code") or variant.get("PFFRG_JIT") == "0":
        assert core.stats()["gram_rows"] == 0
    n = core.n_arrays
    cut = d["cutoff"]
    for step in dumped_steps(d):
        pre = f"step{step}/"
        state = [np.ascontiguousarray(d[pre + f"state/v4_{c}"]) for c in range(n)]
        core.setState(float(d[pre + "state/cutoff"]), np.ascontiguousarray(d[pre + "state/v2"]), state)
        diverged = core.computeStep()
        assert not diverged
        flow = core.flow()
        assert_parity(flow.v2, d[pre + "flow/v2"], f"{case} step {step} v2 flow")
        for c in range(n):
            assert_parity(flow.v4[c], d[pre + f"flow/v4_{c}"], f"{case} step {step} v4 flow channel {c}")
        # Euler update against the reference's own next state where it was dumped
        if step + 1 < len(cut):
            core.finalizeStep(float(cut[step + 1]))
            new = core.flowingFunctional()
            assert new.cutoff == float(cut[step + 1])
            want = [state[c] + (float(cut[step + 1]) - float(cut[step])) * d[pre + f"flow/v4_{c}"] for c in range(n)]
            for c in range(n):
                assert_parity(new.v4[c], want[c], f"{case} step {step} Euler channel {c}")
    core.close()


@pytest.mark.parametrize("variant", [{}, {"PFFRG_CLUSTER": "4"}, {"PFFRG_SUBCTAS": "3", "PFFRG_THREADS": "64", "PFFRG_JIT_NBT": "16", "PFFRG_JIT_NB": "8", "PFFRG_CLUSTER": "2"},
                                     {"PFFRG_ORDER": "t", "PFFRG_CLUSTER": "1"}, {"PFFRG_ORDER": "t", "PFFRG_RPA": "gram"},
                                     {"PFFRG_RPA": "gram", "PFFRG_SPLIT": "1"}, {"PFFRG_RPA": "gram", "PFFRG_SPLIT": "1", "PFFRG_PERSISTENT": "0"}],
                         ids=lambda v: ",".join(f"{k}={x}" for k, x in v.items()) or "default")
@pytest.mark.parametrize("case", ["su2_kagome_r4_nw8", "xyz_kagome_r4_nw8"])
def test_item_range_with_padded_grid(case, variant, monkeypatch):
    """A slice of the work items that is not a multiple of the cluster / sub-CTA size (what a rank of a sharded run computes): the
    grid is padded to whole clusters, padding CTAs only take part in the barriers, and the slice equals the reference's."""
    for k, x in variant.items():
        monkeypatch.setenv(k, x)
    d = golden(case)
    name, core = _core(d)
    L, n = core.tables.n_sites, core.n_arrays
    step = dumped_steps(d)[-1]
    pre = f"step{step}/"
    core.setState(float(d[pre + "state/cutoff"]), np.ascontiguousarray(d[pre + "state/v2"]), [np.ascontiguousarray(d[pre + f"state/v4_{c}"]) for c in range(n)])
    begin, end = 7, 7 + 101
    core.setItemRange(begin, end)
    assert not core.computeStep()
    flow = core.flow()
    for c in range(n):
        assert_parity(flow.v4[c][begin * L:end * L], d[pre + f"flow/v4_{c}"][begin * L:end * L], f"{case} items [{begin}, {end}) channel {c}")
    core.close()


@pytest.mark.parametrize("case", ["su2_square_r3_nw10", "xyz_honeycomb_kitaev_r3_nw10", "tri_kagome_dm_r3_nw6"])
def test_float32_host_arrays_round_trip(case):
    d = golden(case)
    name, core = _core(d)
    n = core.n_arrays
    rng = np.random.default_rng(7)
    v2 = rng.uniform(0, 0.5, core.tables.n_frequencies).astype(np.float32)
    v4 = [rng.uniform(-1, 1, core.array_length).astype(np.float32) for _ in range(n)]
    core.setState(1.25, v2, v4)
    back = core.flowingFunctional(np.float32)
    assert back.cutoff == 1.25
    assert np.array_equal(back.v2, v2)
    for c in range(n):
        assert np.array_equal(back.v4[c], v4[c])
    core.close()


@pytest.mark.parametrize("case", CASES)
def test_device_correlation_measurement_matches_reference(case):
    """K5: chi(Lambda) computed on the device against the datasets the unmodified reference wrote for the same state
    (tests/golden/*.f64.pfd, `h5/obs/<Core>Cor<mu nu>/data/measurement_<step>/data`). North-star tolerance 1e-8 relative with an
    absolute floor of 1e-10 of the largest correlation (off-site values of symmetry-forbidden components are round-off)."""
    from spinparser_b200.frgcore import correlation_datasets
    d = golden(case)
    name, core = _core(d)
    n = core.n_arrays
    n_basis = int(d["lattice/nBasis"])
    rid = [d[f"lattice/range{b}_fwd_rid"] for b in range(n_basis)]
    perm = [d[f"lattice/range{b}_fwd_perm"] for b in range(n_basis)]
    checked = 0
    for step in dumped_steps(d):
        pre = f"step{step}/"
        state = [np.ascontiguousarray(d[pre + f"state/v4_{c}"]) for c in range(n)]
        core.setState(float(d[pre + "state/cutoff"]), np.ascontiguousarray(d[pre + "state/v2"]), state)
        got = correlation_datasets(name, core.measureCorrelation(), rid, perm)
        scale = max(np.abs(d[f"h5/obs/{key}/data/measurement_{step}/data"]).max() for key in got)
        for key, values in got.items():
            want = d[f"h5/obs/{key}/data/measurement_{step}/data"]
            assert values.shape == want.shape, key
            assert float(d[f"h5/obs/{key}/data/measurement_{step}@cutoff"][0]) == float(d[pre + "state/cutoff"])
            err = np.abs(values - want)
            assert (err <= 1e-8 * np.abs(want) + 1e-10 * scale).all(), f"{case} step {step} {key}: max deviation {err.max():.3e} (scale {scale:.3e})"
            checked += 1
    assert checked >= 4
    core.close()


@pytest.mark.parametrize("case", ["su2_kagome_r4_nw8", "xyz_honeycomb_kitaev_r3_nw10", "tri_kagome_dm_r3_nw6"])
def test_initial_condition_built_on_the_device(case):
    """pffrg_set_initial_condition against the initial state the reference constructs (SU2EffectiveAction.hpp:38-60 etc.)."""
    d = golden(case)
    name, core = _core(d)
    L = core.tables.n_sites
    if name == "TRI":
        bare = list(np.asarray(d["initial/v4_0"][: 16 * L]).reshape(16, L))
    else:
        bare = [np.asarray(d[f"initial/v4_{c}"][:L]) for c in range(core.n_arrays)]
    core.setInitialCondition(bare, float(d["initial/cutoff"]))
    got = core.flowingFunctional()
    assert got.cutoff == float(d["initial/cutoff"])
    assert np.array_equal(got.v2, d["initial/v2"])
    for c in range(core.n_arrays):
        assert np.array_equal(got.v4[c], d[f"initial/v4_{c}"]), c
    core.close()


def test_call_sequence_violations_are_reported():
    """PFFRG_ERR_STATE (-4) instead of undefined behaviour: compute / measure before a state exists, finalize before compute."""
    from spinparser_b200 import PffrgError
    d = golden("su2_square_r3_nw10")
    name, core = _core(d)
    for call in (core.computeStep, core.measureCorrelation, lambda: core.finalizeStep(1.0), core.flow, core.flowingFunctional):
        with pytest.raises(PffrgError) as err:
            call()
        assert err.value.code == -4, call
    core.setState(float(d["step0/state/cutoff"]), np.ascontiguousarray(d["step0/state/v2"]), [np.ascontiguousarray(d[f"step0/state/v4_{c}"]) for c in range(2)])
    with pytest.raises(PffrgError) as err:
        core.finalizeStep(1.0)
    assert err.value.code == -4
    with pytest.raises(PffrgError):
        core.setState(1.0, np.zeros(3), [np.zeros(5), np.zeros(5)])  # wrong sizes are rejected on the host side
    core.close()


@pytest.mark.parametrize("case", ["su2_square_r3_nw10", "xyz_honeycomb_kitaev_r3_nw10", "tri_kagome_dm_r3_nw6"])
def test_divergence_is_reported(case):
    """A vertex that overflows in the flow equations: compute_step reports `diverged` (the reference finds NaN in `_flow`,
    SU2EffectiveAction.hpp:212-230, and stops the loop, SpinParser.cpp:151-155); a finite state does not."""
    d = golden(case)
    name, core = _core(d)
    n = core.n_arrays
    step = dumped_steps(d)[-1]
    pre = f"step{step}/"
    v2 = np.ascontiguousarray(d[pre + "state/v2"])
    v4 = [np.ascontiguousarray(d[pre + f"state/v4_{c}"]) for c in range(n)]
    core.setState(float(d[pre + "state/cutoff"]), v2, v4)
    assert core.computeStep() is False
    core.setState(float(d[pre + "state/cutoff"]), v2, [a * 1e160 for a in v4])
    assert core.computeStep() is True
    flow = core.flow()
    assert any(np.isnan(a).any() for a in flow.v4)
    # and the core recovers with a finite state
    core.setState(float(d[pre + "state/cutoff"]), v2, v4)
    assert core.computeStep() is False
    core.close()


def _one_step_matches(core, d, case):
    n = core.n_arrays
    step = dumped_steps(d)[-1]
    pre = f"step{step}/"
    core.setState(float(d[pre + "state/cutoff"]), np.ascontiguousarray(d[pre + "state/v2"]), [np.ascontiguousarray(d[pre + f"state/v4_{c}"]) for c in range(n)])
    assert not core.computeStep()
    flow = core.flow()
    for c in range(n):
        assert_parity(flow.v4[c], d[pre + f"flow/v4_{c}"], f"{case} step {step} channel {c}")


@pytest.mark.parametrize("case", ["su2_kagome_r7_nw6", "xyz_kagome_r4_nw8"])
def test_failed_runtime_compilation_falls_back_to_the_precompiled_kernels(case, monkeypatch, capfd):
    """A failure of the run-time compilation (no NVRTC, a compiler error, ...; injected here) is not a reason to refuse a lattice the
    precompiled kernels can run: pffrg_create reports it on stderr and goes on without the specialised kernel. PFFRG_JIT_STRICT=1, or an
    explicit request for a kernel form, keeps it an error."""
    from spinparser_b200 import PffrgError
    monkeypatch.setenv("PFFRG_JIT_INJECT_FAILURE", "1")
    d = golden(case)
    name, core = _core(d)
    assert "continuing with the precompiled kernels" in capfd.readouterr().err
    assert core.stats()["jit_rpa"] == 0 and core.stats()["gather_threads"] == 0
    _one_step_matches(core, d, case)
    core.close()
    monkeypatch.setenv("PFFRG_JIT_STRICT", "1")
    with pytest.raises(PffrgError, match="injected failure"):
        _core(d)
    monkeypatch.delenv("PFFRG_JIT_STRICT")
    if case.startswith("su2"):
        monkeypatch.setenv("PFFRG_RPA", "gram")
        with pytest.raises(PffrgError, match="injected failure"):
            _core(d)


def test_rejected_cache_entry_is_deleted_and_compiled_afresh(monkeypatch, tmp_path, capfd):
    """The on-disk kernel cache: an entry the driver rejects (truncated here) is deleted and the kernel compiled again."""
    monkeypatch.setenv("PFFRG_CACHE_DIR", str(tmp_path))
    case = "su2_kagome_r7_nw6"
    d = golden(case)
    name, core = _core(d)
    assert core.stats()["jit_rpa"] == 1
    core.close()
    entries = list(tmp_path.glob("*.cubin"))
    assert entries
    sizes = {e: e.stat().st_size for e in entries}
    for e in entries:
        e.write_bytes(e.read_bytes()[:4096])
    capfd.readouterr()
    name, core = _core(d)
    assert "compiling afresh" in capfd.readouterr().err
    assert core.stats()["jit_rpa"] == 1
    _one_step_matches(core, d, case)
    core.close()
    assert all(e.stat().st_size == sizes[e] for e in entries if e.exists()) and any(e.exists() for e in entries)
