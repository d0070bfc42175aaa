#!/usr/bin/env python
"""bench.py -- pf-FRG cutoff steps/s of the B200 flow core (and of the reference CPU core with --impl reference).

A "step" is one cutoff step of the flow equations: computeStep (self-energy flow, quadrature node table, vertex flow of
all Nw^2(Nw+1)/2 frequency triples x L sites x C channels) + finalizeStep (Euler update, multi-GPU exchange).

Default workload (BASELINE.json config 5, the largest single-GPU configuration and the only one whose vertex exceeds the L2):
pyrochlore Heisenberg -- SU2 core, pyrochlore lattice range 8 (L = 103 representatives, 49 334 overlap terms), 64 positive
frequencies (133 120 work items, 27.4 M vertex entries, 222 MB FP64 vertex), cutoff grid 50 * 0.98^k. `--workload` selects the other
configurations (cubic-J1J2 = configs[1], square-Heisenberg = configs[0], kagome-DM on the TRI core, Kitaev honeycomb on the XYZ core).
The timed steps start at k = 211 (cutoff 0.704, ~62 quadrature nodes per item and channel) from the PHYSICAL state reached by
running the flow from the bare couplings on the GPU (untimed setup). The reference arm (--impl reference) times the reference's own
CPU core on the same cutoff steps.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--start-step S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Prints ONE JSON line on rank 0. Multi-GPU is strong scaling: the work items of every step are sharded over the ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from spinparser_b200.pfd import read_pfd, write_pfd  # noqa: E402

WORKLOADS = {
    "cubic_r7_su2_nw64": "cubic-J1J2 (SU2, cubic r=7, L=31, Nw=64, 133120 items, 8.25M vertex entries)",
    "square_r4_su2_nw32": "square-Heisenberg (SU2, square r=4, L=9, Nw=32, 16896 items)",
    "pyrochlore_r8_su2_nw64": "pyrochlore-Heisenberg (SU2, pyrochlore r=8, L=103, Nw=64, 133120 items, 27.4M vertex entries)",
    "honeycomb_kitaev_r7_xyz_nw64": "honeycomb-Kitaev (XYZ, honeycomb r=7, L=18, Nw=64, 133120 items, 9.6M vertex entries)",
    "kagome_dm_r7_tri_nw64": "kagome-DM (TRI, kagome r=7, L=34, Nw=64, 133120 items, 72.4M vertex entries)",
    "pyrochlore_r10_su2_nw64": "pyrochlore-Heisenberg (SU2, pyrochlore r=10, L=185, Nw=64, 133120 items, 49.3M vertex entries)",
    "honeycomb_kitaev_r10_xyz_nw64": "honeycomb-Kitaev (XYZ, honeycomb r=10, L=33, Nw=64, 133120 items, 17.6M vertex entries)",
}
N_CH = {"SU2": 2, "XYZ": 4, "TRI": 16}
DEFAULT_WORKLOAD = "pyrochlore_r8_su2_nw64"
# CPU legs: every n-th work item, n coprime to Nw (items are su-major / t-minor: a stride sharing a factor with Nw would only ever
# visit a few t values), sized for 10-30 s of CPU work per pass on a 16-core host
CPU_STRIDE = {"cubic_r7_su2_nw64": 17, "square_r4_su2_nw32": 3, "pyrochlore_r8_su2_nw64": 67, "honeycomb_kitaev_r7_xyz_nw64": 17,
              "kagome_dm_r7_tri_nw64": 131, "pyrochlore_r10_su2_nw64": 131, "honeycomb_kitaev_r10_xyz_nw64": 33}


def node_counts(d, cutoff):
    """Quadrature nodes (kernel evaluations) per channel for every transfer frequency of the mesh at this cutoff: the conventional
    single-scale terms plus the nodes of the three Katanin segments (src/SU2/SU2FrgCore.cpp:351-392 x src/lib/Integrator.hpp:138-287,
    node counts as in SURVEY.md 8a). Pure host arithmetic on the mesh (no device, no oracle): used to scale a strided CPU sample to the
    whole step by the exact ratio of kernel evaluations."""
    mesh = [float(x) for x in d["frequency"]]
    nw = len(mesh)

    def first_greater(w):
        lo, hi = 1, nw
        while lo < hi:
            mid = (lo + hi) // 2
            if mesh[mid] > w:
                hi = mid
            else:
                lo = mid + 1
        return lo

    def greater_pos(w):
        if w <= mesh[0]:
            return 0
        i = first_greater(w)
        return i if i < nw else nw - 1

    def lesser_pos(w):
        if w <= mesh[0]:
            return 0
        i = first_greater(w)
        return i - 1 if i < nw else nw - 1

    lesser = lambda w: -(greater_pos(-w) + 1) if w < 0 else lesser_pos(w)
    greater = lambda w: -(lesser_pos(-w) + 1) if w < 0 else greater_pos(w)
    out = []
    for x in mesh:
        n = 1 + (1 if x > 2.0 * cutoff else 0)
        if -(x + cutoff) > -mesh[-1]:
            umax = lesser(-(x + cutoff))
            n += (umax + nw) + 2 if umax != -nw else 2
        if x - cutoff > cutoff:
            umin, umax = greater(cutoff - x), lesser(-cutoff)
            n += (umax - umin) + 3 if umax >= umin else 2
        if cutoff < mesh[-1]:
            umin = greater(cutoff)
            n += (nw - 1 - umin) + 2 if umin != nw - 1 else 2
        out.append(n)
    return out


def evaluations_of_items(counts, items):
    """Kernel evaluations (s + t + u channel) of the listed work items; item = su * Nw + t, su = so (so + 1) / 2 + uo."""
    nw = len(counts)
    c = np.asarray(counts, dtype=np.int64)
    items = np.asarray(items, dtype=np.int64)
    su, t = items // nw, items % nw
    so = ((np.sqrt(8.0 * su + 1.0) - 1.0) * 0.5).astype(np.int64)
    so = np.where((so + 1) * (so + 2) // 2 <= su, so + 1, so)
    so = np.where(so * (so + 1) // 2 > su, so - 1, so)
    uo = su - so * (so + 1) // 2
    return int((c[so] + c[uo] + c[t]).sum())


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region (B200_PROFILING.md recipe)."""

    QUERY = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device: int):
        self.device, self.proc, self.lines, self.first, self.last = device, None, [], 0, None

    def mark_begin(self):
        """Start of the timed region: samples taken before it (warm-up) are dropped."""
        self.first = len(self.lines)

    def mark_end(self):
        """End of the timed region (the sample in flight still belongs to it)."""
        self.last = len(self.lines) + 1

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
            time.sleep(0.5)  # nvidia-smi takes a few hundred ms to deliver its first sample
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        window = self.lines[self.first:self.last] or self.lines[-1:]
        for line in window:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons), "samples": len(sm)}


def profiled_record(workload):
    """Summary of the committed `ncu --set full` capture of this workload's flow kernel (profiles/ncu_summary.json, written by
    tools/ncu_summary.py): context for `roofline` (DRAM bytes per launch, L1 data-pipe wavefronts, ...), not a live measurement.
    Empty when there is no capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            return json.load(f).get(workload) or {}
    except Exception:  # never let a reporting extra break the bench line
        return {}


def load_tables(workload):
    return read_pfd(os.path.join(ROOT, "bench_data", workload + ".tables.pfd"))


def synthetic_state(d, seed=20261017):
    """Seeded synthetic state: v4 ~ U(-0.1, 0.1), v2 ~ U(0, 0.5). The cost of a step does not depend on the values."""
    core = bytes(d["core"]).decode()
    nw, L = len(d["frequency"]), int(d["lattice/size"])
    nf = nw * nw * (nw + 1) // 2
    rng = np.random.default_rng(seed)
    n_arrays = {"SU2": 2, "XYZ": 4, "TRI": 1}[core]
    length = nf * L * (16 if core == "TRI" else 1)
    return rng.uniform(0.0, 0.5, nw), [rng.uniform(-0.1, 0.1, length) for _ in range(n_arrays)]


def time_reference_cpu(workload, d, step, v2, v4, stride, repeat, warmup, offset=0, want_rows=False, binary_name="oracle32"):
    """Time the reference's own CPU core (oracle/_ref/oracle32 = unmodified reference sources, FP32 as shipped, OpenMP
    `parallel for schedule(guided)` over work items as in src/lib/LoadManager.hpp:551-557) on every `stride`-th work item of
    cutoff step `step`, starting from the given state. The sample is scaled to the whole step by the exact ratio of kernel
    evaluations (quadrature nodes of the s, t and u channel of every item, `node_counts`). Falls back to the plain-C port (FP64)
    when the reference binary was not built. Returns (seconds per full step for every timed pass, info, sampled flow rows or None)."""
    nw, L = len(d["frequency"]), int(d["lattice/size"])
    nf = nw * nw * (nw + 1) // 2
    core = bytes(d["core"]).decode()
    per = L * (16 if core == "TRI" else 1)
    cores = os.cpu_count() or 1
    items = np.arange(offset, nf, stride, dtype=np.int32)
    counts = node_counts(d, float(d["cutoff"][step]))
    scale = evaluations_of_items(counts, np.arange(nf)) / evaluations_of_items(counts, items)
    binary = os.path.join(ROOT, "oracle", "_ref", binary_name)
    if os.path.exists(binary):
        with tempfile.TemporaryDirectory() as tmp:
            state, out = os.path.join(tmp, "state.pfd"), os.path.join(tmp, "out.pfd")
            write_pfd(state, {"v2": np.asarray(v2, dtype=np.float64), **{f"v4_{c}": np.asarray(a, dtype=np.float64) for c, a in enumerate(v4)}})
            cmd = [binary, "-r", os.path.join(ROOT, "oracle", "res"), os.path.join(ROOT, "bench_data", "tasks", workload + ".xml"),
                   "--mode", "time", "--time-compact", "--load-state", state, "--start-step", str(step), "--time-stride", str(stride), "--time-offset", str(offset),
                   "--time-repeat", str(repeat), "--time-warmup", str(warmup), "--no-lattice", "--threads", str(cores)] + (["--out", out] if want_rows else [])
            stdout = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout
            rows = None
            if want_rows:
                r = read_pfd(out)
                assert np.array_equal(r["time/itemIds"], items)
                rows = (items, np.asarray(r["time/flow/v2"], dtype=np.float64), [np.asarray(r[f"time/flowItems/v4_{c}"], dtype=np.float64).reshape(len(items), per) for c in range(len(v4))])
        rec = json.loads([ln for ln in stdout.splitlines() if ln.startswith("{")][-1])
        assert rec["items"] == len(items)
        return [sec * scale for sec in rec["seconds"]], {"kind": "reference", "cores": rec["threads"], "dtype": "f32" if binary_name == "oracle32" else "f64", "items": len(items), "scale": scale,
                                                          "sample": f"every {stride}th work item of cutoff step {step} ({len(items)} of {nf}; stride coprime to Nw), {warmup} warm-up + {repeat} timed passes, "
                                                                    f"scaled by the exact ratio of kernel evaluations ({scale:.2f})"}, rows
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_port import OraclePort
    port = OraclePort(d)
    cutoff = float(d["cutoff"][step])
    times, rows = [], None
    v2 = np.ascontiguousarray(v2, dtype=np.float64)
    v4 = [np.ascontiguousarray(a, dtype=np.float64) for a in v4]
    for rep in range(warmup + repeat):
        t0 = time.perf_counter()
        f2 = port.v2_flow(cutoff, v2, v4)
        full = port.v4_flow(cutoff, v2, f2, v4, items)
        if rep >= warmup:
            times.append((time.perf_counter() - t0) * scale)
        if want_rows:
            rows = (items, f2, [a.reshape(-1, per)[items] for a in full])
    return times, {"kind": "port", "cores": cores, "dtype": "f64", "items": len(items), "scale": scale,
                   "sample": f"every {stride}th work item of cutoff step {step} ({len(items)} of {nf}), {warmup} warm-up + {repeat} timed passes, scaled by the exact ratio of kernel evaluations ({scale:.2f})"}, rows


def workload_config(args, d, first_step, steps):
    """`config` of the JSON line: identical for both arms (what is computed, not how)."""
    cutoffs = [float(x) for x in d["cutoff"]]
    return {"workload": WORKLOADS[args.workload], "core": bytes(d["core"]).decode(), "n_frequencies": len(d["frequency"]), "n_sites": int(d["lattice/size"]),
            "cutoff_steps": [first_step, first_step + steps - 1], "cutoff": [cutoffs[first_step], cutoffs[first_step + steps - 1]]}


def run_reference(args, d):
    """--impl reference: the reference CPU core on this box's host cores, same workload, metric, unit and cutoff steps as our arm.
    Step i of the K timed steps is a bounded sample of cutoff step start + i (a different sample offset for every step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v2, v4 = synthetic_state(d)
    stride = args.cpu_stride or CPU_STRIDE[args.workload]
    per_step, info = [], None
    for w in range(args.warmup):
        time_reference_cpu(args.workload, d, args.start_step, v2, v4, stride, 1, 0, offset=w % stride)
    for i in range(args.steps):
        times, info, _ = time_reference_cpu(args.workload, d, args.start_step + i, v2, v4, stride, 1, 0, offset=i % stride)
        per_step.append(times[0])
    sec = sum(per_step) / len(per_step)
    core = bytes(d["core"]).decode()
    nw, L = len(d["frequency"]), int(d["lattice/size"])
    nf = nw * nw * (nw + 1) // 2
    value = 1.0 / sec
    line = {
        "impl": "reference", "metric": "pf-FRG cutoff steps/s", "value": value, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": info["dtype"], "data": "synthetic",
        "config": workload_config(args, d, args.start_step, args.steps),
        "notes": {"state": "seeded synthetic vertex (the cost of a step depends on the mesh and the cutoff only, not on the values)", "device": "host CPU",
                  "extrapolated": "ms_per_step is the sampled time scaled to the full step; the run itself is bounded"},
        "vertex_entries_per_s": N_CH[core] * L * nf * value,
        "cpu_baseline": {"value": value, "unit": "steps/s", **{k: info[k] for k in ("cores", "kind", "sample")}},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--start-step", type=int, default=211)
    ap.add_argument("--cpu-stride", type=int, default=0, help="CPU legs: time every n-th work item (default: per workload, coprime to Nw)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--items", type=int, default=0, help="profiling runs: restrict every step to the first N work items (not a benchmark)")
    ap.add_argument("--synthetic-state", action="store_true", help="start from a seeded synthetic vertex at --start-step instead of running the flow there (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    # launch-shape autotuning of the run-time compiled kernel (a documented library option, see pffrg.cu setupJit); every shape is
    # covered by the parity tests
    os.environ.setdefault("PFFRG_AUTOTUNE", "1")
    d = load_tables(args.workload)
    if args.impl == "reference":
        run_reference(args, d)
        return

    import torch
    import torch.distributed as dist
    from spinparser_b200 import FrgCoreFactory, ProblemTables

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU core")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    core_name = bytes(d["core"]).decode()
    nw, L = len(d["frequency"]), int(d["lattice/size"])
    nf = nw * nw * (nw + 1) // 2
    C = N_CH[core_name]
    cutoffs = [float(x) for x in d["cutoff"]]
    opts = {"spin": str(float(d["spinLength"]))} if core_name == "SU2" else {}
    if world > 1:
        # rank 0 tunes the launch shape, the other ranks adopt its choice: one shape on all ranks (bit-identical sharding)
        core = FrgCoreFactory.newFrgCore(core_name, ProblemTables.from_pfd(d), opts, device=local) if rank == 0 else None
        shared = [(core.uniqueId(), core.shapeEnvironment()) if rank == 0 else None]
        dist.broadcast_object_list(shared, src=0)
        if rank != 0:
            os.environ.update(shared[0][1])
            core = FrgCoreFactory.newFrgCore(core_name, ProblemTables.from_pfd(d), opts, device=local)
        core.initCommunicator(shared[0][0], rank, world)
    else:
        core = FrgCoreFactory.newFrgCore(core_name, ProblemTables.from_pfd(d), opts, device=local)

    if args.items > 0:
        core.setItemRange(0, min(args.items, nf))
    stream = torch.cuda.ExternalStream(core.stream, device=torch.device("cuda", local))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=f"cuda:{local}")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- untimed setup: the real flow from the bare couplings down to the first timed cutoff
    step = 0
    if args.synthetic_state:
        sv2, sv4 = synthetic_state(d)
        step = max(0, args.start_step - (args.warmup + 2))
        core.setState(cutoffs[step], sv2, sv4)
    else:
        core.setInitialCondition(list(d["bare"]), cutoffs[0])
    t_setup = time.perf_counter()
    untimed = args.warmup + 2  # warm-up steps + the two steps the clock sampler needs to come up
    setup_target = step if args.synthetic_state else max(0, args.start_step - untimed)
    while step < setup_target:
        if core.computeStep():
            raise SystemExit(f"flow diverged during setup at step {step}")
        step += 1
        core.finalizeStep(cutoffs[step])
    barrier()
    t_setup = time.perf_counter() - t_setup

    def one_step(timed):
        nonlocal step
        with torch.cuda.stream(stream):
            flush.zero_()  # L2 flush between iterations (outside the event bracket)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        diverged = core.computeStep()
        st = core.stats()
        step += 1
        core.finalizeStep(cutoffs[step])
        e1.record(stream)
        st2 = core.stats()
        st["ms_finalize"], st["ms_exchange"], st["launches"] = st2["ms_finalize"], st2["ms_exchange"], st2["launches"]
        if diverged:
            raise SystemExit(f"flow diverged at step {step}")
        return e0, e1, st

    for _ in range(args.warmup):
        one_step(False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # nvidia-smi needs a moment to come up: started ahead of two extra untimed steps, the window is marked below
    for _ in range(2):
        one_step(False)
    barrier()
    sampler.mark_begin()
    first_timed = step
    wall0 = time.perf_counter()
    records = [one_step(True) for _ in range(args.steps)]
    barrier()
    wall = time.perf_counter() - wall0
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    step_ms = torch.tensor([e0.elapsed_time(e1) for e0, e1, _ in records], dtype=torch.float64, device=f"cuda:{local}")
    kern_ms = torch.tensor([st["ms_v4_flow"] for _, _, st in records], dtype=torch.float64, device=f"cuda:{local}")
    sums = torch.tensor([sum(st[k] for _, _, st in records) for k in ("kernel_evals", "alg_bytes", "alg_flops", "exec_flops")], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    total_ms = float(step_ms.sum())
    ms_per_step = total_ms / args.steps
    value = 1e3 / ms_per_step
    kernel_ms_avg = float(kern_ms.mean())
    evals, alg_bytes, alg_flops, exec_flops = (float(x) / args.steps for x in sums)
    launches = sum(st["launches"] for _, _, st in records)  # this rank's kernels (library counter): v2 flow, node table, vertex flow, Euler x2, cutoff

    # ---- end to end through the public API with HOST buffers: upload state, step, download state, every step
    host = core.pinnedEffectiveAction()
    core.flowingFunctional(into=host)
    e2e_times = []
    for _ in range(args.e2e_steps):
        barrier()
        t0 = time.perf_counter()
        core.setState(host.cutoff, host.v2, host.v4, sharded=world > 1)
        if core.computeStep():
            raise SystemExit("flow diverged in the end-to-end leg")
        step += 1
        core.finalizeStep(cutoffs[step])
        if world > 1:
            core.flowingFunctional(into=host, items=core.uploadSlice())  # every rank reads back its share of the rows
        else:
            core.flowingFunctional(into=host)
        barrier()
        e2e_times.append(time.perf_counter() - t0)
    e2e_t = torch.tensor(e2e_times, dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = 1.0 / float(e2e_t.mean()) if args.e2e_steps else None
    if world > 1 and args.e2e_steps:
        # the host copies are only partially refreshed by the slice downloads: bring them back in line for what follows
        core.flowingFunctional(into=host)
    state_bytes = 8 * (nw + C * L * nf)
    state_dev_mb = 8e-6 * C * ((L + 3) // 4 * 4) * nf

    # ---- parity inside the bench job + CPU baseline: the reference's own CPU core (FP32 as shipped) evaluates a strided sample of
    # the NEXT cutoff step from the GPU's current state; the GPU evaluates the same step; the sampled rows must agree to FP32 accuracy
    parity, cpu_baseline = None, None
    if world == 1 and not args.no_cpu_baseline and args.items == 0:
        host_state = core.flowingFunctional()
        stride = args.cpu_stride or CPU_STRIDE[args.workload]
        # The reference computes in FP32: entries of the FP64 state below 1e-15 of the largest one (round-off in symmetry-forbidden components)
        # become DENORMAL floats there, or their products do -- values a pure-FP32 run never holds and which slow x86 arithmetic 3-4 x
        # (measured). They are set to zero for the CPU leg (eight orders below FP32 resolution and the 1e-3 comparison;
        # the FP64 parity leg below gets the state as it is)
        def fp32_safe(a):
            a = np.asarray(a)
            return np.where(np.abs(a) < 1e-15 * max(float(np.abs(a).max()), 1e-300), 0.0, a)
        times, info, rows = time_reference_cpu(args.workload, d, step, fp32_safe(host_state.v2), [fp32_safe(a) for a in host_state.v4], stride, 1, 1, want_rows=True)
        cpu_baseline = {"value": 1.0 / (sum(times) / len(times)), "unit": "steps/s", "cores": info["cores"], "kind": info["kind"],
                        "sample": info["sample"] + "; state entries below 1e-15 of the largest one (denormal range of the FP32 core) set to zero", "cutoff_step": step, "cutoff": cutoffs[step]}
        if core.computeStep():
            raise SystemExit("flow diverged in the parity leg")
        gpu_flow = core.flow()
        per = L * (16 if core_name == "TRI" else 1)

        def deviation(rows, rel, floor):
            """(max norm-wise deviation, all entries within rel * |x| + floor * max|x| per array)"""
            items, ref_v2, ref_rows = rows
            worst, ok = 0.0, True
            for got, want in [(gpu_flow.v2, ref_v2)] + [(gpu_flow.v4[c].reshape(-1, per)[items], want) for c, want in enumerate(ref_rows)]:
                scale = float(np.abs(want).max())
                err = np.abs(got - want)
                worst = max(worst, float(err.max() / max(scale, 1e-300)))
                ok = ok and bool((err <= rel * np.abs(want) + floor * scale).all())
            return worst, ok

        # (1) against the FP64 build of the unmodified reference on a coarser sample: the north-star criterion, 1e-10 relative per entry
        #     with a floor of 1e-12 of the largest entry; (2) against the FP32 reference core that was just timed: the
        #     reference's own FP32 and FP64 builds differ by 1e-5 .. 3e-4 norm-wise in one step, so 1e-3 only catches gross errors
        _, info64, rows64 = time_reference_cpu(args.workload, d, step, host_state.v2, host_state.v4, stride * 8 + 1, 1, 0, offset=3, want_rows=True, binary_name="oracle64")
        worst64, ok64 = deviation(rows64, 1e-10, 1e-12) if info64["dtype"] == "f64" else (None, True)
        worst32, ok32 = deviation(rows, 1e-3, 1e-3)
        parity = {"items": int(info64["items"]), "max_normwise": worst64, "tolerance": "|d| <= 1e-10 |x| + 1e-12 max|x| per array", "against": f"{info64['kind']} CPU core (f64 build of the unmodified reference) from the GPU's own state at cutoff step {step}",
                  "f32_reference": {"items": int(len(rows[0])), "max_normwise": worst32, "tolerance": 1e-3}, "ok": bool(ok64 and ok32)}

    if rank == 0:
        peak, peak_src = measured_peaks()
        from spinparser_b200._capi import lib as _lib
        fp64_peak = float(_lib.pffrg_fp64_peak(local))  # measured here: 16-chain DFMA loop on every SM
        t_kernel = kernel_ms_avg * 1e-3
        alg_gbs = alg_bytes / world / t_kernel / 1e9  # per GPU: this rank's share over its kernel time
        exec_tflops = exec_flops / world / t_kernel / 1e12
        prof = profiled_record(args.workload) if world == 1 else {}
        # what can bind the flow kernel: DRAM traffic (ncu capture), the FP64 pipe (live: flops as executed / measured DFMA peak), the L1
        # data pipe (ncu capture: LSU wavefronts, 128 B per clock and SM). `bound` is the largest fraction; the contractual figure from
        # ALGORITHMIC bytes is kept as `alg_hbm` (gathers that hit in L1 / L2 never reach DRAM, so it can exceed 1 and is no HBM fraction).
        candidates = {"fp64": {"achieved": exec_tflops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": exec_tflops / fp64_peak if fp64_peak > 0 else None, "source": "live: executed flops / kernel time; peak = DFMA loop measured in this run"}}
        traffic = None
        if prof.get("dram_bytes_per_launch") and prof.get("duration_ms_under_ncu"):
            traffic = prof["dram_bytes_per_launch"]
            dram_gbs = traffic / (prof["duration_ms_under_ncu"] * 1e-3) / 1e9
            candidates["hbm"] = {"achieved": dram_gbs, "peak": peak, "unit": "GB/s", "frac": dram_gbs / peak, "source": f"profiles/{prof.get('source')}: dram__bytes_read.sum + dram__bytes_write.sum per launch / launch duration under ncu"}
        if prof.get("l1_lsu_wavefronts_pct") is not None:
            l1_peak = 148 * 128 * (clocks["sm_max_mhz"] or 1965.0) * 1e6 / 1e9 if clocks else 148 * 128 * 1.965e3
            candidates["l1"] = {"achieved": prof["l1_lsu_wavefronts_pct"] / 100 * l1_peak, "peak": l1_peak, "unit": "GB/s", "frac": prof["l1_lsu_wavefronts_pct"] / 100,
                                "source": f"profiles/{prof.get('source')}: l1tex__data_pipe_lsu_wavefronts, % of peak (one 128-byte wavefront per clock and SM)"}
        bound = max((k for k in candidates if candidates[k]["frac"] is not None), key=lambda k: candidates[k]["frac"])
        line = {
            "metric": "pf-FRG cutoff steps/s", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, d, first_timed, args.steps),
            "notes": {**({"PROFILING_ONLY_item_range": [0, args.items]} if args.items > 0 else {}),
                      "state": "seeded synthetic vertex" if args.synthetic_state else f"physical: flow run on the GPU from the bare couplings for {first_timed} steps ({t_setup:.1f} s of setup, untimed)",
                      "parallelism": f"work items sharded over {world} GPU(s), vertex replicated, exchange of the updated slices over NVLink",
                      "l2": f"flushed between timed iterations (256 MiB write; device vertex {state_dev_mb:.0f} MB vs 126 MB L2)", "timing": "CUDA events on the library stream per step, max over ranks"},
            "vertex_entries_per_s": C * L * nf * value,
            "kernel_evals_per_step": evals,
            "alg_gb_per_step": alg_bytes / 1e9, "alg_gflop_per_step": alg_flops / 1e9, "exec_gflop_per_step": exec_flops / 1e9,
            "wall_ms_per_step_incl_flush": wall * 1e3 / args.steps,
            "breakdown_ms": {k: statistics.mean(st[k] for _, _, st in records) for k in ("ms_v2_flow", "ms_node_table", "ms_v4_flow", "ms_finalize", "ms_exchange")},
            "launch_shape": {k: records[0][2][k] for k in ("jit_rpa", "gram_rows", "rpa_terms_merged", "threads", "smem_bytes", "node_batch", "rpa_batch", "rpa_warps", "min_blocks", "sub_ctas", "autotuned_shapes", "jit_compile_ms", "gather_threads", "producer_warps")},
            "roofline": {"bound": bound, "kernel": "pffrg_v4flow_jit" if records[0][2]["jit_rpa"] else "pffrg::v4FlowKernel", **{k: candidates[bound][k] for k in ("achieved", "peak", "unit", "frac")},
                         "traffic": traffic, "candidates": candidates,
                         "alg_hbm": {"achieved": alg_gbs, "peak": peak, "unit": "GB/s", "frac": alg_gbs / peak, "peak_source": peak_src,
                                     "what": "ALGORITHMIC gather + output bytes (SURVEY 8d: 128*L*C per kernel evaluation + 24*L*C per item) / kernel time, per GPU"},
                         "alg_fp64": {"achieved": alg_flops / world / t_kernel / 1e12, "peak": fp64_peak, "unit": "TFLOP/s", "what": "flops of the reference formulation (every overlap term per node)"},
                         "profile_capture": {k: prof.get(k) for k in ("source", "duration_ms_under_ncu", "captured_items", "l1_hit_pct", "l2_hit_pct", "fp64_pipe_pct", "issue_active_pct", "warps_active_pct", "registers_per_thread")} if prof else None},
            "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": state_bytes,
                    "what": "setState(pinned host arrays) + computeStep + finalizeStep + flowingFunctional(download) per step, wall clock, max over ranks"
                            + ("; every rank holds the full host state, uploads 1/N of the rows (distributed over NVLink) and reads back its 1/N of the result: bytes are the totals over all ranks" if world > 1 else "")},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        if parity:
            line["parity"] = parity
        print(json.dumps(line), flush=True)
        if parity and not parity["ok"]:
            raise SystemExit(f"PARITY FAILURE inside the bench run: {parity}")
    core.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
