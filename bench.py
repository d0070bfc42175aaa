#!/usr/bin/env python
"""bench.py -- pf-FRG cutoff steps/s of the B200 flow core (and of the reference CPU core with --impl reference).

A "step" is one cutoff step of the flow equations: computeStep (self-energy flow, quadrature node table, vertex flow of
all Nw^2(Nw+1)/2 frequency triples x L sites x C channels) + finalizeStep (Euler update, multi-GPU exchange).

Workload (BASELINE.json configs[1]): examples/cubic-J1J2.xml geometry -- SU2 core, cubic lattice range 7 (L = 31
representatives, 575 sites in range, 8311 overlap terms), 64 positive frequencies (133 120 work items, 8.25 M vertex
entries), cutoff grid 50 * 0.98^k. The timed steps start at k = 211 (cutoff 0.704, ~62 quadrature nodes per item and
channel) from the PHYSICAL state reached by running the flow from the bare couplings on the GPU (untimed setup).

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
}
N_CH = {"SU2": 2, "XYZ": 4, "TRI": 16}


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


def profiled_traffic(workload):
    """DRAM bytes per launch of the flow kernel from the committed `ncu --set full` capture of this workload
    (profiles/ncu_summary.json, written by tools/ncu_summary.py), or None when there is none."""
    path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        rec = json.load(f).get(workload)
    return rec.get("dram_bytes_per_launch") if rec and not rec.get("captured_items") else None


def profiled_limits(workload):
    """What the committed full capture says binds the flow kernel (percent of the unit's peak, from profiles/ncu_summary.json):
    context for `roofline`, not a live measurement. Empty when there is no capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            rec = json.load(f).get(workload) or {}
        keys = {"l1_lsu_wavefronts_pct": "l1_data_pipe_wavefronts", "fp64_pipe_pct": "fp64_pipe", "issue_active_pct": "issue_slots",
                "gcc_instruction_requests_pct": "gpc_instruction_cache_requests", "gcc_fill_pct": "gpc_instruction_cache_fill"}
        return {name: round(float(rec[k]), 1) for k, name in keys.items() if rec.get(k) is not None}
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


def time_reference_cpu(workload, d, start_step, v2, v4, stride, repeat, warmup):
    """Time the reference's own CPU core (oracle/_ref/oracle32 = unmodified reference sources, FP32 as shipped, OpenMP
    `parallel for schedule(guided)` over work items as in src/lib/LoadManager.hpp:551-557) on every `stride`-th work item of
    one step. Falls back to the plain-C port (FP64) when the reference binary was not built. Returns (seconds per full step, info)."""
    nw, L = len(d["frequency"]), int(d["lattice/size"])
    nf = nw * nw * (nw + 1) // 2
    cores = os.cpu_count() or 1
    binary = os.path.join(ROOT, "oracle", "_ref", "oracle32")
    if os.path.exists(binary):
        with tempfile.TemporaryDirectory() as tmp:
            state = os.path.join(tmp, "state.pfd")
            write_pfd(state, {"v2": np.asarray(v2, dtype=np.float64), **{f"v4_{c}": np.asarray(a, dtype=np.float64) for c, a in enumerate(v4)}})
            cmd = [binary, "-r", os.path.join(ROOT, "oracle", "res"), os.path.join(ROOT, "bench_data", "tasks", workload + ".xml"),
                   "--mode", "time", "--load-state", state, "--start-step", str(start_step), "--time-stride", str(stride),
                   "--time-repeat", str(repeat), "--time-warmup", str(warmup), "--no-lattice", "--threads", str(cores)]
            out = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout
        rec = json.loads([ln for ln in out.splitlines() if ln.startswith("{")][-1])
        per_pass = rec["seconds"]
        scale = rec["items_total"] / rec["items"]
        return [s * scale for s in per_pass], {"kind": "reference", "cores": rec["threads"], "dtype": "f32",
                                                "sample": f"every {stride}th work item ({rec['items']} of {rec['items_total']}), {warmup} warm-up + {repeat} timed passes of one cutoff step, scaled by {scale:.1f}"}
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_port import OraclePort
    port = OraclePort(d)
    cutoff = float(d["cutoff"][start_step])
    items = np.arange(0, nf, stride, dtype=np.int32)
    times = []
    for rep in range(warmup + repeat):
        t0 = time.perf_counter()
        f2 = port.v2_flow(cutoff, v2, v4)
        port.v4_flow(cutoff, v2, f2, v4, items)
        if rep >= warmup:
            times.append((time.perf_counter() - t0) * nf / len(items))
    return times, {"kind": "port", "cores": cores, "dtype": "f64",
                   "sample": f"every {stride}th work item ({len(items)} of {nf}), {warmup} warm-up + {repeat} timed passes, scaled"}


def run_reference(args, d):
    """--impl reference: the reference CPU core on this box's host cores, same workload / metric / unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v2, v4 = synthetic_state(d)
    times, info = time_reference_cpu(args.workload, d, args.start_step, v2, v4, args.cpu_stride, args.steps, args.warmup)
    sec = sum(times) / len(times)
    core = bytes(d["core"]).decode()
    nw, L = len(d["frequency"]), int(d["lattice/size"])
    nf = nw * nw * (nw + 1) // 2
    value = 1.0 / sec
    line = {
        "impl": "reference", "metric": "pf-FRG cutoff steps/s", "value": value, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": info["dtype"], "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "cutoff_step": args.start_step, "cutoff": float(d["cutoff"][args.start_step]),
                   "state": "seeded synthetic vertex (step cost is state independent)", "device": "host CPU"},
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
    ap.add_argument("--workload", default="cubic_r7_su2_nw64", choices=sorted(WORKLOADS))
    ap.add_argument("--start-step", type=int, default=211)
    ap.add_argument("--cpu-stride", type=int, default=16, help="CPU baseline: time every n-th work item")
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
    core = FrgCoreFactory.newFrgCore(core_name, ProblemTables.from_pfd(d), opts, device=local)
    if world > 1:
        ids = [core.uniqueId() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        core.initCommunicator(ids[0], rank, world)

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
        core.setState(cutoffs[args.start_step], sv2, sv4)
        step = args.start_step
    else:
        core.setInitialCondition(list(d["bare"]), cutoffs[0])
    t_setup = time.perf_counter()
    while step < args.start_step:
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
        st["ms_finalize"], st["ms_exchange"] = st2["ms_finalize"], st2["ms_exchange"]
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
    sums = torch.tensor([sum(st[k] for _, _, st in records) for k in ("kernel_evals", "alg_bytes", "alg_flops")], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    total_ms = float(step_ms.sum())
    ms_per_step = total_ms / args.steps
    value = 1e3 / ms_per_step
    kernel_ms_avg = float(kern_ms.mean())
    evals, alg_bytes, alg_flops = (float(x) / args.steps for x in sums)

    # ---- end to end through the public API with HOST buffers: upload state, step, download state, every step
    host = core.pinnedEffectiveAction()
    core.flowingFunctional(into=host)
    e2e_times = []
    for _ in range(args.e2e_steps):
        barrier()
        t0 = time.perf_counter()
        core.setState(host.cutoff, host.v2, host.v4)
        if core.computeStep():
            raise SystemExit("flow diverged in the end-to-end leg")
        step += 1
        core.finalizeStep(cutoffs[step])
        core.flowingFunctional(into=host)
        barrier()
        e2e_times.append(time.perf_counter() - t0)
    e2e_t = torch.tensor(e2e_times, dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = 1.0 / float(e2e_t.mean()) if args.e2e_steps else None
    state_bytes = 8 * (nw + C * L * nf)
    state_dev_mb = 8e-6 * C * ((L + 3) // 4 * 4) * nf

    if rank == 0:
        peak, peak_src = measured_peaks()
        from spinparser_b200._capi import lib as _lib
        fp64_peak = float(_lib.pffrg_fp64_peak(local))  # measured here: 16-chain DFMA loop on every SM
        achieved = alg_bytes / world / (kernel_ms_avg * 1e-3) / 1e9  # per GPU: this rank's share over its kernel time
        line = {
            "metric": "pf-FRG cutoff steps/s", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload], "cutoff_steps": [first_timed, first_timed + args.steps - 1],
                       "cutoff": [cutoffs[first_timed], cutoffs[first_timed + args.steps - 1]],
                       **({"PROFILING_ONLY_item_range": [0, args.items]} if args.items > 0 else {}),
                       "state": "seeded synthetic vertex" if args.synthetic_state else f"physical: flow run on the GPU from the bare couplings for {first_timed} steps ({t_setup:.1f} s for the first {args.start_step}, untimed)",
                       "parallelism": f"work items sharded over {world} GPU(s), vertex replicated, ncclBroadcast exchange of the updated slices",
                       "l2": f"flushed between timed iterations (256 MiB write; device vertex {state_dev_mb:.0f} MB vs 126 MB L2)", "timing": "CUDA events on the library stream per step, max over ranks"},
            "vertex_entries_per_s": C * L * nf * value,
            "kernel_evals_per_step": evals,
            "alg_gb_per_step": alg_bytes / 1e9, "alg_gflop_per_step": alg_flops / 1e9,
            "wall_ms_per_step_incl_flush": wall * 1e3 / args.steps,
            "breakdown_ms": {k: statistics.mean(st[k] for _, _, st in records) for k in ("ms_v2_flow", "ms_node_table", "ms_v4_flow", "ms_finalize", "ms_exchange")},
            "launch_shape": {k: records[0][2][k] for k in ("jit_rpa", "threads", "smem_bytes", "node_batch", "rpa_batch", "rpa_warps", "min_blocks", "sub_ctas", "autotuned_shapes", "jit_compile_ms")},
            "roofline": {"bound": "hbm", "kernel": "pffrg::v4FlowKernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": profiled_traffic(args.workload), "profiled_pct_of_peak": profiled_limits(args.workload),
                         "peak_source": peak_src, "note": "achieved = algorithmic gather+output bytes (SURVEY 8d: 128*L*C per kernel evaluation + 24*L*C per item) / kernel time, per GPU; gathers that hit in L2 do not reach DRAM, so `traffic` (ncu dram bytes per launch) is far below the algorithmic bytes and frac can exceed 1",
                         "fp64_tflops_achieved": alg_flops / world / (kernel_ms_avg * 1e-3) / 1e12,
                         "fp64_tflops_peak_measured": fp64_peak, "fp64_frac": alg_flops / world / (kernel_ms_avg * 1e-3) / 1e12 / fp64_peak if fp64_peak > 0 else None},
            "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": state_bytes,
                    "what": "setState(pinned host arrays) + computeStep + finalizeStep + flowingFunctional(download) per step, wall clock, max over ranks"},
            "gpu_launches": 6 * args.steps,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            host_state = core.flowingFunctional()
            times, info = time_reference_cpu(args.workload, d, step, host_state.v2, host_state.v4, args.cpu_stride, 1, 1)
            line["cpu_baseline"] = {"value": 1.0 / (sum(times) / len(times)), "unit": "steps/s", "cores": info["cores"], "kind": info["kind"], "sample": info["sample"],
                                    "cutoff": cutoffs[step]}
        print(json.dumps(line), flush=True)
    core.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
