#!/usr/bin/env python
"""Summarise `ncu --set full` reports of the flow kernel into profiles/ (run here, on the machine without a GPU).

    python tools/ncu_summary.py <workload> <report.ncu-rep> [<round tag> [<captured items>]]

Writes profiles/<tag>_<workload>_flow_kernel.csv (selected raw metrics, one line per captured launch) and updates
profiles/ncu_summary.json[workload] with the per-launch averages bench.py reports as `roofline.traffic`.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    # instruction delivery: per-SM instruction cache (ICC), GPC-level instruction cache (GCC) and its fill path from the crossbar
    "sm__icc_request_hit_rate.pct", "sm__icc_requests.sum.pct_of_peak_sustained_elapsed", "gcc__average_cache_request_hit_rate.pct",
    "gcc__cache_requests_type_instruction.sum", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
    "gcc__xbar2gcc_sectors.sum", "gcc__xbar2gcc_sectors.sum.pct_of_peak_sustained_elapsed",
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "second": 1.0}


def main(workload, report, tag="r1", items=0):
    raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(header)}
    name = col["Kernel Name"]
    keep = [m for m in METRICS if m in col]
    out = os.path.join(ROOT, "profiles", f"{tag}_{workload}_flow_kernel.csv")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{m} [{units[col[m]]}]" for m in keep])
        for r in data:
            w.writerow([r[name]] + [r[col[m]] for m in keep])

    def avg(metric):
        scale = UNIT.get(units[col[metric]], 1.0)
        vals = [float(r[col[metric]].replace(",", "")) * scale for r in data]
        return sum(vals) / len(vals)

    summary_path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    summary = json.load(open(summary_path)) if os.path.exists(summary_path) else {}
    summary[workload] = {
        "source": os.path.basename(out), "kernel": data[0][name], "launches": len(data),
        "duration_ms_under_ncu": avg("gpu__time_duration.sum") * 1e3,
        "dram_bytes_per_launch": avg("dram__bytes_read.sum") + avg("dram__bytes_write.sum"),
        "l2_bytes_per_launch": avg("lts__t_bytes.sum") if "lts__t_bytes.sum" in col else None,
        "l2_hit_pct": avg("lts__t_sector_hit_rate.pct"), "l1_hit_pct": avg("l1tex__t_sector_hit_rate.pct"),
        "fp64_pipe_pct": avg("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": avg("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": avg("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": avg("launch__registers_per_thread"),
        "l1_lsu_wavefronts_pct": avg("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") if "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed" in col else None,
        "gcc_instruction_requests_pct": avg("gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed") if "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed" in col else None,
        "gcc_fill_pct": avg("gcc__xbar2gcc_sectors.sum.pct_of_peak_sustained_elapsed") if "gcc__xbar2gcc_sectors.sum.pct_of_peak_sustained_elapsed" in col else None,
        "gcc_hit_pct": avg("gcc__average_cache_request_hit_rate.pct") if "gcc__average_cache_request_hit_rate.pct" in col else None,
        # 0 = the whole step; otherwise the capture was restricted to the first N work items (bench.py --items) and the
        # per-launch figures are NOT those of a full step
        "captured_items": int(items),
    }
    with open(summary_path, "w") as f:
        json.dump(summary, f, indent=1, sort_keys=True)
    print(json.dumps(summary[workload], indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:])
