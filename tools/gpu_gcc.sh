#!/bin/bash
# Instruction-delivery counters (GPC-level instruction cache "GCC", per-SM instruction cache "ICC") of the flow kernel.
# Usage (under gpurun): bash tools/gpu_gcc.sh <tag> <workload> "<ENV=..>" [...]
set -u
TAG=$1; WL=$2; shift 2
mkdir -p gpurun_out
export PFFRG_CACHE_DIR=$PWD/.jitcache
M=gpu__time_duration.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,gcc__xbar2gcc_sectors.sum,gcc__xbar2gcc_sectors.sum.pct_of_peak_sustained_elapsed,gcc__average_cache_request_hit_rate.pct,sm__icc_request_hit_rate.pct,sm__icc_requests.sum,sm__icc_requests.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active
i=0
for cfg in "$@"; do
  i=$((i+1))
  env PFFRG_AUTOTUNE=0 $cfg timeout 600 ncu --metrics $M --clock-control none -k regex:v4 -s 3 -c 1 --csv --log-file gpurun_out/${TAG}_gcc_${WL}_$i.csv \
    python bench.py --workload $WL --steps 1 --warmup 3 --synthetic-state --no-cpu-baseline --e2e-steps 0 > gpurun_out/${TAG}_gcc_${WL}_$i.log 2>&1
  echo "== $WL | $cfg"; python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_gcc_${WL}_$i.csv")) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r)); print("  %-75s %s %s" % (d["Metric Name"], d["Metric Value"], d["Metric Unit"]))
PY
done
