#!/bin/bash
# ncu full captures of the flow kernel. Usage (under gpurun): bash tools/gpu_ncu.sh <tag> <workload>[:items] [...]
# `:items` restricts the captured launch to the first N work items (big workloads replay ~45 passes).
set -u
TAG=${1:-r1}; shift || true
mkdir -p gpurun_out
for spec in "$@"; do
  wl=${spec%%:*}; items=0; [[ "$spec" == *:* ]] && items=${spec##*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:v4 -s 3 -c 1 -o gpurun_out/${TAG}_prof_${wl} -f \
    python bench.py --workload $wl --steps 1 --warmup 3 --synthetic-state --no-cpu-baseline --e2e-steps 0 --items $items > gpurun_out/${TAG}_ncu_full_${wl}.log 2>&1
  ls -la gpurun_out/${TAG}_prof_${wl}.ncu-rep
done
