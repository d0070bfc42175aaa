#!/bin/bash
# ncu captures of the flow kernel for the given workloads: launch list + one full capture each.
# Usage (under gpurun): bash tools/gpu_ncu.sh <tag> <workload> [...]
set -u
TAG=${1:-r1}; shift || true
mkdir -p gpurun_out
for wl in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:v4 -s 3 -c 1 -o gpurun_out/${TAG}_prof_${wl} -f \
    python bench.py --workload $wl --steps 1 --warmup 3 --synthetic-state --no-cpu-baseline --e2e-steps 0 > gpurun_out/${TAG}_ncu_full_${wl}.log 2>&1
  ls -la gpurun_out/${TAG}_prof_${wl}.ncu-rep
done
