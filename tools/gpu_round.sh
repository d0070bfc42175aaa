#!/bin/bash
# One GPU session: parity tests, smoke, benches of every workload, launch list and a full ncu capture of the flow kernel.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [ncu workload ...]
set -u
TAG=${1:-r1}; shift || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_cubic.json 2> gpurun_out/${TAG}_bench_cubic.err; tail -c 600 gpurun_out/${TAG}_bench_cubic.json
for wl in honeycomb_kitaev_r7_xyz_nw64 kagome_dm_r7_tri_nw64 pyrochlore_r8_su2_nw64 square_r4_su2_nw32; do
  timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --synthetic-state --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_bench_${wl}.json 2> gpurun_out/${TAG}_bench_${wl}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${wl}.json").read().strip().splitlines()[-1])
    print("${wl}", "ms/step", round(d["ms_per_step"],2), "frac", round(d["roofline"]["frac"],3), "TF64", round(d["roofline"]["fp64_tflops_achieved"],2))
except Exception as e:
    print("${wl} FAILED", e, open("gpurun_out/${TAG}_bench_${wl}.err").read()[-800:])
PY
done
for wl in "$@"; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches_${wl}.csv \
    python bench.py --workload $wl --steps 2 --warmup 3 --synthetic-state --no-cpu-baseline --e2e-steps 0 > gpurun_out/${TAG}_ncu_launch_${wl}.log 2>&1
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:v4 -s 3 -c 1 -o gpurun_out/${TAG}_prof_${wl} -f \
    python bench.py --workload $wl --steps 1 --warmup 3 --synthetic-state --no-cpu-baseline --e2e-steps 0 > gpurun_out/${TAG}_ncu_full_${wl}.log 2>&1
  ls -la gpurun_out/${TAG}_prof_${wl}.ncu-rep
done
