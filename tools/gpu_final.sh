#!/bin/bash
# Round-end GPU session: full GPU test suite, smoke, the default bench (with CPU baseline) and the reference arm, benches of
# the other workloads, the ncu launch list of the default bench and full captures of the flow kernel.
# Usage (under gpurun): bash tools/gpu_final.sh <tag> [workload[:items] ...]
set -u
TAG=${1:-r1}; shift || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log | cut -c1-200
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_cubic_r7_su2_nw64.json 2> gpurun_out/${TAG}_bench_cubic.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
for wl in honeycomb_kitaev_r7_xyz_nw64 kagome_dm_r7_tri_nw64 pyrochlore_r8_su2_nw64 square_r4_su2_nw32; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --synthetic-state > gpurun_out/${TAG}_bench_${wl}.json 2> gpurun_out/${TAG}_bench_${wl}.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(f.split("_bench_")[1][:-5], "| steps/s %.3f" % d["value"], "| ms/step %.2f" % d["ms_per_step"], "| e2e %.3f" % (d["e2e"]["value"] or 0), "| hbm frac %s" % r.get("frac"), "| fp64 %s / %s" % (r.get("fp64_tflops_achieved"), r.get("fp64_tflops_peak_measured")), "| cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches_cubic_r7_su2_nw64.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/${TAG}_ncu_launch.log 2>&1
bash tools/gpu_ncu.sh ${TAG} "$@"
