#!/bin/bash
# Round-end GPU session (second half of round 1): full GPU test suite, smoke, the default bench (with CPU baseline) and the
# reference arm, benches of the other workloads, launch list of the TIMED steps of the default bench, instruction-delivery
# counters and one full ncu capture of the flow kernel in the autotuned shape.
# Usage (under gpurun): bash tools/gpu_final.sh <tag>
set -u
TAG=${1:-r3f}
mkdir -p gpurun_out
export PFFRG_CACHE_DIR=$PWD/.jitcache
python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -n 1 gpurun_out/${TAG}_smoke.log | cut -c1-200
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_cubic_r7_su2_nw64.json 2> gpurun_out/${TAG}_bench_cubic.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
for wl in honeycomb_kitaev_r7_xyz_nw64 pyrochlore_r8_su2_nw64 kagome_dm_r7_tri_nw64 square_r4_su2_nw32; do
  timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --synthetic-state > gpurun_out/${TAG}_bench_${wl}.json 2> gpurun_out/${TAG}_bench_${wl}.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(f.split("_bench_")[1][:-5], "| steps/s %.3f" % d["value"], "| ms/step %.2f" % d["ms_per_step"], "| e2e %.3f" % (d["e2e"]["value"] or 0), "| bound %s frac %s" % (r.get("bound"), r.get("frac")), "| alg-hbm %s" % (r.get("alg_hbm") or {}).get("frac"), "| parity", d.get("parity"), "| cpu", (d.get("cpu_baseline") or {}).get("value"), "|", d.get("launch_shape"))
    except Exception as e:
        print(f, "FAILED", e)
PY
# launch list of the same command (synthetic state so that the list is short; autotuning launches come first, the timed steps last)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_cubic_r7_su2_nw64.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --synthetic-state > gpurun_out/${TAG}_ncu_launch.log 2>&1
python tools/launch_share.py gpurun_out/${TAG}_launches_cubic_r7_su2_nw64.csv 2 | tee gpurun_out/${TAG}_launch_share.txt
# full capture + instruction-delivery counters of the flow kernel in the shape the autotuner picks (read from the bench line)
SHAPE=$(python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_cubic_r7_su2_nw64.json").read().strip().splitlines()[-1])["launch_shape"]
s = d["sub_ctas"]
print(f"PFFRG_THREADS={d['threads'] // s} PFFRG_JIT_NBT={d['rpa_batch'] // s} PFFRG_JIT_NB={d['node_batch']} PFFRG_JIT_MINBLOCKS={d['min_blocks']}" + (f" PFFRG_SUBCTAS={s}" if s > 1 else ""))
PY
)
echo "autotuned shape: $SHAPE"
bash tools/gpu_gcc.sh ${TAG} cubic_r7_su2_nw64 "$SHAPE" | tee gpurun_out/${TAG}_gcc_cubic.txt
bash tools/gpu_session.sh ${TAG} --ncu cubic_r7_su2_nw64:0 "PFFRG_AUTOTUNE=0 $SHAPE"
