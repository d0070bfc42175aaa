#!/bin/bash
# One GPU session of the second half of round 1: parity tests, A/B sweeps of kernel options (one bench line per
# configuration), optional full ncu captures summarised ON THE BOX (phases, stalls, raw metrics as text; the .ncu-rep of a
# kernel with ~1 MB of code is too large to bring back).
# Usage (under gpurun): bash tools/gpu_session.sh <tag> [--tests] [--sweep <workload> "<ENV=..>" ...]... [--ncu <workload>:<items> "<ENV=..>"]...
set -u
TAG=$1; shift
mkdir -p gpurun_out
export PFFRG_CACHE_DIR=$PWD/.jitcache   # cubins precompiled in the build container travel with the snapshot
while [ $# -gt 0 ]; do
  case "$1" in
    --tests)
      shift
      python -m pytest tests -m gpu -x -q --timeout 180 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/${TAG}_pytest_gpu.log
      ;;
    --sweep)
      WL=$2; shift 2
      cfgs=()
      while [ $# -gt 0 ] && [[ "$1" != --* ]]; do cfgs+=("$1"); shift; done
      bash tools/gpu_sweep.sh $TAG $WL "${cfgs[@]}"
      ;;
    --ncu)
      spec=$2; cfg=${3:-X=1}; shift 3 || shift $#
      wl=${spec%%:*}; items=0; [[ "$spec" == *:* ]] && items=${spec##*:}
      rep=/tmp/${TAG}_prof_${wl}
      env $cfg timeout 900 ncu --set full --clock-control none --import-source on -k regex:v4 -s 3 -c 1 -o $rep -f \
        python bench.py --workload $wl --steps 1 --warmup 3 --synthetic-state --no-cpu-baseline --e2e-steps 0 --items $items > gpurun_out/${TAG}_ncu_full_${wl}.log 2>&1
      ls -la $rep.ncu-rep
      python tools/ncu_segments.py $rep.ncu-rep 0.3 > gpurun_out/${TAG}_${wl}_phases.txt 2>&1
      python tools/ncu_stalls.py $rep.ncu-rep 30 > gpurun_out/${TAG}_${wl}_stalls.txt 2>&1
      ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/${TAG}_${wl}_raw.csv 2>/dev/null
      head -n 12 gpurun_out/${TAG}_${wl}_phases.txt
      sz=$(stat -c %s $rep.ncu-rep); [ "$sz" -lt 20000000 ] && cp $rep.ncu-rep gpurun_out/
      ;;
    *) echo "unknown argument $1"; shift;;
  esac
done
