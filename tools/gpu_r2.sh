#!/bin/bash
# Round-2 GPU sessions on one GPU. Usage (under gpurun): bash tools/gpu_r2.sh <tag> <stage> [...]
#   tests            whole GPU test suite (incl. the size-parity tests) + smoke
#   bench            default bench line (pyrochlore-r8, with parity + CPU baseline) and the reference arm
#   bench:<workload> full bench line of another workload
#   sweep:<workload> "<ENV=..>" ...   (until the next stage word) one bench line per configuration (synthetic state)
#   ncu:<workload>:<items> "<ENV=..>" launch list + one full capture of the flow kernel, summarised on the box
set -u
TAG=$1; shift
mkdir -p gpurun_out
export PFFRG_CACHE_DIR=$PWD/.jitcache
while [ $# -gt 0 ]; do
  stage=$1; shift
  case "$stage" in
    tests)
      python -m pytest tests -m gpu -q --timeout 900 --durations=15 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -n 25 gpurun_out/${TAG}_pytest_gpu.log
      grep -h "parity at size" gpurun_out/${TAG}_pytest_gpu.log
      python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -n 1 gpurun_out/${TAG}_smoke.log | cut -c1-200
      ;;
    sizetests)
      python -m pytest tests/test_parity_at_size.py -m gpu -q -s --timeout 900 > gpurun_out/${TAG}_pytest_size.log 2>&1; tail -n 12 gpurun_out/${TAG}_pytest_size.log
      ;;
    bench)
      timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; tail -c 600 gpurun_out/${TAG}_bench_default.err
      timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
      python tools/bench_digest.py gpurun_out/${TAG}_bench_default.json gpurun_out/${TAG}_bench_reference.json
      ;;
    bench:*)
      wl=${stage#bench:}
      timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_${wl}.json 2> gpurun_out/${TAG}_bench_${wl}.err; tail -c 400 gpurun_out/${TAG}_bench_${wl}.err
      python tools/bench_digest.py gpurun_out/${TAG}_bench_${wl}.json
      ;;
    sweep:*)
      wl=${stage#sweep:}
      cfgs=()
      while [ $# -gt 0 ] && [[ "$1" == *=* ]]; do cfgs+=("$1"); shift; done
      bash tools/gpu_sweep.sh $TAG $wl "${cfgs[@]}"
      ;;
    ncu:*)
      spec=${stage#ncu:}; wl=${spec%%:*}; items=${spec##*:}; cfg=${1:-X=1}; shift || true
      env $cfg timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_${wl}.csv \
        python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --synthetic-state > gpurun_out/${TAG}_ncu_launch_${wl}.log 2>&1
      python tools/launch_share.py gpurun_out/${TAG}_launches_${wl}.csv 2 | tee gpurun_out/${TAG}_launch_share_${wl}.txt
      rep=/tmp/${TAG}_prof_${wl}
      env $cfg timeout 900 ncu --set full --clock-control none --import-source on -k regex:v4 -s 3 -c 1 -o $rep -f \
        python bench.py --workload $wl --steps 1 --warmup 3 --synthetic-state --no-cpu-baseline --e2e-steps 0 --items $items > gpurun_out/${TAG}_ncu_full_${wl}.log 2>&1
      ls -la $rep.ncu-rep
      python tools/ncu_segments.py $rep.ncu-rep 0.3 > gpurun_out/${TAG}_${wl}_phases.txt 2>&1
      python tools/ncu_stalls.py $rep.ncu-rep 30 > gpurun_out/${TAG}_${wl}_stalls.txt 2>&1
      ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/${TAG}_${wl}_raw.csv 2>/dev/null
      head -n 14 gpurun_out/${TAG}_${wl}_phases.txt
      sz=$(stat -c %s $rep.ncu-rep); [ "$sz" -lt 30000000 ] && cp $rep.ncu-rep gpurun_out/${TAG}_prof_${wl}.ncu-rep
      ;;
    *) echo "unknown stage $stage";;
  esac
done
