#!/bin/bash
# Tuning sweep of the run-time compiled flow kernel: one bench line per configuration into gpurun_out/<tag>_sweep.txt
# Usage (under gpurun): bash tools/gpu_sweep.sh <tag> <workload> "<ENV=.. ENV=..>" ["<...>" ...]
set -u
TAG=$1; WL=$2; shift 2
mkdir -p gpurun_out
OUT=gpurun_out/${TAG}_sweep_${WL}.txt
for cfg in "$@"; do
  env $cfg timeout ${SWEEP_TIMEOUT:-900} python bench.py --workload $WL --steps 3 --warmup 3 --synthetic-state --no-cpu-baseline --e2e-steps 0 2> gpurun_out/${TAG}_sweep.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$WL | $cfg |', 'v4 ms %.2f' % d['breakdown_ms']['ms_v4_flow'], '| step ms %.2f' % d['ms_per_step'], '| TF64 exec %.2f' % d['roofline']['candidates']['fp64']['achieved'], '|', d['launch_shape'])
except Exception as e:
    print('$WL | $cfg | FAILED', e, open('gpurun_out/${TAG}_sweep.err').read()[-300:])
" >> $OUT
  tail -1 $OUT
done
