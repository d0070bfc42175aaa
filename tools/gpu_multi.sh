#!/bin/bash
# Multi-GPU session: sharded-vs-single-GPU bit-exactness tests and bench lines at N GPUs.
# Usage (under gpurun --gpus N): bash tools/gpu_multi.sh <tag> <N> [workload ...]
set -u
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
export PFFRG_CACHE_DIR=$PWD/.jitcache
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 300 > gpurun_out/${TAG}_pytest_multigpu_n$N.log 2>&1; tail -n 3 gpurun_out/${TAG}_pytest_multigpu_n$N.log
for wl in "$@"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --workload $wl --e2e-steps 1 --synthetic-state > gpurun_out/${TAG}_bench_${wl}_n$N.json 2> gpurun_out/${TAG}_bench_${wl}_n$N.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${wl}_n$N.json").read().strip().splitlines()[-1])
    print("$wl n=$N value", round(d["value"],3), "ms/step", round(d["ms_per_step"],2), d["breakdown_ms"], d["launch_shape"])
except Exception as e:
    print("$wl FAILED", e, open("gpurun_out/${TAG}_bench_${wl}_n$N.err").read()[-1500:])
PY
done
