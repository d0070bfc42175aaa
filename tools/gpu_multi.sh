#!/bin/bash
# Multi-GPU session: sharded-vs-single-GPU bit-exactness tests (peer-memory and NCCL exchange, sharded upload, shape broadcast), the
# two-rank run through the C++ adapter, and bench lines at N GPUs.
# Usage (under gpurun --gpus N): bash tools/gpu_multi.sh <tag> <N> [--notests] [workload[:ENV=..,ENV=..] ...]
set -u
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
export PFFRG_CACHE_DIR=$PWD/.jitcache
if [ "${1:-}" == "--notests" ]; then shift; else
  timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_host_adapter.py::test_two_ranks_through_the_adapter -m gpu -q --timeout 600 > gpurun_out/${TAG}_pytest_multigpu_n$N.log 2>&1; tail -n 8 gpurun_out/${TAG}_pytest_multigpu_n$N.log
fi
for spec in "$@"; do
  wl=${spec%%:*}; cfg="X=1"; tag=""; [[ "$spec" == *:* ]] && { cfg=${spec#*:}; cfg=${cfg//,/ }; tag="_$(echo $cfg | tr -c 'A-Za-z0-9=' '_')"; }
  env $cfg timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --workload $wl --e2e-steps 2 > gpurun_out/${TAG}_bench_${wl}_n$N$tag.json 2> gpurun_out/${TAG}_bench_${wl}_n$N$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${wl}_n$N$tag.json").read().strip().splitlines()[-1])
    print("$wl n=$N $cfg | value", round(d["value"],3), "| ms/step", round(d["ms_per_step"],3), "| e2e", round(d["e2e"]["value"],3), "|", {k: round(v,3) for k,v in d["breakdown_ms"].items()}, d["launch_shape"]["threads"], d["launch_shape"]["gram_rows"])
except Exception as e:
    print("$wl FAILED", e, open("gpurun_out/${TAG}_bench_${wl}_n$N$tag.err").read()[-1500:])
PY
done
