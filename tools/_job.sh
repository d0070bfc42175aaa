N=$1
python -m pytest tests/test_multi_gpu.py -m gpu -q -k xyz > gpurun_out/r2b_pytest_multigpu_n$N.log 2>&1; tail -3 gpurun_out/r2b_pytest_multigpu_n$N.log
for wl in cubic_r7_su2_nw64 pyrochlore_r8_su2_nw64 kagome_dm_r7_tri_nw64; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --workload $wl --e2e-steps 1 --synthetic-state > gpurun_out/r2b_bench_${wl}_n$N.json 2> gpurun_out/r2b_bench_${wl}_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2b_bench_${wl}_n$N.json").read().strip().splitlines()[-1])
    print("$wl n=$N value", round(d["value"],3), "ms/step", round(d["ms_per_step"],2), d["breakdown_ms"])
except Exception as e:
    print("$wl FAILED", e, open("gpurun_out/r2b_bench_${wl}_n$N.err").read()[-1500:])
PY
done
