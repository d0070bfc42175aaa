python -m pytest tests -m gpu -q > gpurun_out/r1r_pytest_gpu.log 2>&1; tail -4 gpurun_out/r1r_pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1r_bench_cubic.json 2> gpurun_out/r1r_bench_cubic.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r1r_bench_cubic.json").read().strip().splitlines()[-1])
print("cubic value", round(d["value"],3), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), d["breakdown_ms"])
PY
