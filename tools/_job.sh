python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tri" > gpurun_out/r1y_pytest_gpu.log 2>&1; tail -3 gpurun_out/r1y_pytest_gpu.log
bash tools/gpu_sweep.sh r1y kagome_dm_r7_tri_nw64 "X=1"
