python -m pytest tests -m gpu -q > gpurun_out/r1i_pytest_gpu.log 2>&1; tail -8 gpurun_out/r1i_pytest_gpu.log
bash tools/gpu_ncu.sh r1i kagome_dm_r7_tri_nw64:3000
