python -m pytest tests/test_gpu_parity.py -m gpu -q -k correlation > gpurun_out/r1s_pytest_gpu.log 2>&1; tail -30 gpurun_out/r1s_pytest_gpu.log | cut -c1-250
