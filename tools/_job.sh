python -m pytest tests/test_host_adapter.py -m gpu -q > gpurun_out/r1t_pytest_gpu.log 2>&1; tail -30 gpurun_out/r1t_pytest_gpu.log | cut -c1-250
