python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/r1k_pytest_multigpu.log 2>&1; tail -6 gpurun_out/r1k_pytest_multigpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r1k_bench_cubic_n2.json 2> gpurun_out/r1k_bench_cubic_n2.err; tail -c 1500 gpurun_out/r1k_bench_cubic_n2.json; tail -3 gpurun_out/r1k_bench_cubic_n2.err
python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r1k_bench_cubic_n1.json 2> gpurun_out/r1k_bench_cubic_n1.err; tail -c 800 gpurun_out/r1k_bench_cubic_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1k_bench_cubic_ref.json 2> gpurun_out/r1k_bench_cubic_ref.err; tail -c 800 gpurun_out/r1k_bench_cubic_ref.json
