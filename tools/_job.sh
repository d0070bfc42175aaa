python -m pytest tests -m gpu -x -q > gpurun_out/r1c_pytest_gpu.log 2>&1; tail -5 gpurun_out/r1c_pytest_gpu.log
bash tools/gpu_sweep.sh r1c cubic_r7_su2_nw64 "X=1" "PFFRG_JIT_TILES=2" "PFFRG_JIT_NBT=32" "PFFRG_JIT_NBT=32 PFFRG_JIT_TILES=4" "PFFRG_JIT_NBT=128" "PFFRG_JIT_NBT=128 PFFRG_THREADS=512" "PFFRG_JIT_NBT=128 PFFRG_JIT_TILES=2 PFFRG_THREADS=512" "PFFRG_JIT_CHUNK=16 PFFRG_JIT_ACC=8"
bash tools/gpu_sweep.sh r1c honeycomb_kitaev_r7_xyz_nw64 "X=1" "PFFRG_JIT_NBT=64" "PFFRG_JIT_NBT=64 PFFRG_JIT_TILES=2"
bash tools/gpu_sweep.sh r1c pyrochlore_r8_su2_nw64 "X=1"
