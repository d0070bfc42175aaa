python -m pytest tests -m gpu -x -q > gpurun_out/r1e_pytest_gpu.log 2>&1; tail -3 gpurun_out/r1e_pytest_gpu.log
bash tools/gpu_sweep.sh r1e cubic_r7_su2_nw64 "X=1" "PFFRG_JIT_TILES=2" "PFFRG_JIT_NBT=32 PFFRG_JIT_MINBLOCKS=3 PFFRG_JIT_CHUNK=16" "PFFRG_JIT_NBT=32 PFFRG_JIT_TILES=2 PFFRG_JIT_MINBLOCKS=3 PFFRG_JIT_CHUNK=16" "PFFRG_JIT_PREFETCH=4" "PFFRG_JIT_ACC=12 PFFRG_JIT_CHUNK=24"
bash tools/gpu_sweep.sh r1e honeycomb_kitaev_r7_xyz_nw64 "X=1" "PFFRG_JIT_NBT=64 PFFRG_JIT_NB=16" "PFFRG_JIT_NBT=64 PFFRG_JIT_NB=16 PFFRG_JIT_TILES=2" "PFFRG_JIT_NBT=64 PFFRG_JIT_TILES=2"
bash tools/gpu_sweep.sh r1e pyrochlore_r8_su2_nw64 "X=1" "PFFRG_JIT_TILES=2" "PFFRG_JIT_CHUNK=64 PFFRG_JIT_ACC=16"
