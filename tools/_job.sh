python -m pytest tests -m gpu -q > gpurun_out/r2a_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2a_pytest_gpu.log
for wl in cubic_r7_su2_nw64 honeycomb_kitaev_r7_xyz_nw64 square_r4_su2_nw32; do PFFRG_JIT_VERBOSE=1 bash tools/gpu_sweep.sh r2a $wl "PFFRG_AUTOTUNE=1"; grep autotune gpurun_out/r2a_sweep.err | tail -4; done
