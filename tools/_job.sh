python -m pytest tests -m gpu -q > gpurun_out/r1w_pytest_gpu.log 2>&1; tail -4 gpurun_out/r1w_pytest_gpu.log
for wl in cubic_r7_su2_nw64 honeycomb_kitaev_r7_xyz_nw64 pyrochlore_r8_su2_nw64 square_r4_su2_nw32; do bash tools/gpu_sweep.sh r1w $wl "X=1"; done
