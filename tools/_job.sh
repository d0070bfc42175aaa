bash tools/gpu_sweep.sh r1u honeycomb_kitaev_r7_xyz_nw64 "X=1" "PFFRG_PAD_GROUPS=1" "PFFRG_PAD_GROUPS=1 PFFRG_THREADS=512 PFFRG_JIT_MINBLOCKS=1"
bash tools/gpu_sweep.sh r1u kagome_dm_r7_tri_nw64 "PFFRG_PAD_GROUPS=1"
bash tools/gpu_sweep.sh r1u square_r4_su2_nw32 "X=1" "PFFRG_PAD_GROUPS=1"
