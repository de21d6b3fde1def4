cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:k_grid_forward -s 2 -c 1 -f -o gpurun_out/r2_grid_occ_xfast python scripts/gpu/occ_kernels.py > gpurun_out/ncu_grid.log 2>&1
tail -3 gpurun_out/ncu_grid.log
ncu -i gpurun_out/r2_grid_occ_xfast.ncu-rep --page raw --csv > gpurun_out/r2_grid_occ_xfast_raw.csv 2>/dev/null
ls -la gpurun_out/r2_grid_occ_xfast*
