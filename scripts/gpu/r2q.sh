cd $GRAFT_REPO_ROOT
K='regex:k_occ_cell|k_occ_partial|k_occ_store|k_occ_ema|k_deform_forward_umma|k_grid_forward|k_sigma_forward|distribution|cumsum|scan|k_packbits'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 600 --csv --log-file gpurun_out/r2_occ_full_launches.csv python scripts/gpu/occ_once.py full 1 > gpurun_out/ncu_occ.log 2>&1
tail -2 gpurun_out/ncu_occ.log
python scripts/agg_launches.py gpurun_out/r2_occ_full_launches.csv 0.2 | head -30
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 900 --csv --log-file gpurun_out/r2_occ_partial_launches.csv python scripts/gpu/occ_once.py partial 1 > gpurun_out/ncu_occ2.log 2>&1
python scripts/agg_launches.py gpurun_out/r2_occ_partial_launches.csv 0.2 | head -30
