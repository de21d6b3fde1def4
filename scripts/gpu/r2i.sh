cd $GRAFT_REPO_ROOT
python scripts/gpu/stages.py march > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_march_rays_train_chain -s 10 -c 1 -f -o gpurun_out/r2_march_chain python scripts/gpu/stages.py march > gpurun_out/ncu_march.log 2>&1
tail -1 gpurun_out/ncu_march.log
