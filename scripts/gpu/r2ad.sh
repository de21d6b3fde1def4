cd $GRAFT_REPO_ROOT
python scripts/gpu/occ_kernels.py 0 2>&1 | tail -4
python scripts/gpu/occ_kernels.py 500 2>&1 | tail -4
