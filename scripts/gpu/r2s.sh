cd $GRAFT_REPO_ROOT
python scripts/gpu/occ_kernels.py 2>&1 | tail -4
