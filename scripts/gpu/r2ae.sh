cd $GRAFT_REPO_ROOT
python scripts/gpu/occ_prof.py partial 0 2>&1 | grep -v Warn | tail -17
python scripts/gpu/occ_prof.py partial 500 2>&1 | grep -v Warn | tail -17
