cd $GRAFT_REPO_ROOT
for g in 2 4; do SEALD_UMMA_G=$g python scripts/gpu/deform_small.py 2>&1 | grep "^G"; done
