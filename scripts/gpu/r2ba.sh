cd $GRAFT_REPO_ROOT
timeout 300 python scripts/gpu/density_bench.py 2>&1 | tail -6
SEALD_UMMA_G=2 timeout 300 python scripts/gpu/density_bench.py 2>&1 | grep umma
SEALD_UMMA_G=4 timeout 300 python scripts/gpu/density_bench.py 2>&1 | grep umma
