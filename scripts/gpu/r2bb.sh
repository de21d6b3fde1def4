cd $GRAFT_REPO_ROOT
export SEALD_DENSITY_IMPL=split
echo "== baseline"; timeout 300 python scripts/gpu/occ_once.py full 3 2>&1 | tail -2
echo "== G2"; SEALD_UMMA_G=2 timeout 300 python scripts/gpu/occ_once.py full 3 2>&1 | tail -2
echo "== G2 slim"; SEALD_UMMA_G=2 SEALD_SIGMA_SLIM=1 timeout 300 python scripts/gpu/occ_once.py full 3 2>&1 | tail -2
echo "== G2 slim partial"; SEALD_UMMA_G=2 SEALD_SIGMA_SLIM=1 timeout 300 python scripts/gpu/occ_once.py partial 3 2>&1 | tail -2
echo "== slim only"; SEALD_SIGMA_SLIM=1 timeout 300 python scripts/gpu/occ_once.py full 3 2>&1 | tail -2
