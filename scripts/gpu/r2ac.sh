cd $GRAFT_REPO_ROOT
python scripts/gpu/occ_once.py full 3 0 2>&1 | tail -3
python scripts/gpu/occ_once.py full 3 500 2>&1 | tail -4
python scripts/gpu/occ_once.py partial 3 500 2>&1 | tail -3
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active --format=csv
