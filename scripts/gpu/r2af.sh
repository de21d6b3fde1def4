cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_raymarch.py tests/test_gpu_occupancy.py tests/test_gpu_trainer.py -x -q 2>&1 | grep -v Warn | grep "^E\|passed\|failed\|Error\|^FAILED\|^tests.*py:[0-9]" | head -20
python scripts/gpu/occ_prof.py partial 500 2>&1 | grep -v Warn | tail -17 | head -8
python scripts/gpu/occ_prof.py full 500 2>&1 | grep -v Warn | tail -17 | head -8
