cd $GRAFT_REPO_ROOT
for w in 2 3 4 6; do SEALD_WGRAD_ROWW=$w python scripts/gpu/wgrad_split.py 2>&1 | tail -1; done
timeout 300 python -m pytest tests/test_gpu_umma.py -x -q 2>&1 | grep -v Warn | grep "^E\|passed\|failed\|Error\|^FAILED" | head
