cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_trainer.py -x -q -k "tail or ema" 2>&1 | grep -v Warn | grep "^E\|passed\|failed\|Error" | head -30
