cd $GRAFT_REPO_ROOT
for i in 1 2 3 4 5 6; do
timeout 300 python -m pytest tests/test_gpu_trainer.py -x -q -k "fused_optimizer_tail" 2>&1 | grep -v Warn | grep "^E  \|passed\|failed" | head -6 | cut -c1-300
done
