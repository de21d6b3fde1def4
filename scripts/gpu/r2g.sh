cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v Warn | grep "^E\|passed\|failed\|Error\|^tests\|^FAILED" | head -60
timeout 900 python bench.py --steps 200 --warmup 20 > gpurun_out/r2g_bench.log 2> gpurun_out/r2g_bench.err
echo bench rc $?
tail -c 600 gpurun_out/r2g_bench.err
