cd $GRAFT_REPO_ROOT
timeout 900 python tests/golden/make_ref_golden.py gpurun_out/golden_ref pretrain 2>&1 | grep -v Warn | tail -8 | cut -c1-1500
timeout 900 python -m pytest tests/test_gpu_seal.py tests/test_gpu_ref_parity.py -x -q 2>&1 | grep -v Warn | grep "^E\|passed\|failed\|Error\|^FAILED\|^tests.*py:[0-9]" | head -30
