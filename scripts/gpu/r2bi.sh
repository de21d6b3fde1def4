cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_renderer.py tests/test_gpu_ref_parity.py tests/test_gpu_seal.py -m gpu -x -q 2>&1 | grep -v Warn | tail -6
timeout 300 python scripts/gpu/frame_share.py 1 8 2>&1 | tail -2
