cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_renderer.py tests/test_gpu_ref_parity.py -m gpu -x -q 2>&1 | grep -v Warn | tail -4
timeout 300 python scripts/gpu/frame_share.py 1 8 2>&1 | tail -2
SEALD_RENDER_COARSE=0 timeout 300 python scripts/gpu/frame_share.py 1 8 2>&1 | tail -2
SEALD_RENDER_PACK=1 timeout 200 python scripts/gpu/frame_stages.py 2>&1 | grep "march\|sum"
