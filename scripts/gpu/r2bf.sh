cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_renderer.py -m gpu -x -q 2>&1 | grep -v Warn | tail -6
for cfg in "0 4" "1 1" "1 4" "1 6"; do
set -- $cfg
echo "== pack $1 nstep0 $2"
SEALD_RENDER_PACK=$1 SEALD_RENDER_NSTEP0=$2 timeout 300 python scripts/gpu/frame_share.py 1 4 8 2>&1 | tail -3
done
