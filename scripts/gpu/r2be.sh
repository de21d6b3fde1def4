cd $GRAFT_REPO_ROOT
for mult in 1 2 4 8; do
echo "== slots x$mult"
SEALD_RENDER_SLOTS_MULT=$mult timeout 300 python scripts/gpu/frame_share.py 1 4 8 2>&1 | tail -3
done
