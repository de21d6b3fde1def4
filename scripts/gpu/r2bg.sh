cd $GRAFT_REPO_ROOT
for cfg in "1 4" "1.5 4" "2 4" "2 8" "3 8"; do
set -- $cfg
echo "== slots x$1 nstep0 $2"
SEALD_RENDER_SLOTS_MULT=$1 SEALD_RENDER_NSTEP0=$2 timeout 300 python scripts/gpu/frame_share.py 1 4 8 2>&1 | tail -3
done
