cd $GRAFT_REPO_ROOT
for cfg in "1 8" "1 16" "1 32" "2 8" "2 16" "2 32" "4 32"; do
set -- $cfg
SEALD_RENDER_SLOTS_MULT=$1 SEALD_RENDER_MAX_NSTEP=$2 python - <<PY
import torch, json
from seald_nerf_b200 import microbench
r = microbench.frame_render(torch.device('cuda:0'), times=(0.5,), reps=5)
print("$cfg", r["t=0.50"], flush=True)
PY
done
