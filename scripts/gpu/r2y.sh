cd $GRAFT_REPO_ROOT
for lim in 65536 1000000; do
SEALD_MARCH_CHAIN_MAX=$lim python - <<'PY'
import torch, json, os
from seald_nerf_b200 import microbench
r = microbench.march_composite(torch.device('cuda:0'))
print(os.environ['SEALD_MARCH_CHAIN_MAX'], json.dumps({k: r[k] for k in r if 'march' in k})[:600], flush=True)
PY
done
