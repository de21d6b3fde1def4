cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_grid.py tests/test_gpu_occupancy.py -x -q 2>&1 | grep -v Warn | grep "^E\|passed\|failed\|Error\|^FAILED\|^tests.*py:[0-9]" | head -20
python scripts/gpu/occ_kernels.py 2>&1 | tail -4
python - <<'PY'
import torch, json
from seald_nerf_b200 import microbench
dev=torch.device('cuda:0')
print(json.dumps(microbench.grid_throughput(dev)) if hasattr(microbench,'grid_throughput') else [n for n in dir(microbench) if not n.startswith('_')])
PY
