cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py > gpurun_out/dp_check_r2.log 2>&1
echo rc $?
grep -v Warn gpurun_out/dp_check_r2.log | grep "OK\|MISMATCH\|Error" | cut -c1-250 | head -16
timeout 600 python -m pytest tests/test_gpu_renderer.py tests/test_gpu_ref_parity.py -m gpu -x -q 2>&1 | grep -v Warn | tail -3
python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from seald_nerf_b200 import microbench
r = microbench.frame_render(torch.device("cuda:0"), times=(0.5,), reps=5)
print("1-GPU frame", r)
PY
