cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_field.py tests/test_gpu_trainer.py tests/test_gpu_renderer.py tests/test_gpu_raymarch.py tests/test_gpu_ref_parity.py tests/test_gpu_seal.py -x -q 2>&1 | grep -v Warn | grep "^E\|passed\|failed\|Error\|^tests" | head -30
python - <<'PY'
import torch, json
from seald_nerf_b200 import microbench
dev=torch.device('cuda:0')
for lm in (20, 17, 15):
    r=microbench.field_throughput(dev, log2_M=lm, reps=10)
    print(lm, r["deform_fwd_tcgen05"], r["heads_fwd"], flush=True)
print(json.dumps(microbench.frame_render(dev)), flush=True)
PY
timeout 900 python bench.py --steps 200 --warmup 20 --no-extras > gpurun_out/r2e_bench.log 2> gpurun_out/r2e_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2e_bench.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['config'].get('skipped_steps'))
print(d['roofline']['stage_ms'])
PY
tail -c 400 gpurun_out/r2e_bench.err
