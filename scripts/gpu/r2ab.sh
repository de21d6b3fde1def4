cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_field.py tests/test_gpu_trainer.py tests/test_gpu_ref_parity.py tests/test_gpu_renderer.py tests/test_gpu_occupancy.py tests/test_gpu_seal.py -x -q 2>&1 | grep -v Warn | grep "^E\|passed\|failed\|Error\|^FAILED\|^tests.*py:[0-9]" | head -20
for impl in tile warp; do
SEALD_HEADS_IMPL=$impl python - <<'PY'
import torch, json, os
from seald_nerf_b200 import microbench
dev=torch.device('cuda:0')
for lm in (20, 15):
    r=microbench.field_throughput(dev, log2_M=lm, reps=10)
    print(os.environ['SEALD_HEADS_IMPL'], lm, r["heads_fwd"], flush=True)
PY
done
python scripts/gpu/occ_kernels.py 2>&1 | head -1
timeout 900 python bench.py --steps 300 --warmup 30 --no-ref-gpu --no-cpu-baseline > gpurun_out/r2ab_bench.log 2> gpurun_out/r2ab_bench.err
tail -c 300 gpurun_out/r2ab_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ab_bench.log').read().strip().splitlines()[-1])
print(round(d['value']/1e6,3), round(d['ms_per_step'],4), round(d['e2e']['value']/1e6,3), d['gpu_launches'])
print(d['roofline']['stage_ms'])
for k in ('frame','occupancy_update','seald'):
    print(k, json.dumps(d.get(k))[:400])
PY
