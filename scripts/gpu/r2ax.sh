cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_train_kernels.py tests/test_gpu_checkpoint.py tests/test_gpu_ref_parity.py -m gpu -x -q 2>&1 | tail -5
for ff in 1 0 1 0; do
SEALD_FLAGS_FINAL=$ff timeout 600 python bench.py --steps 300 --warmup 30 --no-extras --no-cpu-baseline > gpurun_out/r2ax.log 2> gpurun_out/r2ax.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2ax.log').read().strip().splitlines()[-1])
print("flags_final=$ff", round(d['value']/1e6,3), round(d['ms_per_step'],4), round(d['e2e']['value']/1e6,3), d['gpu_launches'])
PY
done
