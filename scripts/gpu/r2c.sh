set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_checkpoint.py tests/test_gpu_train_kernels.py -x -q 2>&1 | tail -25
timeout 900 python bench.py --steps 200 --warmup 20 --no-extras > gpurun_out/r2c_bench.log 2> gpurun_out/r2c_bench.err
echo bench rc $?
tail -c 600 gpurun_out/r2c_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c_bench.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['config'].get('skipped_steps'))
print(d['roofline']['stage_ms'])
PY
