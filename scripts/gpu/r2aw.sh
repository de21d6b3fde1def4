cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_checkpoint.py tests/test_gpu_rays.py -x -q 2>&1 | grep -v Warn | grep "^E\|passed\|failed\|Error\|^FAILED\|^tests.*py:[0-9]" | head -10
for c in 0 12 12 0; do
SEALD_GRAPH_CONDITION=$c timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2aw_bench.log 2> gpurun_out/r2aw_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2aw_bench.log').read().strip().splitlines()[-1])
print("condition=$c", round(d['value']/1e6,3), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']/1e6,3), d['gpu_launches'], d['config']['skipped_steps'], d['config']['final_loss'])
PY
done
tail -c 300 gpurun_out/r2aw_bench.err
