cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_trainer.py tests/test_gpu_field.py tests/test_gpu_ref_parity.py tests/test_gpu_checkpoint.py -x -q 2>&1 | grep -v Warn | grep "^E\|passed\|failed\|Error\|^FAILED" | head -30
timeout 600 python bench.py --steps 300 --warmup 30 --no-extras --no-cpu-baseline > gpurun_out/r2k.log 2> gpurun_out/r2k.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2k.log').read().strip().splitlines()[-1])
print(round(d['value']/1e6,3), round(d['ms_per_step'],4), round(d['e2e']['value']/1e6,3), d['gpu_launches'], d['config'].get('final_loss'))
print(d['roofline']['stage_ms'])
PY
tail -c 300 gpurun_out/r2k.err
