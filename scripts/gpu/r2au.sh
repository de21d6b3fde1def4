cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_umma.py tests/test_gpu_field.py tests/test_gpu_trainer.py tests/test_gpu_ref_parity.py -x -q 2>&1 | grep -v Warn | grep "^E\|passed\|failed\|Error\|^FAILED\|^tests.*py:[0-9]" | head -30
python scripts/gpu/deform_small.py 2>&1 | tail -4
timeout 900 python bench.py --steps 300 --warmup 30 --no-extras --no-cpu-baseline > gpurun_out/r2au_bench.log 2> gpurun_out/r2au_bench.err
tail -c 300 gpurun_out/r2au_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2au_bench.log').read().strip().splitlines()[-1])
print(round(d['value']/1e6,3), round(d['ms_per_step'],4), round(d['e2e']['value']/1e6,3), d['gpu_launches'], d['config']['final_loss'])
print(d['roofline']['stage_ms'])
PY
