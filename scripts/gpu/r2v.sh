cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_grid.py tests/test_gpu_occupancy.py tests/test_gpu_trainer.py tests/test_gpu_ref_parity.py -x -q 2>&1 | grep -v Warn | grep "^E\|passed\|failed\|Error\|^FAILED\|^tests.*py:[0-9]" | head -20
timeout 900 python bench.py --steps 200 --warmup 20 --no-ref-gpu --no-cpu-baseline > gpurun_out/r2v_bench.log 2> gpurun_out/r2v_bench.err
echo bench rc $?
tail -c 300 gpurun_out/r2v_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_bench.log').read().strip().splitlines()[-1])
print(round(d['value']/1e6,3), round(d['ms_per_step'],4), round(d['e2e']['value']/1e6,3), d['gpu_launches'])
print(d['roofline']['stage_ms'])
for k in ('frame','occupancy_update','hashgrid','hashgrid_4d'):
    print(k, json.dumps(d.get(k))[:700])
PY
