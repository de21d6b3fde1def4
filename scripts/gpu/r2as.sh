cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v Warn | grep "^E\|passed\|failed\|Error\|^FAILED\|^tests.*py:[0-9]" | head -30
timeout 900 python bench.py --steps 300 --warmup 30 > gpurun_out/r2as_bench.log 2> gpurun_out/r2as_bench.err
echo bench rc $?
tail -c 300 gpurun_out/r2as_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2as_bench.log').read().strip().splitlines()[-1])
print(round(d['value']/1e6,3), round(d['ms_per_step'],4), round(d['e2e']['value']/1e6,3), d['gpu_launches'], d['config']['skipped_steps'])
print(d['roofline']['kernel'], round(d['roofline']['frac'],3), d['roofline']['stage_ms'])
for k in ('frame','occupancy_update','seald','hashgrid','hashgrid_4d','e2e_resident_dataset'):
    print(k, json.dumps(d.get(k))[:420])
print(json.dumps(d.get('ref_gpu',{}).get('ours_over_ref')))
PY
