cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_trainer.py -x -q -k "pipelined" 2>&1 | grep -v Warn | grep "^E\|passed\|failed\|Error\|^FAILED\|^tests.*py:[0-9]" | head -20
timeout 900 python bench.py --steps 300 --warmup 30 --no-extras --no-cpu-baseline > gpurun_out/r2z_bench.log 2> gpurun_out/r2z_bench.err
tail -c 300 gpurun_out/r2z_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench.log').read().strip().splitlines()[-1])
print(round(d['value']/1e6,3), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']/1e6,3), round(d['e2e']['ms_per_step'],4), 'sync', round(d['e2e']['synchronous']['value']/1e6,3))
PY
