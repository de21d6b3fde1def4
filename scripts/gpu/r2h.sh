cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_raymarch.py tests/test_gpu_trainer.py tests/test_gpu_seal.py tests/test_gpu_ref_parity.py -x -q 2>&1 | grep -v Warn | grep "^E\|passed\|failed\|Error\|^FAILED" | head -30
for c in 0 1; do
SEALD_MARCH_CHAIN=$c timeout 900 python bench.py --steps 200 --warmup 20 --no-extras > gpurun_out/r2h_bench_$c.log 2> gpurun_out/r2h_bench_$c.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2h_bench_$c.log').read().strip().splitlines()[-1])
print($c, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['config'].get('skipped_steps'))
print(d['roofline']['stage_ms'])
PY
done
