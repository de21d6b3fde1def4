cd $GRAFT_REPO_ROOT
for w in 8 12; do SEALD_WGRAD_ROWW=$w python scripts/gpu/wgrad_split.py 2>&1 | tail -1; done
timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_trainer.py tests/test_gpu_field.py tests/test_gpu_ref_parity.py tests/test_gpu_checkpoint.py tests/test_gpu_seal.py -x -q 2>&1 | grep -v Warn | grep "^E\|passed\|failed\|Error\|^FAILED" | head -30
for w in 6 10; do
SEALD_WGRAD_ROWW=$w timeout 600 python bench.py --steps 300 --warmup 30 --no-extras --no-cpu-baseline > gpurun_out/r2m_$w.log 2> gpurun_out/r2m.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2m_$w.log').read().strip().splitlines()[-1])
print($w, round(d['value']/1e6,3), round(d['ms_per_step'],4), round(d['e2e']['value']/1e6,3), d['gpu_launches'], d['config'].get('final_loss'))
print(d['roofline']['stage_ms'])
PY
done
tail -c 300 gpurun_out/r2m.err
