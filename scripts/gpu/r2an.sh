cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v Warn | grep "^E\|passed\|failed\|Error\|^FAILED\|^tests.*py:[0-9]" | head -30
for f in 0 1; do
SEALD_FUSE_GRID_HEADS=$f timeout 900 python bench.py --steps 300 --warmup 30 --no-ref-gpu --no-cpu-baseline > gpurun_out/r2an_bench$f.log 2> gpurun_out/r2an_bench.err
tail -c 300 gpurun_out/r2an_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2an_bench$f.log').read().strip().splitlines()[-1])
print("fuse=$f", round(d['value']/1e6,3), round(d['ms_per_step'],4), round(d['e2e']['value']/1e6,3), d['gpu_launches'], d['config']['final_loss'])
print(d['roofline']['stage_ms'])
for k in ('frame','occupancy_update'):
    print(k, json.dumps(d.get(k))[:330])
PY
done
