cd $GRAFT_REPO_ROOT
for cfg in "0 296" "start 296" "start 148" "start 592" "after_march 296" "after_march 592"; do
set -- $cfg
SEALD_DEFER=$1 SEALD_ADAM_BLOCKS=$2 timeout 600 python bench.py --steps 300 --warmup 30 --no-extras --no-cpu-baseline > gpurun_out/r2j.log 2> gpurun_out/r2j.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2j.log').read().strip().splitlines()[-1])
print("$cfg", round(d['value']/1e6,3), round(d['ms_per_step'],4), round(d['e2e']['value']/1e6,3), d['gpu_launches'])
PY
done
