cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py > gpurun_out/dp_check_r2.log 2>&1
grep -v Warn gpurun_out/dp_check_r2.log | grep "OK\|MISMATCH\|Error" | cut -c1-200 | head -14
for cfg in 1 0 1 0; do
SEALD_DP_TAIL=$cfg timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 300 --warmup 30 --no-extras > gpurun_out/r2ar_bench2.log 2> gpurun_out/r2ar_bench2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2ar_bench2.log').read().strip().splitlines()[-1])
print("dp_tail=$cfg", round(d['value']/1e6,3), round(d['ms_per_step'],4), round(d['e2e']['value']/1e6,3), d['gpu_launches']//300, d['dp_consistent']['ok'], d['config']['skipped_steps'])
PY
done
