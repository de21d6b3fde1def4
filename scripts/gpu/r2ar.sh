cd $GRAFT_REPO_ROOT
for cfg in "1 0" "0 0" "1 1" "0 1" "1 0"; do
set -- $cfg
SEALD_FUSED_TAIL=$1 SEALD_GRID_BWD_SPLIT=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 300 --warmup 30 --no-extras > gpurun_out/r2ar_bench2.log 2> gpurun_out/r2ar_bench2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2ar_bench2.log').read().strip().splitlines()[-1])
print("tail=$1 split=$2", round(d['value']/1e6,3), round(d['ms_per_step'],4), round(d['e2e']['value']/1e6,3), d['gpu_launches']//300, d['dp_consistent']['ok'])
PY
done
