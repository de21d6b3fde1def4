cd $GRAFT_REPO_ROOT
N=${NGPU:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r2aa_bench$N.log 2> gpurun_out/r2aa_bench$N.err
echo rc $?
tail -c 500 gpurun_out/r2aa_bench$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2aa_bench$N.log').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value']/1e6,3), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']/1e6,3), d['gpu_launches'], d['dp_consistent']['ok'])
for k in ('frame','seald','occupancy_update'):
    print(k, json.dumps(d.get(k))[:400])
PY
