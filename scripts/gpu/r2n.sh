cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py > gpurun_out/dp_check_r2.log 2>&1
grep -v Warn gpurun_out/dp_check_r2.log | grep -B2 -A12 "Traceback\|OK\|MISMATCH" | head -60
