cd $GRAFT_REPO_ROOT
export PYTHONFAULTHANDLER=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/r2aq_bench2.log 2> gpurun_out/r2aq_bench2.err
grep -n "Fatal\|File \"/root\|File \".*seald\|Segmentation\|Current thread" -A12 gpurun_out/r2aq_bench2.err | head -60 | cut -c1-200
