# launch list of OUR kernels for 3 eager (non-graph) training steps; per-launch times are cold-cache and serialised
set -x
timeout 600 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
