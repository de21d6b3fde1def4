# narrow `ncu --set full` capture of a few kernels of the eager training step (cheap re-profile after a kernel changed)
set -x
TAG=${TAG:-r1d}
OUT=gpurun_out
K=${NCU_KERNELS:-regex:k_adam|k_grid_scatter}
timeout 300 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > $OUT/plain.log 2>&1 && \
ncu --set full --clock-control none -k "$K" -s ${NCU_SKIP:-12} -c ${NCU_COUNT:-6} -o /tmp/${TAG}_narrow -f python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > $OUT/ncu_narrow.log 2>&1
tail -2 $OUT/ncu_narrow.log
ncu -i /tmp/${TAG}_narrow.ncu-rep --page raw --csv > $OUT/${TAG}_narrow_raw.csv 2>/dev/null
ls -la $OUT/${TAG}_narrow_raw.csv
