# ncu evidence of the training step (eager launches, one GPU): (1) per-launch durations of ~20 steps, (2) --set full of two whole steps.
# The set-up kernels (packbits, per-frame occupancy boxes) are filtered out by name so the skip count lands inside the steps.
set -x
TAG=${TAG:-r1}
OUT=gpurun_out
K='regex:k_select_frame|k_march|k_deform|k_grid|k_heads|k_composite|k_wgrad|k_adam|k_grad_finite|k_cast_pad|k_pack_umma|k_loss_scale|k_mse'
timeout 600 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > $OUT/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > $OUT/ncu1.log 2>&1
tail -2 $OUT/ncu1.log
ncu --set full --clock-control none -k "$K" -s ${NCU_SKIP:-60} -c ${NCU_COUNT:-45} -o /tmp/${TAG}_train -f python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > $OUT/ncu2.log 2>&1
tail -2 $OUT/ncu2.log
ncu -i /tmp/${TAG}_train.ncu-rep --page raw --csv > $OUT/${TAG}_train_raw.csv 2>/dev/null
ls -la $OUT/ | tail -8
