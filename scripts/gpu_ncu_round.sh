# ncu evidence for profiles/: (1) launch list of 3 eager training steps, (2) --set full of the kernels of one eager train step,
# (3) --set full of the big-batch kernels (grid encoder 2^22 points, UMMA deform MLP at 1M samples, frame march/composite).
# The .ncu-rep files stay on the box unless small; the raw-page CSVs (what profiles/ summarises) come back.
set -x
TAG=${TAG:-r1}
OUT=gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > $OUT/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > $OUT/ncu1.log 2>&1
tail -2 $OUT/ncu1.log
# the training step launches ~21 of our kernels; skip the 3 warm-up steps + calibration and capture one whole step
ncu --set full --clock-control none -k regex:^k_ -s ${NCU_SKIP:-150} -c ${NCU_COUNT:-30} -o /tmp/${TAG}_train -f python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > $OUT/ncu2.log 2>&1
tail -2 $OUT/ncu2.log
ncu -i /tmp/${TAG}_train.ncu-rep --page raw --csv > $OUT/${TAG}_train_raw.csv 2>/dev/null
timeout 300 python -m seald_nerf_b200.microbench ncu > $OUT/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:k_grid_forward|k_grid_scatter|k_grid_input_backward|k_deform_forward_umma|k_march_rays_train|k_composite_train|k_packbits" -c 14 -o /tmp/${TAG}_big -f python -m seald_nerf_b200.microbench ncu > $OUT/ncu3.log 2>&1
tail -2 $OUT/ncu3.log
ncu -i /tmp/${TAG}_big.ncu-rep --page raw --csv > $OUT/${TAG}_big_raw.csv 2>/dev/null
ls -la /tmp/*.ncu-rep
for f in /tmp/${TAG}_train.ncu-rep /tmp/${TAG}_big.ncu-rep; do if [ $(stat -c %s $f) -lt 20000000 ]; then cp $f $OUT/; fi; done
ls -la $OUT/
