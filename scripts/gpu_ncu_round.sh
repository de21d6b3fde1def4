# ncu evidence for profiles/: (1) launch list of 3 eager training steps, (2) --set full of the heavy kernels of the train step,
# (3) --set full of the big-batch kernels (grid encoder 2^22 points, UMMA deform MLP at 1M samples, frame march/composite)
set -x
TAG=${TAG:-r1}
timeout 600 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log
ncu --set full --clock-control none --import-source on -k "regex:k_wgrad|k_deform_forward_umma|k_grid_backward|k_composite_train_bwd|k_deform_backward|k_march_rays_train_warp|k_adam|k_grid_forward|k_heads" -s 60 -c 14 -o gpurun_out/${TAG}_train -f python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
timeout 300 python -m seald_nerf_b200.microbench ncu > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:k_grid_forward|k_grid_backward|k_deform_forward_umma|k_march_rays_train|k_composite_train|k_packbits" -c 8 -o gpurun_out/${TAG}_big -f python -m seald_nerf_b200.microbench ncu > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
ls -la gpurun_out/
