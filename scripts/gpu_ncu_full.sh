# one --set full capture of the heavy kernels of an eager training step (after the same command ran clean without ncu)
set -x
timeout 600 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:${NCU_KERNELS:-k_wgrad|k_deform_forward|k_grid_backward|k_composite_train_bwd|k_deform_backward}" -s ${NCU_SKIP:-0} -c ${NCU_COUNT:-7} -o gpurun_out/prof -f python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log; ls -la gpurun_out/
