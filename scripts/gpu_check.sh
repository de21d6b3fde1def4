# quick GPU check: selected parity tests + a short bench
set -x
timeout 900 python -m pytest ${TESTS:-tests/test_gpu_field.py tests/test_gpu_trainer.py tests/test_gpu_grid.py} -m gpu -q --tb=short -x 2>&1 | tail -15
timeout 600 python bench.py --steps ${STEPS:-100} --warmup 10 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -2 gpurun_out/bench.log
