# full GPU check: every gpu-marked test, then the default bench line
set -x
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -25 > gpurun_out/t.log; tail -25 gpurun_out/t.log
timeout 900 python bench.py --steps ${STEPS:-200} --warmup 20 > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
