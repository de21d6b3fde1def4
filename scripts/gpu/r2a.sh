set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python tests/golden/make_ref_golden.py gpurun_out/golden_ref > gpurun_out/r2a_parity.log 2> gpurun_out/r2a_parity.err
echo parity rc $?
timeout 600 python oracle/ref_bench.py --steps 100 --warmup 20 > gpurun_out/r2a_refbench.log 2> gpurun_out/r2a_refbench.err
echo refbench rc $?
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/r2a_bench.log 2> gpurun_out/r2a_bench.err
echo bench rc $?
tail -c 3000 gpurun_out/r2a_parity.log; tail -c 1500 gpurun_out/r2a_parity.err; tail -c 3000 gpurun_out/r2a_refbench.log; tail -c 1500 gpurun_out/r2a_refbench.err
