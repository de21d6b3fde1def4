set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25
timeout 600 python tests/golden/make_ref_golden.py gpurun_out/golden_ref train > gpurun_out/r2b_parity.log 2> gpurun_out/r2b_parity.err
echo parity rc $?
timeout 900 python bench.py --steps 200 --warmup 20 > gpurun_out/r2b_bench.log 2> gpurun_out/r2b_bench.err
echo bench rc $?
tail -c 600 gpurun_out/r2b_bench.err
