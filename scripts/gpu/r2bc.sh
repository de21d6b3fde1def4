cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_umma.py tests/test_gpu_occupancy.py tests/test_gpu_field.py -m gpu -x -q 2>&1 | tail -12
