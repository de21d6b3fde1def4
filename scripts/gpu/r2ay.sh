cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_train_kernels.py tests/test_gpu_checkpoint.py tests/test_gpu_ref_parity.py tests/test_gpu_umma.py -m gpu -q 2>&1 | tail -8
