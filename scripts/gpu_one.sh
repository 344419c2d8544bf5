cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_gpu_mirror.py -m gpu -q -x -k "sdf_func" 2>&1 | tail -15
