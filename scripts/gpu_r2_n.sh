cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
mkdir -p gpurun_out
bash scripts/gpu_r2_l.sh vJ vK vL
MP_LIB=$GRAFT_REPO_ROOT/multiply_b200/_variants/lib_vL.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
MP_TC_RZ_SCALE=1.0 timeout 300 python scripts/gpu_normal_diag.py 2>&1 | grep "^surface tc\|^origin tc\|^tc rendered"
