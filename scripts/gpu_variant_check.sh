#!/bin/bash
# usage: gpu_variant_check.sh <variant> : parity subset + bench + trace with multiply_b200/_variants/lib_<variant>.so
cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
v=$1
export MP_LIB=$GRAFT_REPO_ROOT/multiply_b200/_variants/lib_$v.so
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "implicit or forward_vs_oracle or golden or precision or sdf_grid or background" 2>&1 | tail -5
