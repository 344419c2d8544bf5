#!/bin/bash
# usage: scripts/build_variant.sh <name> [nvcc defines...]   e.g.  scripts/build_variant.sh poly -DMP_SP_MODE=2
# Builds multiply_b200/_variants/lib_<name>.so = the current objects of multiply_b200/_build with mlp_tc.cu recompiled
# under the given defines; select it at run time with MP_LIB=<path> (scripts/gpu_variants.sh does that on the GPU box).
set -e
cd "$(dirname "$0")/.."
python -m multiply_b200.build > /dev/null
name=$1; shift
mkdir -p multiply_b200/_variants /tmp/mpvar_$name
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -DMP_BUILDING "$@" \
     -c multiply_b200/csrc/mlp_tc.cu -o /tmp/mpvar_$name/mlp_tc.o 2>&1 | grep -v "warning\|\^\|^$\|const int tid" || true
objs=$(ls multiply_b200/_build/*.o | grep -v mlp_tc.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o multiply_b200/_variants/lib_$name.so $objs /tmp/mpvar_$name/mlp_tc.o -lcudart -lcuda
cuobjdump --dump-resource-usage multiply_b200/_variants/lib_$name.so 2>/dev/null | grep -A1 "tc_chain" | grep -o "REG:[0-9]*\|STACK:[0-9]*" | paste - - | head -2
