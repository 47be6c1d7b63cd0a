#!/bin/bash
# Build liblgca_b200 from the WORKING TREE with extra nvcc flags into <out.so> for same-box A/B timing:
#   scripts/build_variant.sh ab_shf.so -DLGCA_FMA_SHIFT=0 ; LGCA_B200_LIB=$PWD/ab_shf.so python scripts/quick_bench.py k5
set -e
OUT=$(realpath $1); shift; ROOT=$(cd $(dirname $0)/.. && pwd)
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O3 --use_fast_math -shared "$@" -o $OUT $ROOT/lgca_b200/csrc/*.cu
echo built $OUT with "$@"
