#!/bin/bash
# Build liblgca_b200 from another git revision into <out.so> for same-box A/B timing:
#   scripts/build_ab.sh HEAD ab_old.so ; LGCA_B200_LIB=$PWD/ab_old.so python scripts/quick_bench.py k4
set -e
REF=${1:-HEAD}; OUT=$(realpath ${2:-ab_old.so}); ROOT=$(cd $(dirname $0)/.. && pwd)
TMP=$(mktemp -d)
git -C $ROOT archive $REF lgca_b200/csrc include | tar -x -C $TMP
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O3 --use_fast_math -shared -o $OUT $TMP/lgca_b200/csrc/*.cu
rm -rf $TMP; echo built $OUT from $REF
