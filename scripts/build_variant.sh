#!/bin/bash
# Build liblgca_b200 from the WORKING TREE with extra nvcc flags into <out.so> for same-box A/B timing:
#   scripts/build_variant.sh ab_shf.so -DLGCA_FMA_SHIFT=0 ; LGCA_B200_LIB=$PWD/ab_shf.so python scripts/quick_bench.py k5
set -e
OUT=$(realpath $1); shift; ROOT=$(cd $(dirname $0)/.. && pwd)
cd $ROOT && LGCA_B200_OUT=$OUT LGCA_B200_NVCC_EXTRA="$*" python -m lgca_b200.build --force > /dev/null
echo built $OUT with "$@"
