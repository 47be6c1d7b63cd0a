#!/bin/bash
# Round-1 (re-entry) GPU call 2: parity on the new default build (31-LOP3 network, strided plane addressing, SHF),
# parity of the loads-only variant, then same-box A/B timings.
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu --maxfail=5 -q ) > gpurun_out/c2_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c2_tests.log
tail -4 gpurun_out/c2_tests.log
( LGCA_B200_LIB=$PWD/ab_ldonly.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu --maxfail=5 -q ) > gpurun_out/c2_tests_ldonly.log 2>&1
tail -2 gpurun_out/c2_tests_ldonly.log
for lib in default ab_ldonly.so ab_shf.so; do
  for k in k5 k6; do
    if [ $lib = default ]; then
      echo "== default $k"; timeout 120 python scripts/quick_bench.py $k
    else
      echo "== $lib $k"; LGCA_B200_LIB=$PWD/$lib timeout 120 python scripts/quick_bench.py $k
    fi
  done
done > gpurun_out/c2_ab.log 2>&1
cat gpurun_out/c2_ab.log
