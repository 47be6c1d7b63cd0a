#!/bin/bash
# Round-1 (re-entry) GPU call 3: parity on the default build (snapshot mutex, irregular-width path restored), then
# occupancy A/B: K=6 capped at 96 registers, K=4 at 80 registers, wall variants capped.
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu --maxfail=5 -q ) > gpurun_out/c3_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c3_tests.log
tail -4 gpurun_out/c3_tests.log
{
  echo "== default k5"; timeout 120 python scripts/quick_bench.py k5
  echo "== default k6"; timeout 120 python scripts/quick_bench.py k6
  echo "== ab_k6cap.so k6"; LGCA_B200_LIB=$PWD/ab_k6cap.so timeout 120 python scripts/quick_bench.py k6
  echo "== ab_k4_80.so k4"; LGCA_B200_LIB=$PWD/ab_k4_80.so timeout 120 python scripts/quick_bench.py k4
  echo "== ab_nscap.so k5"; LGCA_B200_LIB=$PWD/ab_nscap.so timeout 120 python scripts/quick_bench.py k5
  echo "== ab_nscap.so k6"; LGCA_B200_LIB=$PWD/ab_nscap.so timeout 120 python scripts/quick_bench.py k6
} > gpurun_out/c3_ab.log 2>&1
cat gpurun_out/c3_ab.log
