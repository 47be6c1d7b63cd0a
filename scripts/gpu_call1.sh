#!/bin/bash
# Round-1 (re-entry) GPU call 1: parity on the new default build, then same-box A/B timings.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
( time timeout 420 python -m pytest tests -m gpu --maxfail=5 -q ) > gpurun_out/c1_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c1_tests.log
tail -3 gpurun_out/c1_tests.log
for lib in default ab_head.so ab_shf.so ab_mb0.so; do
  for k in k5 k6 k4; do
    if [ $lib = default ]; then
      echo "== default $k"; timeout 120 python scripts/quick_bench.py $k
    else
      echo "== $lib $k"; LGCA_B200_LIB=$PWD/$lib timeout 120 python scripts/quick_bench.py $k
    fi
  done
done > gpurun_out/c1_ab.log 2>&1
# chunk-height / occupancy-model sensitivity of the default build on C5
for cr in 120 160 200 244 328; do
  echo "== default k5 chunk_rows=$cr"; LGCA_B200_CHUNK_ROWS=$cr timeout 120 python scripts/quick_bench.py k5 2>&1 | grep "32768x32768"
done >> gpurun_out/c1_ab.log 2>&1
cat gpurun_out/c1_ab.log
