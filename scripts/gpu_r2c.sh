#!/bin/bash
TAG=${1:-r02c}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "res or extreme or canonical" ) > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -8 gpurun_out/${TAG}_tests.log
timeout 600 python scripts/small_lattices.py ${2:-0 4 6} > gpurun_out/${TAG}_small.md 2>&1
cat gpurun_out/${TAG}_small.md
