#!/bin/bash
# chained launches: parity first, then A-B timing
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_chain.py -x -q ) > gpurun_out/r03b_chain_tests.log 2>&1
echo "chain tests rc=$?" >> gpurun_out/r03b_chain_tests.log
tail -5 gpurun_out/r03b_chain_tests.log
( time timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "k1 or k2 or k3 or k4 or k6 or default or full_size" ) > gpurun_out/r03b_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/r03b_parity.log
tail -4 gpurun_out/r03b_parity.log
timeout 300 python scripts/chain_ab.py > gpurun_out/r03b_chain_ab.log 2>&1
cat gpurun_out/r03b_chain_ab.log
