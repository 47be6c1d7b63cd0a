#!/bin/bash
TAG=${1:-r02e}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -k regex:step_resident -c 1 -f"
timeout 300 $NCU -o gpurun_out/${TAG}_res_c1 python scripts/prof_one.py FHP_I 1400 700 pipe 0 0 400 > gpurun_out/${TAG}_ncu_c1.log 2>&1
timeout 300 $NCU -o gpurun_out/${TAG}_res_karman python scripts/prof_one.py FHP_III 4400 2200 karman 0 0 200 > gpurun_out/${TAG}_ncu_karman.log 2>&1
timeout 300 $NCU -o gpurun_out/${TAG}_res_hpp python scripts/prof_one.py HPP 4096 4096 periodic 0 0 200 > gpurun_out/${TAG}_ncu_hpp.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_c1.log
ls -la gpurun_out/*.ncu-rep
