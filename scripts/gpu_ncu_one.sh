#!/bin/bash
# one ncu --set full capture of a kernel: scripts/gpu_ncu_one.sh <tag> <kernel regex> <prof_one.py args...>
TAG=$1; KREGEX=$2; shift 2
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -c 1 -f -o gpurun_out/$TAG python scripts/prof_one.py "$@" > gpurun_out/${TAG}.log 2>&1
tail -2 gpurun_out/${TAG}.log; ls -la gpurun_out/$TAG.ncu-rep
