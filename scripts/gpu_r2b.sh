#!/bin/bash
# Round-2 GPU call B (2 GPUs): multi-device tests, 2-GPU bench with the decomposition-invariance check, C++ app on 2 GPUs.
TAG=${1:-r02b}
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 900 python -m pytest tests/test_gpu_group.py tests/test_host_apps.py -m gpu -q -k "distinct or equals_oracle or several_strips" ) > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -15 gpurun_out/${TAG}_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
echo "bench rc=$?"
tail -c 2500 gpurun_out/${TAG}_bench_n2.json; tail -5 gpurun_out/${TAG}_bench_n2.err
( time timeout 300 lgca_b200/host/bin/lgca-karman --steps 1000 --hash-every 500 --quiet --gpus 2 ) 2>&1 | tail -12
