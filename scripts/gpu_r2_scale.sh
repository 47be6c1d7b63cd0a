#!/bin/bash
# Round-2 scaling call (8 GPUs): bench at N = 8 (and 4, 2 when asked) with the decomposition-invariance check and the C4 box extra;
# lgca-karman on 2 GPUs, lgca-box on 8.
TAG=${1:-r02y}; NS=${2:-"8 4"}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in $NS; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
  echo "bench N=$N rc=$?"; python scripts/show_bench.py gpurun_out/${TAG}_bench_n$N.json; tail -3 gpurun_out/${TAG}_bench_n$N.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/${TAG}_bench_n$N.json').read().strip().splitlines()[-1])
print('   launch_ms_rank_min_max', d['roofline'].get('launch_ms_rank_min_max'), 'e2e probe', d['e2e'].get('host_link_probe'))"
done
( time timeout 300 lgca_b200/host/bin/lgca-karman --steps 1000 --hash-every 500 --quiet --gpus 2 ) 2>&1 | tail -7
( time timeout 600 lgca_b200/host/bin/lgca-box --dims 65536 32768 --model FHP_II --gpus 8 --steps 120 --pp-interval 60 --no-cell-fields --quiet ) 2>&1 | tail -6
