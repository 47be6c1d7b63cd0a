#!/bin/bash
# Round-2 GPU call A (1 GPU): full GPU test suite, smoke, bench (both arms).  Artefacts -> gpurun_out/.
TAG=${1:-r02a}
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -12 gpurun_out/${TAG}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
( time timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err ) 2>&1 | tail -3
python scripts/show_bench.py gpurun_out/${TAG}_bench_n1.json 2>/dev/null | head -60 || tail -c 3000 gpurun_out/${TAG}_bench_n1.json
tail -5 gpurun_out/${TAG}_bench_n1.err
