#!/bin/bash
# Round-end GPU call: ncu captures of the fused-step kernel (C5 and C3 at the default K), full GPU test suite,
# bench.py (both arms), ncu launch list of the bench command, five-config table.  Artefacts -> gpurun_out/.
TAG=${1:-r01b}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -k regex:step_wave --launch-skip 3 -c 1 -f"
timeout 240 $NCU -o gpurun_out/${TAG}_wave_c5 python scripts/prof_one.py FHP_III 32768 32768 periodic 0 0 30 > gpurun_out/${TAG}_ncu_c5.log 2>&1
timeout 240 $NCU -o gpurun_out/${TAG}_wave_c3 python scripts/prof_one.py FHP_III 16384 8192 karman 0 0 30 > gpurun_out/${TAG}_ncu_c3.log 2>&1
K=$(python -c "import lgca_b200; e=lgca_b200.Engine('FHP_III',32768,32768); print(e.info().k_fuse)")
python scripts/summarize_ncu.py full gpurun_out/${TAG}_wave_c5.ncu-rep profiles/${TAG}_wave_k${K}_c5_full.md periodic_k${K} > /dev/null
python scripts/summarize_ncu.py full gpurun_out/${TAG}_wave_c3.ncu-rep profiles/${TAG}_wave_k${K}_c3_full.md karman_k${K} > /dev/null
cp profiles/traffic.json profiles/${TAG}_wave_k${K}_c5_full.md profiles/${TAG}_wave_k${K}_c3_full.md gpurun_out/
( time timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
timeout 400 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
cat gpurun_out/${TAG}_bench_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-karman > gpurun_out/${TAG}_launches_bench.log 2>&1
python scripts/summarize_ncu.py launches gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_launches_bench_c5.md > /dev/null
timeout 200 python scripts/config_table.py > gpurun_out/${TAG}_configs.md 2>&1
cat gpurun_out/${TAG}_configs.md
rm -f gpurun_out/${TAG}_wave_c3.ncu-rep   # keep the C5 report (source page) only: size
ls -la gpurun_out | tail -20
