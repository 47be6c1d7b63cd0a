#!/bin/bash
# Round-2 ncu evidence (one GPU): full captures of the wave kernel (C5, C3, HPP K=6, FHP-I K=5, an edge-tile strip launch) and
# of the SM-resident kernel (C1, Karman default, C2), the launch list of the bench command.  Summaries -> profiles/ (tracked).
TAG=${1:-r02}
mkdir -p gpurun_out
NCUW="ncu --set full --clock-control none --import-source on -k regex:step_wave --launch-skip 3 -c 1 -f"
NCUR="ncu --set full --clock-control none --import-source on -k regex:step_resident -c 1 -f"
cap() { # cap <name> <ncu prefix...> -- <command...>
  local name=$1; shift; local pre=(); while [ "$1" != "--" ]; do pre+=("$1"); shift; done; shift
  timeout 300 "${pre[@]}" -o gpurun_out/${TAG}_$name "$@" > gpurun_out/${TAG}_ncu_$name.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_$name.log
}
K=$(python -c "import lgca_b200; e=lgca_b200.Engine('FHP_III',32768,32768); print(e.info().k_fuse)")
cap wave_c5      $NCUW -- python scripts/prof_one.py FHP_III 32768 32768 periodic 0 1 30
cap wave_c3      $NCUW -- python scripts/prof_one.py FHP_III 16384 8192 karman 0 1 30
cap wave_hpp_k6  $NCUW -- python scripts/prof_one.py HPP 4096 4096 periodic 6 5 30
cap wave_fhp1_k5 $NCUW -- python scripts/prof_one.py FHP_I 1400 700 pipe 5 5 30
cap wave_strip   ncu --set full --clock-control none --import-source on -k regex:step_wave --launch-skip 6 -c 1 -f -- python scripts/prof_strip.py FHP_III 16384 8192 periodic 8
cap res_c1       $NCUR -- python scripts/prof_one.py FHP_I 1400 700 pipe 0 1 300
cap res_karman   $NCUR -- python scripts/prof_one.py FHP_III 4400 2200 karman 0 1 200
cap res_hpp      $NCUR -- python scripts/prof_one.py HPP 4096 4096 periodic 0 1 200
S=scripts/summarize_ncu.py
python $S full gpurun_out/${TAG}_wave_c5.ncu-rep profiles/${TAG}_wave_k${K}_c5_full.md periodic_k${K} > /dev/null
python $S full gpurun_out/${TAG}_wave_c3.ncu-rep profiles/${TAG}_wave_k${K}_c3_full.md karman_k${K} > /dev/null
python $S full gpurun_out/${TAG}_wave_hpp_k6.ncu-rep profiles/${TAG}_wave_k6_hpp4096_full.md > /dev/null
python $S full gpurun_out/${TAG}_wave_fhp1_k5.ncu-rep profiles/${TAG}_wave_k5_fhp1_c1_full.md > /dev/null
python $S full gpurun_out/${TAG}_wave_strip.ncu-rep profiles/${TAG}_wave_strip_edge_tile_full.md > /dev/null
python $S full gpurun_out/${TAG}_res_c1.ncu-rep profiles/${TAG}_resident_c1_full.md > /dev/null
python $S full gpurun_out/${TAG}_res_karman.ncu-rep profiles/${TAG}_resident_karman_default_full.md > /dev/null
python $S full gpurun_out/${TAG}_res_hpp.ncu-rep profiles/${TAG}_resident_hpp4096_full.md > /dev/null
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-karman --no-box --no-app-tick > gpurun_out/${TAG}_launches_bench.log 2>&1
python $S launches gpurun_out/${TAG}_launches.csv profiles/${TAG}_launches_bench_c5.md > /dev/null
mkdir -p gpurun_out/profiles_${TAG}; cp profiles/traffic.json profiles/${TAG}_*.md gpurun_out/profiles_${TAG}/
rm -f gpurun_out/${TAG}_wave_c3.ncu-rep gpurun_out/${TAG}_wave_hpp_k6.ncu-rep gpurun_out/${TAG}_wave_fhp1_k5.ncu-rep gpurun_out/${TAG}_res_karman.ncu-rep gpurun_out/${TAG}_res_hpp.ncu-rep
ls -la gpurun_out/profiles_${TAG}; head -30 profiles/${TAG}_launches_bench_c5.md
