TAG=r02w
NCUR="ncu --set full --clock-control none --import-source on -k regex:step_resident -c 1 -f"
cap() { local name=$1; shift; local pre=(); while [ "$1" != "--" ]; do pre+=("$1"); shift; done; shift
  timeout 300 "${pre[@]}" -o gpurun_out/${TAG}_$name "$@" > gpurun_out/${TAG}_ncu_$name.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_$name.log; }
cap res_c1       $NCUR -- python scripts/prof_one.py FHP_I 1400 700 pipe 0 1 300
cap res_karman   $NCUR -- python scripts/prof_one.py FHP_III 4400 2200 karman 0 1 200
cap res_hpp      $NCUR -- python scripts/prof_one.py HPP 4096 4096 periodic 0 1 200
for n in c1 karman hpp; do python scripts/summarize_ncu.py full gpurun_out/${TAG}_res_$n.ncu-rep gpurun_out/${TAG}_resident_${n}_full.md > /dev/null; done
rm -f gpurun_out/${TAG}_res_karman.ncu-rep gpurun_out/${TAG}_res_hpp.ncu-rep
