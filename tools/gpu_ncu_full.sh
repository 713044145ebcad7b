#!/bin/bash
# ncu --set full captures (text summaries only come back): usage gpu_ncu_full.sh tag "mode what class n chunk name" ...
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=$1; shift
for spec in "$@"; do
  set -- $spec
  CW_MODE=$1 timeout 120 python tools/prof_one.py $2 $3 $4 $5 > gpurun_out/plain_$6.log 2>&1 || { cat gpurun_out/plain_$6.log; continue; }
  CW_MODE=$1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:cw_ -s 1 -c 1 -f -o gpurun_out/prof_${T}_$6 python tools/prof_one.py $2 $3 $4 $5 > gpurun_out/ncu_$6.log 2>&1
  tail -1 gpurun_out/ncu_$6.log
  python tools/ncu_summary.py gpurun_out/prof_${T}_$6.ncu-rep > gpurun_out/${T}_ncu_$6.txt 2>&1
  ncu -i gpurun_out/prof_${T}_$6.ncu-rep --page raw --csv > gpurun_out/${T}_raw_$6.csv 2>/dev/null
  rm -f gpurun_out/prof_${T}_$6.ncu-rep
done
ls -la gpurun_out | tail -8
