#!/bin/bash
# source-level counters (per-SASS-instruction executed counts) of one kernel; the report comes back for tools/ncu_groups.py
# usage: gpu_ncu_source.sh <mode> <lz4|lzf|fused> <class> <n_chunks> <chunk> <tag>
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CW_MODE=$1 timeout 120 python tools/prof_one.py $2 $3 $4 $5 > gpurun_out/plain_$6.log 2>&1 || { cat gpurun_out/plain_$6.log; exit 1; }
CW_MODE=$1 timeout 600 ncu --section SourceCounters --section InstructionStats --section WarpStateStats --section SchedulerStats --clock-control none --import-source on -k regex:cw_ -s 1 -c 1 -f -o gpurun_out/src_$6 python tools/prof_one.py $2 $3 $4 $5 > gpurun_out/ncu_$6.log 2>&1
tail -2 gpurun_out/ncu_$6.log; ls -la gpurun_out/src_$6.ncu-rep
