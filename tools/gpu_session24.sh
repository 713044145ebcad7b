#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/prof_one.py lz4 mixed 32768 > gpurun_out/s24_one.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cw_lz4_kernel -s 1 -c 1 -f -o gpurun_out/prof_r01_lz4v5_mixed python tools/prof_one.py lz4 mixed 32768 > gpurun_out/ncu_lz4m.log 2>&1; tail -2 gpurun_out/ncu_lz4m.log
