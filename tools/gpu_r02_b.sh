#!/bin/bash
# round 2, session B: state of the restored tree (4-window groups) + one ncu capture of the LZ4 fast kernel on text
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== lz4 8192 fast"; CW_MODE=1 timeout 300 python tools/prof_classes.py lz4 8192 2>&1 | tail -12
echo "== fused 8192 fast"; CW_MODE=1 timeout 300 python tools/prof_classes.py fused 8192 2>&1 | tail -12
} > gpurun_out/r02_b.log 2>&1
tail -40 gpurun_out/r02_b.log
CW_MODE=1 python tools/prof_one.py lz4 file:bible.txt 16384 > gpurun_out/plain_b.log 2>&1 &&
CW_MODE=1 ncu --set full --clock-control none --import-source on -k regex:cw_lz4 -s 1 -c 1 -f -o gpurun_out/prof_r02_lz4fast_bible python tools/prof_one.py lz4 file:bible.txt 16384 > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/ncu_b.log
