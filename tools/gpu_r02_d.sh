#!/bin/bash
# round 2, session D: fast LZ4 kernel variants: tests, per-class table, one ncu capture on text
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== lz4 8192 fast"; CW_MODE=1 timeout 300 python tools/prof_classes.py lz4 8192 2>&1 | tail -12
echo "== fused 8192 fast"; CW_MODE=1 timeout 300 python tools/prof_classes.py fused 8192 2>&1 | tail -12
} > gpurun_out/r02_d.log 2>&1
tail -30 gpurun_out/r02_d.log
CW_MODE=1 python tools/prof_one.py lz4 file:bible.txt 16384 > gpurun_out/plain_d.log 2>&1 &&
CW_MODE=1 ncu --set full --clock-control none --import-source on -k regex:cw_lz4f -s 1 -c 1 -f -o gpurun_out/prof_r02d_lz4fast_bible python tools/prof_one.py lz4 file:bible.txt 16384 > gpurun_out/ncu_d.log 2>&1
tail -3 gpurun_out/ncu_d.log
