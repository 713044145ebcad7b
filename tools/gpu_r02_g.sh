#!/bin/bash
# round 2, session G: fast LZ4 teams with mbarrier tokens; team-size sweep
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for nt in 1 2 3; do
echo "== lz4 8192 fast, team $nt"; CW_LZ4F_TEAM=$nt CW_MODE=1 timeout 300 python tools/prof_classes.py lz4 8192 2>&1 | tail -12
done
for nt in 1 2; do
echo "== fused 8192 fast, team $nt"; CW_LZ4F_TEAM=$nt CW_MODE=1 timeout 300 python tools/prof_classes.py fused 8192 2>&1 | tail -12
done
echo "== lz4 65536 fast, team 4"; CW_LZ4F_TEAM=4 CW_MODE=1 timeout 300 python tools/prof_classes.py lz4 65536 2>&1 | tail -12
} > gpurun_out/r02_g.log 2>&1
tail -80 gpurun_out/r02_g.log
CW_MODE=1 python tools/prof_one.py lz4 file:bible.txt 16384 > gpurun_out/plain_g.log 2>&1 &&
CW_MODE=1 ncu --set full --clock-control none --import-source on -k regex:cw_lz4f -s 1 -c 1 -f -o gpurun_out/prof_r02g_lz4fast_bible python tools/prof_one.py lz4 file:bible.txt 16384 > gpurun_out/ncu_g.log 2>&1
tail -3 gpurun_out/ncu_g.log
