#!/bin/bash
# round 2, session H: fast LZ4 teams, per-warp token chains, 4-byte continuation counts; every run under its own timeout
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for nt in 2 3; do
echo "== lz4 8192 fast, team $nt"; CW_LZ4F_TEAM=$nt CW_MODE=1 timeout 90 python tools/prof_classes.py lz4 8192 2>&1 | tail -12
done
echo "== fused 8192 fast, team 2"; CW_MODE=1 timeout 90 python tools/prof_classes.py fused 8192 2>&1 | tail -12
echo "== fused 8192 fast, team 2, 3 hash warps... (n/a)"
for nt in 2 4; do
echo "== lz4 65536 fast, team $nt"; CW_LZ4F_TEAM=$nt CW_MODE=1 timeout 120 python tools/prof_classes.py lz4 65536 2>&1 | tail -12
done
echo "== lz4 4096 fast, team 2"; CW_MODE=1 timeout 90 python tools/prof_classes.py lz4 4096 2>&1 | tail -12
} > gpurun_out/r02_h.log 2>&1
tail -90 gpurun_out/r02_h.log
CW_MODE=1 timeout 120 python tools/prof_one.py fused file:bible.txt 16384 > gpurun_out/plain_h.log 2>&1 &&
CW_MODE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:cw_fused -s 1 -c 1 -f -o gpurun_out/prof_r02h_fusedfast_bible python tools/prof_one.py fused file:bible.txt 16384 > gpurun_out/ncu_h.log 2>&1
tail -3 gpurun_out/ncu_h.log
