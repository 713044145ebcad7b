#!/bin/bash
# round 2, lean fast-mode LZ4 (E / N entries + one-load walk): GPU suite, per-class tables for team sizes, source-level counters
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=${1:-a}
{
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for nt in 2 3; do
echo "== lz4 8192 fast, team $nt"; CW_LZ4F_TEAM=$nt CW_MODE=1 timeout 90 python tools/prof_classes.py lz4 8192 2>&1 | tail -12
done
for nt in 2 3; do
echo "== fused 8192 fast, team $nt"; CW_LZ4F_TEAM=$nt CW_MODE=1 timeout 90 python tools/prof_classes.py fused 8192 2>&1 | tail -12
done
echo "== lz4 65536 fast, team 4"; CW_MODE=1 timeout 120 python tools/prof_classes.py lz4 65536 2>&1 | tail -12
echo "== lz4 4096 fast, team 2"; CW_MODE=1 timeout 90 python tools/prof_classes.py lz4 4096 2>&1 | tail -12
} > gpurun_out/r02_fast_v2_$T.log 2>&1
tail -100 gpurun_out/r02_fast_v2_$T.log
bash tools/gpu_ncu_source.sh 1 lz4 file:bible.txt 16384 8192 lz4fast_v2_bible_$T
