#!/bin/bash
# A/B of the team size of the kernels whose teams wait on each other (few slots per SM): CW_LZ4F_TEAM = warps per chunk
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests -m gpu -x -q -k "fast" 2>&1 | tail -2
for t in ${1:-0 6 7}; do echo "== lzf 8192 fast team $t (0: default shape)"; CW_LZ4F_TEAM=$t CW_MODE=1 timeout 90 python tools/prof_classes.py lzf 8192 2>&1 | grep -E "mixed|canterbury"; done
for t in ${1:-0 6 7}; do echo "== lz4 65536 fast team $t"; CW_LZ4F_TEAM=$t CW_MODE=1 timeout 120 python tools/prof_classes.py lz4 65536 2>&1 | grep -E "mixed|canterbury"; done
echo "== fusedlzf 8192 default"; CW_MODE=1 timeout 90 python tools/prof_classes.py fusedlzf 8192 2>&1 | grep -E "mixed|canterbury"
} > gpurun_out/r02_team_ab.log 2>&1
cat gpurun_out/r02_team_ab.log
