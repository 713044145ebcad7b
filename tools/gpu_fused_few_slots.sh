#!/bin/bash
# fused fast kernels with few slots per SM (LZF 8 KiB, 64 KiB chunks): tests + tables
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests -m gpu -x -q -k "fast or fused" 2>&1 | tail -2
echo "== fusedlzf 8192 fast"; CW_MODE=1 timeout 90 python tools/prof_classes.py fusedlzf 8192 2>&1 | grep -E "mixed|canterbury"
echo "== fused (lz4) 65536 fast"; CW_MODE=1 timeout 120 python tools/prof_classes.py fused 65536 2>&1 | grep -E "mixed|canterbury"
echo "== fusedlzf 65536 fast"; CW_MODE=1 timeout 120 python tools/prof_classes.py fusedlzf 65536 2>&1 | grep -E "mixed|canterbury"
echo "== fused (lz4) 8192 fast"; CW_MODE=1 timeout 90 python tools/prof_classes.py fused 8192 2>&1 | grep -E "mixed|canterbury"
} > gpurun_out/r02_fused_few_slots_${1:-a}.log 2>&1
cat gpurun_out/r02_fused_few_slots_${1:-a}.log
