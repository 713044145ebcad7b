#!/bin/bash
# quick A/B after a change to the fast-mode kernels: fast-mode GPU tests, per-class tables (stand-alone and fused, 8 KiB)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=${1:-q}
{
timeout 600 python -m pytest tests -m gpu -x -q -k "fast or fused or process_block" 2>&1 | tail -3
echo "== lz4 8192 fast"; CW_MODE=1 timeout 90 python tools/prof_classes.py lz4 8192 2>&1 | tail -12
echo "== fused 8192 fast"; CW_MODE=1 timeout 90 python tools/prof_classes.py fused 8192 2>&1 | tail -12
} > gpurun_out/r02_fast_quick_$T.log 2>&1
cat gpurun_out/r02_fast_quick_$T.log
if [ -n "$2" ]; then
{
echo "== lz4 65536 fast"; CW_MODE=1 timeout 120 python tools/prof_classes.py lz4 65536 2>&1 | tail -12
echo "== lz4 4096 fast"; CW_MODE=1 timeout 90 python tools/prof_classes.py lz4 4096 2>&1 | tail -12
} > gpurun_out/r02_fast_quick_$T.more.log 2>&1
cat gpurun_out/r02_fast_quick_$T.more.log
fi
