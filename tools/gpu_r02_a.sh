#!/bin/bash
# round 2, session A: first run of the window-parallel LZ4 fast parse (csrc/lz4_fast.cuh)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for m in 0 1; do
  echo "== lz4 8192 mode $m"; CW_MODE=$m timeout 300 python tools/prof_classes.py lz4 8192 2>&1 | tail -12
done
echo "== fused 8192 fast"; CW_MODE=1 timeout 300 python tools/prof_classes.py fused 8192 2>&1 | tail -12
echo "== lz4 4096 fast"; CW_MODE=1 timeout 300 python tools/prof_classes.py lz4 4096 2>&1 | tail -12
echo "== lz4 65536 fast"; CW_MODE=1 timeout 300 python tools/prof_classes.py lz4 65536 2>&1 | tail -12
echo "== lz4 65536 compat"; CW_MODE=0 timeout 300 python tools/prof_classes.py lz4 65536 2>&1 | tail -12
echo "== fused 65536 fast"; CW_MODE=1 timeout 300 python tools/prof_classes.py fused 65536 2>&1 | tail -12
echo "== fused 65536 compat"; CW_MODE=0 timeout 300 python tools/prof_classes.py fused 65536 2>&1 | tail -12
echo "== variant REVERSE=0: lz4 8192 fast"; CW_LIB=$GRAFT_REPO_ROOT/compute_war_b200/libcomputewar_norev.so CW_MODE=1 timeout 300 python tools/prof_classes.py lz4 8192 2>&1 | tail -12
} > gpurun_out/r02_a.log 2>&1
tail -100 gpurun_out/r02_a.log
