#!/bin/bash
# 64 KiB chunks (in-place reads): prefetch of the next window's candidate lines into L1, on / off
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for lib in libcomputewar_nodpf.so libcomputewar.so; do
  echo "== $lib"
  export CW_LIB=$GRAFT_REPO_ROOT/compute_war_b200/$lib
  timeout 300 python tools/prof_classes.py lz4 65536 2>&1 | grep "mixed\|tiled\|text\|pattern\|records"
  timeout 300 python tools/prof_classes.py lz4 32768 2>&1 | grep "mixed\|tiled"
  timeout 300 python tools/prof_kernels.py fused 16384 65536 3 2>&1 | tail -1
done
