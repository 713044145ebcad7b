#!/bin/bash
# A/B: how often the parsers publish their records (CW_LZ4_PUB_BATCH 1 / 8 / 16)
cd "$GRAFT_REPO_ROOT" || exit 1
for lib in libcomputewar_pub1.so libcomputewar.so libcomputewar_pub16.so; do
  echo "== $lib"
  CW_LIB=$GRAFT_REPO_ROOT/compute_war_b200/$lib timeout 300 python tools/prof_kernels.py lz4,lzf,fused 262144 8192 3 2>&1 | tail -3
  CW_LIB=$GRAFT_REPO_ROOT/compute_war_b200/$lib timeout 300 python tools/prof_classes.py fused 8192 2>&1 | grep "tiled\|text"
done
