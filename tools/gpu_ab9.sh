#!/bin/bash
# A/B: batched replay of twin-free unit windows in the LZ4 parser.  Build the variant first:
#   bash tools/build_variant.sh chain -DCW_LZ4_CHAIN=1
cd "$GRAFT_REPO_ROOT" || exit 1
for lib in libcomputewar.so libcomputewar_chain.so; do
  echo "== $lib"
  export CW_LIB=$GRAFT_REPO_ROOT/compute_war_b200/$lib
  timeout 300 python tools/prof_kernels.py lz4,fused 262144 8192 3 2>&1 | tail -2
  timeout 300 python tools/prof_classes.py lz4 8192 2>&1 | grep -v "^$"
done
