#!/bin/bash
# A/B: LZF table for 8 KiB chunks: 8192 buckets (32 KB, 5 parsers per SM) vs 4096 buckets + verify (24 KB, 6 parsers)
# variant: bash tools/build_variant.sh lzf13 -DCW_LZF_FORMAT_8K=13
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for lib in libcomputewar_lzf13.so libcomputewar.so; do
  echo "== $lib"
  export CW_LIB=$GRAFT_REPO_ROOT/compute_war_b200/$lib
  timeout 300 python tools/prof_kernels.py lzf,fusedlzf 262144 8192 3 2>&1 | tail -2
  timeout 300 python tools/prof_classes.py lzf 8192 2>&1 | grep -v "^$"
done
