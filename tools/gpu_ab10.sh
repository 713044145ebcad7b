#!/bin/bash
# A/B: emitters grouped on one scheduler in the stand-alone LZ4 kernel (CW_LZ4_ROLE_MAP=0: roles in warp order)
cd "$GRAFT_REPO_ROOT" || exit 1
for v in 0 1; do
  echo "== CW_LZ4_ROLE_MAP=$v"
  CW_LZ4_ROLE_MAP=$v timeout 300 python tools/prof_kernels.py lz4 262144 8192 3 2>&1 | tail -1
  CW_LZ4_ROLE_MAP=$v timeout 300 python tools/prof_kernels.py lz4 262144 4096 3 2>&1 | tail -1
  CW_LZ4_ROLE_MAP=$v CW_MODE=1 timeout 300 python tools/prof_classes.py lz4 8192 2>&1 | grep "mixed\|tiled" | sed 's/^/fast /'
  CW_LZ4_ROLE_MAP=$v timeout 300 python tools/prof_kernels.py lz4 16384 65536 3 2>&1 | tail -1
done
