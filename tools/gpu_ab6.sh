#!/bin/bash
# A/B: next-block prefetch in the Skein loop (fused kernel's hash warp; stand-alone hash kernel)
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for lib in libcomputewar_nopf.so libcomputewar.so libcomputewar_allpf.so; do
  echo "== $lib"
  export CW_LIB=$GRAFT_REPO_ROOT/compute_war_b200/$lib
  timeout 300 python tools/prof_kernels.py skein512,skein256,fused,fusedsha 262144 8192 3 2>&1 | tail -4
  timeout 300 python tools/prof_classes.py fused 8192 2>&1 | grep "mixed\|tiled\|random"
  CW_MODE=1 timeout 300 python tools/prof_classes.py fused 8192 2>&1 | grep "mixed\|random" | sed 's/^/fast /'
done
