#!/bin/bash
# A/B: placement of the fused kernel's roles on the four schedulers.  CW_ROLE_MAP="h..;e.." = scheduler of every hash
# warp ; of every emitter ("0" = roles in warp order)
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest tests -m gpu -x -q -k "fused or process_block or large_batch" 2>&1 | tail -2
echo "## fast mode (13 parsers + 2 emitters + 2 hash warps)"
for rm in "0" "1,2;0,3" "3,2;3,2" "1,2;1,2" "2,3;0,1" "0,1;2,3"; do
  echo "== CW_ROLE_MAP=$rm"; CW_ROLE_MAP="$rm" CW_MODE=1 timeout 300 python tools/prof_classes.py fused 8192 2>&1 | grep "mixed\|tiled"
done
echo "## compat (9 + 2 + 1)"
for rm in "0" "3;3,3" "3;3,2" "3;2,1"; do
  echo "== CW_ROLE_MAP=$rm"; CW_ROLE_MAP="$rm" CW_MODE=0 timeout 300 python tools/prof_classes.py fused 8192 2>&1 | grep "mixed\|tiled"
done
