#!/bin/bash
# A/B: second hash warp / third emitter next to the hash warp's scheduler (compat fused kernel)
cd "$GRAFT_REPO_ROOT" || exit 1
run() { echo "== $*"; env "$@" CW_MODE=0 timeout 300 python tools/prof_classes.py fused 8192 2>&1 | grep "mixed\|tiled\|random"; }
run CW_ROLE_MAP="3;3,3"
run CW_FUSED_HASH_WARPS=2 CW_ROLE_MAP="3,3;3,0"
run CW_FUSED_HASH_WARPS=2 CW_ROLE_MAP="3,0;3,3"
run CW_LZ4_EMITTERS=3 CW_ROLE_MAP="3;3,3,0"
