#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for sl in 3 4; do for mb in 96 128 192 256; do
  CW_SLOTS=$sl CW_SLICE_MB=$mb timeout 300 python tools/prof_e2e.py 2>&1 | head -1
done; done
CW_MODE=1 CW_SLOTS=4 CW_SLICE_MB=128 timeout 300 python tools/prof_e2e.py 2>&1 | head -1
