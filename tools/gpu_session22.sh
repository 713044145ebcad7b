#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/prof_e2e.py
CW_MODE=1 timeout 300 python tools/prof_e2e.py
CW_SLICE_MB=32 timeout 300 python tools/prof_e2e.py
