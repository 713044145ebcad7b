#!/bin/bash
# quick post-change check: GPU parity suite + stand-alone kernel timings
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/prof_kernels.py ${1:-all} 262144 ${2:-8192} 3 2>&1 | tail -8
