#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -12
timeout 600 bash tools/small_batch_sweep.sh
