#!/bin/bash
# LZF fast mode (the LZF flavour of lz4_fast.cuh): GPU suite, then per-class tables compat vs fast, stand-alone and fused
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=${1:-a}
{
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== lzf 8192 fast"; CW_MODE=1 timeout 120 python tools/prof_classes.py lzf 8192 2>&1 | tail -12
echo "== lzf 8192 compat"; CW_MODE=0 timeout 120 python tools/prof_classes.py lzf 8192 2>&1 | tail -12
echo "== fused skein512+lzf 8192 fast"; CW_MODE=1 timeout 120 python tools/prof_classes.py fusedlzf 8192 2>&1 | tail -12
echo "== lzf 4096 fast"; CW_MODE=1 timeout 120 python tools/prof_classes.py lzf 4096 2>&1 | tail -12
echo "== lzf 65536 fast"; CW_MODE=1 timeout 180 python tools/prof_classes.py lzf 65536 2>&1 | tail -12
} > gpurun_out/r02_lzf_fast_$T.log 2>&1
cat gpurun_out/r02_lzf_fast_$T.log
