#!/bin/bash
# A/B of the fast LZ4 kernels: windows per group (SW), warps per chunk; every run under its own timeout
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=$GRAFT_REPO_ROOT/compute_war_b200
{
echo "== lz4 8192 fast, SW=4 team 2 (default)"; CW_MODE=1 timeout 120 python tools/prof_classes.py lz4 8192 2>&1 | grep -E "mixed|canterbury|text|random"
echo "== lz4 8192 fast, SW=2 team 2"; CW_LIB=$L/libcomputewar_sw2.so CW_MODE=1 timeout 120 python tools/prof_classes.py lz4 8192 2>&1 | grep -E "mixed|canterbury|text|random"
echo "== lz4 8192 fast, SW=2 team 3 (39 warps, 48 registers)"; CW_LZ4F_TEAM=3 CW_LIB=$L/libcomputewar_sw2w39.so CW_MODE=1 timeout 120 python tools/prof_classes.py lz4 8192 2>&1 | grep -E "mixed|canterbury|text|random"
echo "== lz4 8192 fast, SW=2 team 2 (39-warp build, 48 registers)"; CW_LZ4F_TEAM=2 CW_LIB=$L/libcomputewar_sw2w39.so CW_MODE=1 timeout 120 python tools/prof_classes.py lz4 8192 2>&1 | grep -E "mixed|canterbury|text|random"
echo "== fused 8192 fast, SW=2 team 2"; CW_LIB=$L/libcomputewar_sw2.so CW_MODE=1 timeout 120 python tools/prof_classes.py fused 8192 2>&1 | grep -E "mixed|canterbury|text|random"
} > gpurun_out/r02_ab_sw.log 2>&1
cat gpurun_out/r02_ab_sw.log
