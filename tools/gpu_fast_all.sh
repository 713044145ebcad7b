#!/bin/bash
# fast-mode kernels after a change: fast / fused GPU tests, then per-class tables: LZ4 and LZF, stand-alone and fused, 8 KiB (+ 64 KiB / 4 KiB with "more")
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=${1:-q}
{
timeout 600 python -m pytest tests -m gpu -x -q -k "fast or fused or process_block or ragged" 2>&1 | tail -3
echo "== lz4 8192 fast"; CW_MODE=1 timeout 90 python tools/prof_classes.py lz4 8192 2>&1 | tail -12
echo "== fused 8192 fast"; CW_MODE=1 timeout 90 python tools/prof_classes.py fused 8192 2>&1 | tail -12
echo "== lzf 8192 fast"; CW_MODE=1 timeout 90 python tools/prof_classes.py lzf 8192 2>&1 | tail -12
echo "== fused skein512+lzf 8192 fast"; CW_MODE=1 timeout 90 python tools/prof_classes.py fusedlzf 8192 2>&1 | tail -12
if [ -n "$2" ]; then
echo "== lz4 65536 fast"; CW_MODE=1 timeout 120 python tools/prof_classes.py lz4 65536 2>&1 | tail -12
echo "== lzf 65536 fast"; CW_MODE=1 timeout 120 python tools/prof_classes.py lzf 65536 2>&1 | tail -12
echo "== lz4 4096 fast"; CW_MODE=1 timeout 90 python tools/prof_classes.py lz4 4096 2>&1 | tail -12
echo "== lzf 4096 fast"; CW_MODE=1 timeout 90 python tools/prof_classes.py lzf 4096 2>&1 | tail -12
fi
} > gpurun_out/r02_fast_all_$T.log 2>&1
cat gpurun_out/r02_fast_all_$T.log
