#!/bin/bash
# ballot-based twin grouping (LZ4 + LZF), LZF split: parity, per-class timings, bench
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s20_tests.txt
cat gpurun_out/s20_tests.txt
for w in lzf lz4 fused; do
  echo "== $w" >> gpurun_out/s20_classes.txt
  timeout 300 python tools/prof_classes.py $w 8192 >> gpurun_out/s20_classes.txt 2>&1
done
echo "== fused fast" >> gpurun_out/s20_classes.txt
CW_MODE=1 timeout 300 python tools/prof_classes.py fused 8192 >> gpurun_out/s20_classes.txt 2>&1
cat gpurun_out/s20_classes.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s20_bench.json 2> gpurun_out/s20_bench.err
cat gpurun_out/s20_bench.json
