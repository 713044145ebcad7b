#!/bin/bash
# LZF parser/emitter split + tagged chains: parity, then per-class timings with 1/2/3 emitters
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s19_tests.txt
cat gpurun_out/s19_tests.txt
for e in 1 2 3; do
  echo "== CW_LZF_EMITTERS=$e" >> gpurun_out/s19_lzf.txt
  CW_LZF_EMITTERS=$e timeout 300 python tools/prof_classes.py lzf 8192 >> gpurun_out/s19_lzf.txt 2>&1
done
echo "== 4096" >> gpurun_out/s19_lzf.txt
timeout 300 python tools/prof_classes.py lzf 4096 >> gpurun_out/s19_lzf.txt 2>&1
cat gpurun_out/s19_lzf.txt
