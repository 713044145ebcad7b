#!/bin/bash
# hybrid twin grouping timings + ncu source-level capture of the LZF kernel on text
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/s21_tests.txt
cat gpurun_out/s21_tests.txt
for w in lzf lz4; do
  echo "== $w" >> gpurun_out/s21_classes.txt
  timeout 300 python tools/prof_classes.py $w 8192 >> gpurun_out/s21_classes.txt 2>&1
done
cat gpurun_out/s21_classes.txt
timeout 300 python tools/prof_one.py lzf file:bible.txt 16384 > gpurun_out/s21_one.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cw_lzf_kernel -s 1 -c 1 -f -o gpurun_out/prof_r01_lzf_bible python tools/prof_one.py lzf file:bible.txt 16384 > gpurun_out/ncu_lzf.log 2>&1; tail -2 gpurun_out/ncu_lzf.log
