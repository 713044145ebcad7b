#!/bin/bash
# A/B: compat LZ4 with chunks of 4-8 KiB read in place (CW_LZ4_COMPAT_DIRECT_ABOVE: slots hold only the table, 13 parsers per SM
# instead of 9) against the staged default; GPU suite + smoke first
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/r02_ab_compat_direct.txt
: > $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $O
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2 >> $O
for above in 16384 2048; do
  for chunk in 8192 4096; do
    echo "## CW_LZ4_COMPAT_DIRECT_ABOVE=$above chunk=$chunk" >> $O
    CW_LZ4_COMPAT_DIRECT_ABOVE=$above timeout 300 python tools/prof_classes.py lz4 $chunk 2>&1 | tail -12 >> $O
  done
done
echo "## byte identity of the in-place variant (compat LZ4 tests with the knob set)" >> $O
CW_LZ4_COMPAT_DIRECT_ABOVE=2048 timeout 600 python -m pytest tests -m gpu -x -q -k "lz4 or compress or corpus" 2>&1 | tail -3 >> $O
cat $O
