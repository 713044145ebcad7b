#!/bin/bash
# A/B: compat LZ4 with chunks read in place (CW_LZ4_COMPAT_DIRECT_ABOVE: slots hold only the table, 13 parsers per SM) against
# staged chunks, stand-alone and fused.  usage: gpu_compat_direct_ab.sh "CHUNKS" "THRESHOLDS"
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/r02_ab_compat_direct.txt
: > $O
for chunk in ${1:-8192 4096}; do
  for above in ${2:-16384 2048}; do
    for what in lz4 fused; do
      echo "## CW_LZ4_COMPAT_DIRECT_ABOVE=$above chunk=$chunk $what" >> $O
      CW_LZ4_COMPAT_DIRECT_ABOVE=$above timeout 300 python tools/prof_classes.py $what $chunk 2>&1 | grep -v "synthetic [PTRX]\|alice\|ptt5\|world" | tail -4 >> $O
    done
  done
done
cat $O
