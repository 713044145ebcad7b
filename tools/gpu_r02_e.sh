#!/bin/bash
# round 2, session E: fast LZ4 as parser + shared emitters (libcomputewar_fq.so) vs self-emitting warps; emitter sweep
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
FQ=$GRAFT_REPO_ROOT/compute_war_b200/libcomputewar_fq.so
{
CW_LIB=$FQ timeout 900 python -m pytest tests -m gpu -x -q -k "fast or lz4 or fused or process_block" 2>&1 | tail -5
for em in 3 5 7; do
echo "== queue form, lz4 8192 fast, emitters $em"; CW_LZ4_EMITTERS=$em CW_LIB=$FQ CW_MODE=1 timeout 300 python tools/prof_classes.py lz4 8192 2>&1 | tail -12
done
for em in 3 5; do for hw in 2; do
echo "== queue form, fused 8192 fast, emitters $em hash warps $hw"; CW_FUSED_HASH_WARPS=$hw CW_LZ4_EMITTERS=$em CW_LIB=$FQ CW_MODE=1 timeout 300 python tools/prof_classes.py fused 8192 2>&1 | tail -12
done; done
} > gpurun_out/r02_e.log 2>&1
tail -80 gpurun_out/r02_e.log
