#!/bin/bash
# A/B: hash warps of the fused fast kernels (variant build libcomputewar_hw3.so allows 3)
cd "$GRAFT_REPO_ROOT" || exit 1
L=$GRAFT_REPO_ROOT/compute_war_b200
mkdir -p gpurun_out
{
for cfg in "default 2" "hw3 2" "hw3 3"; do set -- $cfg
lib=""; [ "$1" = "hw3" ] && lib=$L/libcomputewar_hw3.so
echo "=== build $1, $2 hash warps"
for spec in "fused 8192" "fusedlzf 8192"; do set -- $spec $2; echo "== $1 $2 fast"; CW_FUSED_HASH_WARPS=${cfg##* } CW_LIB=$lib CW_MODE=1 timeout 120 python tools/prof_classes.py $1 $2 2>&1 | grep -E "mixed|canterbury|random"; done
done
} > gpurun_out/r02_ab_hash_warps.log 2>&1
cat gpurun_out/r02_ab_hash_warps.log
