#!/bin/bash
# A/B: how a warp waits for its team's tokens (suspending try_wait vs test_wait + fixed nanosleep back-off; variant builds)
cd "$GRAFT_REPO_ROOT" || exit 1
L=$GRAFT_REPO_ROOT/compute_war_b200
mkdir -p gpurun_out
{
for v in default ws64 ws200; do
lib=""; [ "$v" != "default" ] && lib=$L/libcomputewar_$v.so
echo "=== $v"
for spec in "lz4 8192" "lzf 8192" "lz4 65536" "fused 8192"; do set -- $spec; echo "== $1 $2 fast"; CW_LIB=$lib CW_MODE=1 timeout 120 python tools/prof_classes.py $1 $2 2>&1 | grep -E "mixed|canterbury"; done
done
} > gpurun_out/r02_ab_wait.log 2>&1
cat gpurun_out/r02_ab_wait.log
