#!/bin/bash
# round 2, session L: A/B of the fast fused kernel: doubling rounds, hash warps; bench line with copy ceiling
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
for v in d1 "" d3; do
  lib=$GRAFT_REPO_ROOT/compute_war_b200/libcomputewar${v:+_$v}.so
  echo "== fused 8192 fast, doublings variant '${v:-2 (default)}'"; CW_LIB=$lib CW_MODE=1 timeout 120 python tools/prof_classes.py fused 8192 2>&1 | grep -E "mixed|canterbury|text|random"
done
for hw in 1 2; do
  echo "== fused 8192 fast, hash warps $hw"; CW_FUSED_HASH_WARPS=$hw CW_MODE=1 timeout 120 python tools/prof_classes.py fused 8192 2>&1 | grep -E "mixed|canterbury"
done
echo "== fused 65536 fast (team 4)"; CW_MODE=1 timeout 200 python tools/prof_classes.py fused 65536 2>&1 | grep -E "mixed|canterbury"
echo "== lz4 65536 fast (team 4)"; CW_MODE=1 timeout 200 python tools/prof_classes.py lz4 65536 2>&1 | grep -E "mixed|canterbury"
} > gpurun_out/r02_l.log 2>&1
cat gpurun_out/r02_l.log
timeout 600 python bench.py --no-configs > gpurun_out/r02_l_bench.json 2> gpurun_out/r02_l_bench.err; tail -2 gpurun_out/r02_l_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02_l_bench.json').read().strip().splitlines()[-1])
print('value', l['value'], 'fast', l['fast_mode']['value']); print('e2e', json.dumps(l['e2e'])[:900])
PY
