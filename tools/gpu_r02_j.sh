#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python - <<'PY'
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, compute_war_b200 as cw, oracle_lib as O
eng=cw.Engine(0)
rng=np.random.default_rng(3)
bad=[]
for ln in list(range(1,130))+[200,256,1000,4096]:
    m=rng.integers(0,256,ln,dtype=np.uint8)
    for st,hb in ((512,512),(512,384),(256,256),(256,224),(512,8),(256,128)):
        g=eng.hash_skein(st,hb,m,ln).tobytes()
        if g!=O.skein(st,hb,m.tobytes()): bad.append((ln,st,hb))
print("hash_skein bad", len(bad), bad[:20])
eng.close()
PY
timeout 900 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -14
