#!/bin/bash
# compute-sanitizer memcheck over the tests that mix pack / dedup / packed host output on one context
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python - > gpurun_out/skein_bits_probe.log 2>&1 <<'PY'
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, compute_war_b200 as cw, oracle_lib as O
eng=cw.Engine(0)
rng=np.random.default_rng(3)
bad=[]
for rep in range(3):
    for ln in list(range(1,130))+[200,256,1000,4096]:
        m=rng.integers(0,256,ln,dtype=np.uint8)
        for st,hb in ((512,512),(512,384),(256,256),(256,224)):
            g=eng.hash_skein(st,hb,m,ln).tobytes()
            if g!=O.skein(st,hb,m.tobytes()): bad.append((rep,ln,st,hb))
print("bad", len(bad), bad[:20])
eng.close()
PY
tail -3 gpurun_out/skein_bits_probe.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -x -q \
  -k "pack_grow_then_dedup or packed_host_output or packed_output or dedup_map or skein_kat_vectors" > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/memcheck.log
grep -m5 -A12 "Invalid\|ERROR SUMMARY\|=========" gpurun_out/memcheck.log | head -60
tail -8 gpurun_out/memcheck.log
