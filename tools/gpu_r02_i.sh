#!/bin/bash
# round 2, session I: full GPU suite (new KAT / packed-output / fast-mode checks) + the default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -20 > gpurun_out/r02_i_tests.log
cat gpurun_out/r02_i_tests.log
( time timeout 900 python bench.py > gpurun_out/r02_i_bench.json 2> gpurun_out/r02_i_bench.err ) 2>&1 | tail -4
tail -5 gpurun_out/r02_i_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02_i_bench.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches')})
print('e2e', l['e2e'])
print('fast', l['fast_mode'])
print('roofline', l['roofline'])
print('cpu', l.get('cpu_baseline'))
print('whole', l['config'].get('whole_sample_vs_reference_run'))
for c in (l.get('configs') or []):
    if isinstance(c, dict): print({k:(round(v,4) if isinstance(v,float) else v) for k,v in c.items() if k in ('config','hash','comp','mode','GBps','compression_ratio','parity','parity_first_chunks_vs_oracle','size_vs_reference_on_sample','device_roundtrip_all_blocks')}, (c.get('roofline') or {}).get('frac'))
    else: print(c, l['configs'][c]) 
PY
