#!/bin/bash
# 2 GPUs: full GPU suite (host tool with --devices 2 included), then the 2-GPU bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s26_bench2.json 2> gpurun_out/s26_bench2.err
grep "^{" gpurun_out/s26_bench2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('2gpu value', d['value'], 'e2e', d['e2e']['value'], 'fast', d['fast_mode']['value'])"
