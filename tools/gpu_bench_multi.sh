#!/bin/bash
# bench.py on N GPUs as the driver launches it (N = $1)
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
grep "^{" gpurun_out/bench_${N}gpu.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$N gpus value', d['value'], 'e2e', d['e2e']['value'], 'fast', d['fast_mode']['value'], d['clocks'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_${N}gpu.json 2>> gpurun_out/bench_${N}gpu.err
grep "^{" gpurun_out/bench_ref_${N}gpu.json | cut -c1-200
