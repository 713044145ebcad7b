#!/bin/bash
# N-GPU copy-only probe + bench.py without the per-config table / CPU leg (N GPU-minutes per minute): e2e against the copy ceiling
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/pcie_probe_multi.py 2>/dev/null | grep "^{" | tee -a gpurun_out/pcie_probe_multi_${N}.jsonl
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 5 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
grep "^{" gpurun_out/bench_${N}gpu.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$N gpus value', d['value'], 'e2e', d['e2e'], 'fast', d['fast_mode']['value'], d['clocks'])"
tail -3 gpurun_out/bench_${N}gpu.err
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/topo_${N}gpu.txt
