#!/bin/bash
# N-GPU copy-only probe + bench.py (both arms) on the same box: e2e against its measured copy ceiling (N = $1)
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-4}
mkdir -p gpurun_out
for n in 1 $N; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 tools/pcie_probe_multi.py 2>/dev/null | grep "^{" | tee -a gpurun_out/pcie_probe_multi.jsonl
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
grep "^{" gpurun_out/bench_${N}gpu.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$N gpus value', d['value'], 'e2e', d['e2e'], 'fast', d['fast_mode']['value'], d['clocks'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_${N}gpu.json 2>> gpurun_out/bench_${N}gpu.err
grep "^{" gpurun_out/bench_ref_${N}gpu.json | cut -c1-200
nvidia-smi topo -m 2>/dev/null | head -14
