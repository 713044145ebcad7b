#!/bin/bash
# 8-GPU box: topology, then the bench with and without NUMA binding of the ranks
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( nvidia-smi topo -m; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ "$(cat $d/class)" = "0x030200" ]; then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done; nproc; python -c "import os; print(sorted(os.sched_getaffinity(0)))" ) > gpurun_out/s25_topo.txt 2>&1
cat gpurun_out/s25_topo.txt
for numa in 1 0; do
  CW_BENCH_NUMA=$numa timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/s25_bench8_numa$numa.json 2> gpurun_out/s25_bench8_numa$numa.err
  python - <<PY
import json
l=[x for x in open("gpurun_out/s25_bench8_numa$numa.json") if x.startswith("{")]
d=json.loads(l[-1]); print("numa=$numa value", d["value"], "e2e", d["e2e"]["value"], "fast", d["fast_mode"]["value"], d["config"].get("host_numa_binding_rank0"))
PY
done
