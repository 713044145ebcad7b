"""Copy-only probe for N processes (one per GPU, launched like bench.py under torch.distributed.run): pinned host memory,
64 MiB pieces, H2D alone / D2H alone / both at once / both with the D2H leg at 70 % (what a packed result carries).
Every rank copies its own 1 GiB; ranks start together (barrier) and the slowest rank's time counts, so the figure is
the AGGREGATE host<->device rate of the box, the ceiling of bench.py's e2e leg.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe_multi.py"""
import json
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
N = 1 << 30
piece = 64 << 20
h_in = torch.empty(N, dtype=torch.uint8).pin_memory()
h_out = torch.empty(N, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
h_out.fill_(2)
d_a = torch.empty(N, dtype=torch.uint8, device=dev)
d_b = torch.empty(N, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(device_ids=[local])


def run(h2d, d2h, frac=1.0, reps=3):
    best = None
    for _ in range(reps + 1):
        barrier()
        t0 = time.perf_counter()
        for off in range(0, N, piece):
            if h2d:
                with torch.cuda.stream(s1):
                    d_a[off:off + piece].copy_(h_in[off:off + piece], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    k = int(piece * frac)
                    h_out[off:off + k].copy_(d_b[off:off + k], non_blocking=True)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = float(t.item()) if best is None else min(best, float(t.item()))
    return best


out = {"n_gpus": world, "bytes_per_gpu": N, "piece": piece}
for name, a in (("h2d_only", (True, False)), ("d2h_only", (False, True)), ("both", (True, True)), ("both_d2h_70pct", (True, True, 0.7))):
    t = run(*a)
    out[name + "_GBps_aggregate_per_direction"] = world * N / t / 1e9
if rank == 0:
    try:
        out["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
        out["host_cpus"] = len(os.sched_getaffinity(0))
    except Exception:
        pass
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
