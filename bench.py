#!/usr/bin/env python
"""bench.py -- headline benchmark: fused Skein-512 + LZ4 over 8 KiB chunks, input GB/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the fused hot path (ProcessBlock, HashAndCompress.cpp:231-261, applied to a whole
shard) over one batch of synthetic input:  every rank owns 262144 chunks x 8 KiB = 2 GiB of the seeded
mixed-entropy stream of include/cw_synth.h, rank r holding chunks [r*262144, (r+1)*262144) -- at N = 8 that is
exactly BASELINE.json configs[4] (16 GiB sharded by chunk range).  Weak scaling, no data-path collective.

value   device-resident: the shard is generated in HBM, CUDA events around K steps, max over ranks.
e2e     the same pass through the host-buffer C-ABI call cw_hash_compress (pinned host input, strided
        host output, H2D + D2H inside the timed region, three-slot pipeline inside the library).
roofline / cpu_baseline: see DESIGN.md "Measurement".
--impl reference: the reference's own CPU primitives (oracle/_ref/ref_cli = its Skein source + vendored
        liblz4.a 1.8.2, compress-then-hash per chunk on all host cores) on a bounded sample of the same stream.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CHUNK = 8192
CHUNKS_PER_GPU = 262144  # 2 GiB per GPU
SEED = 0xC0FFEE
METRIC = "input GB/s, fused Skein-512+LZ4 @8KB chunks"
DIGEST = 64


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline
# ------------------------------------------------------------------------------------------------

def cpu_sample_chunks(threads: int) -> int:
    # bounded sample: 64 MiB per thread, at least 256 MiB, at most 4 GiB
    per_thread = 8192
    n = threads * per_thread
    return int(min(max(n, 32768), 524288))


def run_cpu_reference(n_chunks: int, threads: int, reps: int):
    """-> (list of seconds per rep, kind, ratio).  Uses oracle/_ref/ref_cli when present, else the C oracle port."""
    import oracle_lib as O
    if O.have_ref_cli():
        out = O.ref_cli("bench", "skein512", "lz4", str(CHUNK), str(n_chunks), str(threads), "0", hex(SEED), str(reps))
        recs = [json.loads(l) for l in out.strip().splitlines() if l.startswith("{")]
        return [r["seconds"] for r in recs], "reference", recs[-1]["ratio"]
    data = O.synth(CHUNK, 0, n_chunks, SEED)
    secs = []
    ratio = None
    for _ in range(reps):
        sec, _, _, lens = O.hash_compress_mt(O.HASH_SKEIN512, O.COMP_LZ4, data, CHUNK, threads)
        secs.append(sec)
        ratio = float(lens.sum()) / float(n_chunks * CHUNK)
    return secs, "port", ratio


def reference_arm(args, rank: int, world: int):
    if rank != 0:
        return
    threads = host_threads()
    n = cpu_sample_chunks(threads)
    secs, kind, ratio = run_cpu_reference(n, threads, args.warmup + args.steps)
    timed = secs[args.warmup:]
    total = sum(timed)
    value = n * CHUNK * len(timed) / total / 1e9
    sample = f"{n} chunks x {CHUNK} B = {n * CHUNK / 2**20:.0f} MiB prefix of the synthetic stream per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus, "steps": len(timed),
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(timed), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64/u8", "data": "synthetic",
        "config": {"workload": "fused skein512+lz4, 8 KiB chunks, synthetic mixed-entropy stream (CPU: bounded sample)",
                   "chunk_bytes": CHUNK, "sample_chunks": n, "compression_ratio": ratio},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(props):
    """Multi-GPU runs: keep this rank's threads (and so, by first touch, its pinned host buffers) on the NUMA node its
    GPU hangs off, so that the host<->device copies of the end-to-end leg do not cross the socket interconnect.
    CW_BENCH_NUMA=0 turns it off.  Returns what was done, for the JSON line."""
    if os.environ.get("CW_BENCH_NUMA", "1") == "0":
        return {"bound": False, "why": "CW_BENCH_NUMA=0"}
    try:
        dev = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        base = "/sys/bus/pci/devices/" + dev
        node = int(open(base + "/numa_node").read())
        cpus = _parse_cpulist(open(base + "/local_cpulist").read()) & os.sched_getaffinity(0)
        if node < 0 or not cpus:
            return {"bound": False, "why": "no NUMA information for " + dev}
        os.sched_setaffinity(0, cpus)
        return {"bound": True, "node": node, "cpus": len(cpus)}
    except Exception as e:  # sysfs not readable in this container: run unbound
        return {"bound": False, "why": type(e).__name__}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunks-per-gpu", type=int, default=CHUNKS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import compute_war_b200 as cw
    import oracle_lib as O

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(torch.cuda.get_device_properties(local)) if world > 1 else None
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner out of stdout: one JSON line only
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    eng = cw.Engine(local)
    n = args.chunks_per_gpu
    first = rank * n
    stride = cw.compress_bound(cw.COMP_LZ4, CHUNK)
    dev = torch.device("cuda", local)
    src = torch.empty(n * CHUNK, dtype=torch.uint8, device=dev)
    dig = torch.empty(n * DIGEST, dtype=torch.uint8, device=dev)
    blocks = torch.empty(n * stride, dtype=torch.uint8, device=dev)
    lens = torch.empty(n, dtype=torch.int32, device=dev)
    tstream = torch.cuda.Stream(device=dev)  # a real (non-default) stream: kernels and timing events share it
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    eng.synth_dev(src.data_ptr(), CHUNK, first, n, SEED, stream)
    torch.cuda.synchronize()

    def step():
        eng.hash_compress_dev(cw.HASH_SKEIN512, cw.COMP_LZ4, cw.MODE_COMPAT, src.data_ptr(), CHUNK, n, dig.data_ptr(),
                              blocks.data_ptr(), stride, lens.data_ptr(), stream)

    # parity spot check before timing: first and last 128 chunks of the shard against the oracle
    step()
    torch.cuda.synchronize()
    sel = np.r_[0:128, n - 128:n]
    src_h = torch.cat([src[: 128 * CHUNK], src[(n - 128) * CHUNK:]]).cpu().numpy()
    assert (src_h == np.concatenate([O.synth(CHUNK, first, 128, SEED), O.synth(CHUNK, first + n - 128, 128, SEED)])).all()
    dig_h = dig.view(n, DIGEST)[torch.as_tensor(sel, device=dev)].cpu().numpy()
    lens_h = lens[torch.as_tensor(sel, device=dev)].cpu().numpy().astype(np.uint32)
    blk_h = blocks.view(n, stride)[torch.as_tensor(sel, device=dev)].cpu().numpy()
    ob, ol = O.compress_chunks(O.COMP_LZ4, src_h, CHUNK)
    parity_ok = bool((dig_h == O.hash_chunks(O.HASH_SKEIN512, src_h, CHUNK)).all() and (lens_h == ol).all()
                     and all((blk_h[i, : ol[i]] == ob[i, : ol[i]]).all() for i in range(len(sel))))
    if not parity_ok:
        raise SystemExit("bench.py: parity check against the oracle FAILED; refusing to report a number")
    comp_bytes = float(lens.sum(dtype=torch.int64).item())
    ratio = comp_bytes / float(n * CHUNK)

    # size-independent property at the full shard size: every block decodes back to its chunk (GPU decoder,
    # itself pinned against the reference's streams in tests/), and every chunk got a length
    back = torch.empty(n * CHUNK, dtype=torch.uint8, device=dev)
    back_lens = torch.empty(n, dtype=torch.int32, device=dev)

    def roundtrip_all() -> bool:
        back.zero_()
        eng.decompress_dev(cw.COMP_LZ4, blocks.data_ptr(), stride, lens.data_ptr(), CHUNK, n, back.data_ptr(),
                           back_lens.data_ptr(), stream)
        torch.cuda.synchronize()
        return bool((back_lens == CHUNK).all().item()) and bool(torch.equal(back, src))

    roundtrip_compat = roundtrip_all()
    if not roundtrip_compat:
        raise SystemExit("bench.py: device round trip of the compat blocks FAILED; refusing to report a number")

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # nvidia-smi needs a few hundred ms to start: begin before the warm-up, all samples are under load
    for _ in range(max(args.warmup - 1, 0)):
        step()
    barrier()
    eng.prof_enable(True)
    eng.prof_read()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches0
    kernel_ms, kernel_launches = eng.prof_read()
    eng.prof_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    ms_all = max_over_ranks(ms)
    total_bytes = float(n * CHUNK) * world * args.steps
    value = total_bytes / (ms_all * 1e-3) / 1e9

    # ---- fast mode (CW_MODE_FAST: same parse over a 4096-entry table; valid LZ4, ratio within 1 %) as a side figure ----
    fast = None
    try:
        def step_fast():
            eng.hash_compress_dev(cw.HASH_SKEIN512, cw.COMP_LZ4, cw.MODE_FAST, src.data_ptr(), CHUNK, n, dig.data_ptr(),
                                  blocks.data_ptr(), stride, lens.data_ptr(), stream)
        step_fast()
        torch.cuda.synchronize()
        f_lens = lens[torch.as_tensor(sel, device=dev)].cpu().numpy().astype(np.uint32)
        f_blk = blocks.view(n, stride)[torch.as_tensor(sel, device=dev)].cpu().numpy()
        f_dig = dig.view(n, DIGEST)[torch.as_tensor(sel, device=dev)].cpu().numpy()
        fast_ok = bool(O.verify_roundtrip(O.COMP_LZ4, src_h, CHUNK, f_blk, f_lens) == 0 and (f_dig == dig_h).all()
                       and roundtrip_all())
        fast_ratio = float(lens.sum(dtype=torch.int64).item()) / float(n * CHUNK)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            step_fast()
        f1.record()
        barrier()
        fms = max_over_ranks(f0.elapsed_time(f1))
        fast = {"value": float(n * CHUNK) * world * args.steps / (fms * 1e-3) / 1e9, "unit": "GB/s",
                "compression_ratio": fast_ratio, "ratio_vs_compat": fast_ratio / ratio,
                "roundtrip_all_blocks_and_sample_digests_ok": fast_ok}
        step()  # leave compat results in the buffers
        torch.cuda.synchronize()
    except Exception as ex:  # noqa: BLE001
        fast = {"error": str(ex)}

    # ---- end to end through the host-buffer C ABI ----
    e2e = None
    if not args.no_e2e:
        n_e = n
        try:
            import psutil
            avail = psutil.virtual_memory().available
            while n_e > 16384 and (n_e * (CHUNK + stride + DIGEST + 4)) * max(world, 1) > 0.5 * avail:
                n_e //= 2
        except Exception:
            pass
        h_src = torch.empty(n_e * CHUNK, dtype=torch.uint8).pin_memory()
        h_src.copy_(src[: n_e * CHUNK])
        h_blocks = torch.empty(n_e * stride, dtype=torch.uint8).pin_memory()
        h_dig = torch.empty(n_e * DIGEST, dtype=torch.uint8).pin_memory()
        h_lens = torch.empty(n_e, dtype=torch.int32).pin_memory()
        torch.cuda.synchronize()

        def e2e_step():
            eng.process_block_ptr(cw.HASH_SKEIN512, cw.COMP_LZ4, cw.MODE_COMPAT, h_src.data_ptr(), CHUNK, n_e,
                                  h_dig.data_ptr(), h_blocks.data_ptr(), stride, h_lens.data_ptr())

        e2e_step()
        ok_e2e = bool((h_dig.view(n_e, DIGEST)[:128].numpy() == dig_h[:128]).all()
                      and (h_lens[:128].numpy().astype(np.uint32) == ol[:128]).all())
        if not ok_e2e:
            raise SystemExit("bench.py: e2e outputs differ from the device-resident pass")
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        t1 = time.perf_counter()
        barrier()
        sec_all = max_over_ranks(t1 - t0)
        e2e = {"value": float(n_e * CHUNK) * world * args.steps / sec_all / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": int(n_e * CHUNK) * world,
               "d2h_bytes_per_step": int(n_e * (stride + DIGEST + 4)) * world,
               "chunks_per_gpu": n_e, "timer": "host wall clock around the blocking cw_hash_compress call, max over ranks",
               "api": "cw_hash_compress (host pinned buffers, strided output)"}
        del h_src, h_blocks, h_dig, h_lens

    ratio_all = sum_over_ranks(comp_bytes) / (float(n * CHUNK) * world)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        alg_bytes = float(n * CHUNK) * (1.0 + ratio) + float(n * (DIGEST + 4))  # read input, write blocks+digests+lens
        k_ms = kernel_ms / max(kernel_launches, 1)
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("fused_bytes_per_launch")
        except Exception:
            pass
        # integer-pipe view of the same kernel: Skein-512 alone is 1090 64-bit ops per 64-byte block
        #  = 2180 int32 ops; 129 blocks per 8 KiB chunk
        int_ops = float(n) * 129 * 2180
        int_peak = None
        try:
            ub = json.load(open(os.path.join(ROOT, "profiles", "ubench_r01.json")))
            int_peak = ub["int32_alu_plus_fma_tops"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_all / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64/u8", "data": "synthetic",
            "config": {"workload": "fused skein512+lz4 (compat), 8 KiB chunks, synthetic mixed-entropy stream, "
                                   f"{n} chunks = {n * CHUNK / 2**30:.2f} GiB per GPU (16 GiB at 8 GPUs = BASELINE configs[4])",
                       "chunk_bytes": CHUNK, "chunks_per_gpu": n, "sharding": "contiguous chunk range per rank, no collective",
                       "l2": "inputs larger than L2 (2 GiB per pass vs 126 MB)", "compression_ratio": ratio_all,
                       "parity": "first/last 128 chunks of every shard bit-exact vs oracle before timing; "
                                 "all blocks of the shard decoded back on the device == input",
                       "roundtrip_all_blocks": roundtrip_compat, "host_numa_binding_rank0": numa},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "fast_mode": fast,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "cw_fused_kernel<skein512, lz4 compat> (the one kernel of a cw_hash_compress_dev call)",
                         "kernel_ms_per_launch": k_ms, "algorithmic_bytes_per_launch": alg_bytes,
                         "int_pipe": {"int32_ops_per_launch_skein_only": int_ops,
                                      "achieved_tops": int_ops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else None,
                                      "peak_tops": int_peak,
                                      "note": "the pass is integer-issue bound, not HBM bound (DESIGN.md)"}},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            ncpu = cpu_sample_chunks(threads)
            secs, kind, cratio = run_cpu_reference(ncpu, threads, 2)
            best = min(secs)
            line["cpu_baseline"] = {"value": ncpu * CHUNK / best / 1e9, "unit": "GB/s", "cores": threads, "kind": kind,
                                    "sample": f"{ncpu} chunks x {CHUNK} B prefix of the same synthetic stream, best of 2",
                                    "compression_ratio": cratio}
        print(json.dumps(line), flush=True)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
