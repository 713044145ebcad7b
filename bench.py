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


CPU_CHECK = {}  # comp_bytes / digest_fnv of the last reference run (whole-sample check values)


def run_cpu_reference(n_chunks: int, threads: int, reps: int):
    """-> (list of seconds per rep, kind, ratio).  Uses oracle/_ref/ref_cli (its Skein compiled for x86-64-v3 when the
    host runs that) when present, else the C oracle port."""
    import oracle_lib as O
    if O.have_ref_cli():
        exe = O.ref_cli_bench_exe()
        out = subprocess.check_output([exe, "bench", "skein512", "lz4", str(CHUNK), str(n_chunks), str(threads), "0", hex(SEED),
                                       str(reps)], text=True)
        recs = [json.loads(l) for l in out.strip().splitlines() if l.startswith("{")]
        CPU_CHECK.update({"n_chunks": n_chunks, "comp_bytes": recs[-1].get("comp_bytes"), "digest_fnv": recs[-1].get("digest_fnv"),
                          "build": "x86-64-v3" if exe.endswith("_v3") else "portable -O2"})
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
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": threads, "kind": kind, "sample": sample,
                         "build": CPU_CHECK.get("build")},
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



# ------------------------------------------------------------------------------------------------
# the other BASELINE.json configs (rank 0, N = 1): short device-resident runs, each with its own parity check
# ------------------------------------------------------------------------------------------------

# algorithmic integer work per input byte: Skein-512 1090 ops64 = 2180 int32 ops per 64-byte block,
# Skein-256 574 ops64 per 32-byte block, SHA-256 ~1900 int32 ops per 64-byte compression; + 1 output / padding block
def int_ops_per_chunk(hash_name: str, chunk: int) -> float:
    if hash_name == "skein512":
        return ((chunk + 63) // 64 + 1) * 2180.0
    if hash_name == "skein256-128":
        return ((chunk + 31) // 32 + 1) * 1148.0
    return ((chunk + 9 + 63) // 64) * 1900.0


def fnv_of_bytes(b) -> int:
    """ref_cli's digest check value: h = h * 1099511628211 + byte over all digest bytes, mod 2^64, h0 = 0"""
    import numpy as np
    a = np.ascontiguousarray(b, np.uint8).ravel().astype(np.uint64)
    n = len(a)
    P = np.uint64(1099511628211)
    with np.errstate(over="ignore"):
        pw = np.empty(n + 1, np.uint64)
        pw[0] = 1
        # powers by doubling: pw[i] = P^i
        filled = 1
        while filled < n + 1:
            m = min(filled, n + 1 - filled)
            pw[filled: filled + m] = pw[:m] * pw[filled - 1] * P
            filled += m
        h = (a * pw[n - 1:: -1][:n]).sum(dtype=np.uint64) if n else np.uint64(0)
    return int(h)


def run_configs(eng, cw, O, torch, np, dev, stream, peaks, int_peak, steps: int):
    """-> list of per-config dicts.  Throughput is device resident (CUDA events on `stream`, `steps` launches after one
    warm-up); every entry carries the parity check that was run on its own data before timing."""
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    out = []
    HN = {cw.HASH_SKEIN512: "skein512", cw.HASH_SKEIN256_128: "skein256-128", cw.HASH_SHA256: "sha256"}
    CN = {cw.COMP_LZ4: "lz4", cw.COMP_LZF: "lzf"}

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps

    def corpus(which, chunk, min_bytes):
        files = [f for f in O.corpus_files() if os.path.basename(os.path.dirname(f)) in which]
        if not files:
            return None, 0, 0
        d = np.concatenate([np.fromfile(f, np.uint8)[: os.path.getsize(f) // chunk * chunk] for f in files])  # whole chunks only
        reps = max(1, -(-min_bytes // len(d)))
        return np.tile(d, reps), reps, len(d) // chunk

    def hash_entry(name, data_h, chunk, halg, extra):
        n = len(data_h) // chunk
        src = torch.from_numpy(data_h[: n * chunk]).to(dev)
        D = cw.digest_bytes(halg)
        dig = torch.empty(n * D, dtype=torch.uint8, device=dev)
        ms = timed(lambda: eng.hash_dev(halg, src.data_ptr(), chunk, n, dig.data_ptr(), stream))
        k = min(n, 512)
        ok = bool((dig.view(n, D)[:k].cpu().numpy() == O.hash_chunks(halg, data_h[: k * chunk], chunk)).all())
        gbs = n * chunk / ms / 1e6
        tops = n * int_ops_per_chunk(HN[halg], chunk) / ms / 1e9
        e = {"config": name, "hash": HN[halg], "chunk_bytes": chunk, "chunks": n, "GBps": gbs, "ms": ms,
             "parity_first_chunks_vs_oracle": ok, "roofline": {"bound": "int_pipe", "achieved_tops": tops, "peak_tops": int_peak,
                                                               "frac": tops / int_peak if int_peak else None}}
        e.update(extra)
        out.append(e)

    def comp_entry(name, data_h, chunk, calg, mode, halg, extra):
        n = len(data_h) // chunk
        src = torch.from_numpy(data_h[: n * chunk]).to(dev)
        stride = cw.compress_bound(calg, chunk)
        blocks = torch.empty(n * stride, dtype=torch.uint8, device=dev)
        lens = torch.empty(n, dtype=torch.int32, device=dev)
        D = cw.digest_bytes(halg) if halg is not None else 0
        dig = torch.empty(max(n * D, 1), dtype=torch.uint8, device=dev)
        if halg is None:
            fn = lambda: eng.compress_dev(calg, mode, src.data_ptr(), chunk, n, blocks.data_ptr(), stride, lens.data_ptr(), stream)
        else:
            fn = lambda: eng.hash_compress_dev(halg, calg, mode, src.data_ptr(), chunk, n, dig.data_ptr(), blocks.data_ptr(), stride,
                                               lens.data_ptr(), stream)
        ms = timed(fn)
        k = min(n, 64 if chunk > 8192 else 256)
        lens_h = lens[:k].cpu().numpy().astype(np.uint32)
        blk_h = blocks.view(n, stride)[:k].cpu().numpy()
        samp = data_h[: k * chunk]
        ol = O.compress_chunks(calg, samp, chunk)[1]
        if mode == cw.MODE_COMPAT:
            ob = O.compress_chunks(calg, samp, chunk)[0]
            ok = bool((lens_h == ol).all() and all((blk_h[i, : ol[i]] == ob[i, : ol[i]]).all() for i in range(k)))
            how = "first chunks byte-identical to the oracle (= reference library output)"
        else:
            ok = O.verify_roundtrip(calg, samp, chunk, blk_h, lens_h) == 0
            how = "first chunks decode back (restated decoder)"
            if calg == cw.COMP_LZF:  # length 0 = kept raw (lzf.h:59-64): the decoders take such a chunk from its block
                blk_h = blk_h.copy()
                for i in np.nonzero(lens_h == 0)[0]:
                    blk_h[i, :chunk] = samp[i * chunk: (i + 1) * chunk]
            if O.have_ref_cli():
                back = O.ref_decompress_chunks(calg, chunk, blk_h, lens_h)
                ok = ok and len(back) == len(samp) and bool((back == samp).all())
                how = "first chunks restored by the reference's own decompressor (ref_cli)"
            fb, fl = O.lz4_fast_chunks(samp, chunk) if calg == cw.COMP_LZ4 else O.lzf_fast_chunks(samp, chunk)
            ok = ok and bool((lens_h == fl).all() and all((blk_h[i, : fl[i]] == fb[i, : fl[i]]).all() for i in range(k)))
            how += " and byte-identical to the restated fast parse"
        if halg is not None:
            ok = ok and bool((dig.view(n, D)[:k].cpu().numpy() == O.hash_chunks(halg, samp, chunk)).all())
        # all blocks of the set decode back on the device
        back = torch.empty(n * chunk, dtype=torch.uint8, device=dev)
        bl = torch.empty(n, dtype=torch.int32, device=dev)
        if calg == cw.COMP_LZF:  # chunks that did not shrink (length 0) are kept raw: put them where the decoder looks
            raw = (lens == 0).nonzero().flatten()
            if raw.numel():
                blocks.view(n, stride)[raw, :chunk] = src.view(n, chunk)[raw]
        eng.decompress_dev(calg, blocks.data_ptr(), stride, lens.data_ptr(), chunk, n, back.data_ptr(), bl.data_ptr(), stream)
        torch.cuda.synchronize()
        rt = bool((bl == chunk).all().item()) and bool(torch.equal(back, src))
        lh = lens.cpu().numpy().astype(np.int64)
        ratio = float(np.where(lh == 0, chunk, lh).sum()) / (n * chunk)
        ref_ratio_sample = float(np.where(ol == 0, chunk, ol.astype(np.int64)).sum()) / (k * chunk)
        my_ratio_sample = float(np.where(lens_h == 0, chunk, lens_h.astype(np.int64)).sum()) / (k * chunk)
        gbs = n * chunk / ms / 1e6
        alg_bpb = 1.0 + ratio + (4 + D) / chunk
        e = {"config": name, "comp": CN[calg], "mode": "fast" if mode else "compat", "chunk_bytes": chunk, "chunks": n,
             "GBps": gbs, "ms": ms, "compression_ratio": ratio, "size_vs_reference_on_sample": my_ratio_sample / ref_ratio_sample,
             "parity": ok, "parity_how": how, "device_roundtrip_all_blocks": rt}
        if halg is not None:
            tops = n * int_ops_per_chunk(HN[halg], chunk) / ms / 1e9
            e["hash"] = HN[halg]
            e["roofline"] = {"bound": "int_pipe", "achieved_tops": tops, "peak_tops": int_peak, "frac": tops / int_peak if int_peak else None,
                             "hbm": {"achieved": gbs * alg_bpb, "peak": hbm_peak, "frac": gbs * alg_bpb / hbm_peak}}
        else:
            e["roofline"] = {"bound": "hbm", "achieved": gbs * alg_bpb, "peak": hbm_peak, "frac": gbs * alg_bpb / hbm_peak}
        e.update(extra)
        out.append(e)
        del src, blocks, lens, dig, back, bl

    GiB = 1 << 30
    # configs[0/1]: Skein-512 (and the tools' Skein-256/128) digests of 4 KiB chunks over dataset/canterbury
    d, reps, per = corpus(("canterbury",), 4096, GiB // 2)
    if d is not None:
        for h in (cw.HASH_SKEIN512, cw.HASH_SKEIN256_128):
            hash_entry("configs[0]: 4 KiB chunks of dataset/canterbury", d, 4096, h, {"corpus_chunks": per, "tiled_x": reps})
    # configs[1]: Skein-512 + SHA-256 over dataset/canterbury-large
    d, reps, per = corpus(("canterbury-large",), 8192, GiB)
    if d is not None:
        for h in (cw.HASH_SKEIN512, cw.HASH_SHA256):
            hash_entry("configs[1]: 8 KiB chunks of dataset/canterbury-large (E.coli absent from the mount)", d, 8192, h,
                       {"corpus_chunks": per, "tiled_x": reps})
    # configs[2]: LZ4, independent 64 KiB blocks, compat + fast (Silesia is not in the mount: Canterbury + large stand in)
    d, reps, per = corpus(("canterbury", "canterbury-large"), 65536, GiB)
    syn64 = O.synth(65536, 0, 8192)
    for mode in (cw.MODE_COMPAT, cw.MODE_FAST):
        if d is not None:
            comp_entry("configs[2]: LZ4 64 KiB blocks, canterbury + canterbury-large (Silesia absent)", d, 65536, cw.COMP_LZ4, mode, None,
                       {"corpus_chunks": per, "tiled_x": reps})
        comp_entry("configs[2]: LZ4 64 KiB blocks, synthetic mix", syn64, 65536, cw.COMP_LZ4, mode, None, {})
    # configs[3]: fused hc_sklz4 / hc_sklzf / hc_shlz4 at 8 KiB, real corpus tiled and the synthetic mix
    d, reps, per = corpus(("canterbury", "canterbury-large"), 8192, GiB // 4)
    syn8 = O.synth(8192, 0, 32768)
    for halg, calg, modes in ((cw.HASH_SKEIN512, cw.COMP_LZ4, (0, 1)), (cw.HASH_SKEIN512, cw.COMP_LZF, (0, 1)),
                              (cw.HASH_SHA256, cw.COMP_LZ4, (0, 1))):
        tag = "hc_" + {"skein512": "sk", "sha256": "sh"}[HN[halg]] + CN[calg]
        for mode in modes:
            if d is not None:
                comp_entry(f"configs[3]: {tag} 8 KiB, canterbury + canterbury-large", d, 8192, calg, mode, halg,
                           {"corpus_chunks": per, "tiled_x": reps})
            comp_entry(f"configs[3]: {tag} 8 KiB, synthetic mix", syn8, 8192, calg, mode, halg, {})
    return out

def run_small_items(O, np):
    """The reference's calling pattern through the product's own host tool (compute_war_b200/hashandcompress --work-items:
    --c-threads workers, one cw_hash_compress call -- or HashOffload-style cw_submit / cw_wait -- per work item of
    readBlockFactor chunks; HashAndCompress.cpp:160-183,263-272, run_tests:14), from ordinary pageable buffers like the
    reference's, 256 MiB of the synthetic stream in 4 KiB chunks.  Reported next to the whole-file call."""
    exe = os.path.join(ROOT, "compute_war_b200", "hashandcompress")
    if not os.access(exe, os.X_OK):
        return {"error": "host tool not built"}
    import re
    import tempfile
    chunk, n = 4096, 65536
    rows = []
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "items.bin")
        O.synth(chunk, 0, n).tofile(path)

        def run(label, *flags):
            try:
                out = subprocess.run([exe, "-H", "skein512", "-C", "lz4", "-k", str(chunk), "-i", "3", *flags, path],
                                     capture_output=True, text=True, timeout=120)
                m = re.search(r"MBps=([0-9.]+)", out.stdout + out.stderr)
                rows.append({"call": label, "GBps": float(m.group(1)) * 1048576 / 1e9 if m else None})  # the tool prints MiB/s
            except Exception as ex:  # noqa: BLE001 -- a side figure never fails the bench line
                rows.append({"call": label, "error": str(ex)[:120]})

        run("one call per file, page-locked buffers")
        run("one call per file, pageable buffers", "--no-pin")
        for r in (8, 64, 256, 2048):
            run(f"blocking calls of {r * chunk // 1024} KiB from 14 threads, pageable", "--no-pin", "--work-items", "-c", "14", "-r", str(r))
        for r in (8, 64):
            run(f"HashOffload style (cw_submit / cw_wait, 64 tickets per thread) items of {r * chunk // 1024} KiB from 14 threads, pageable",
                "--no-pin", "--work-items", "--gpu-offload", "1", "-c", "14", "-r", str(r))
    return {"workload": f"{n} chunks x {chunk} B of the synthetic stream, fused skein512+lz4 compat, host tool, best-effort side figure",
            "host_threads": host_threads(), "rows": rows}



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunks-per-gpu", type=int, default=CHUNKS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config table (BASELINE configs 0-3)")
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
        # stdout carries ONE JSON line: NCCL's own lines (the "NCCL version ..." banner it prints at NCCL_DEBUG=VERSION and
        # above, warnings) go to stderr
        # (measured on the box: with NCCL_DEBUG unset / VERSION the banner lands on stdout whatever NCCL_DEBUG_FILE says;
        # at WARN it follows NCCL_DEBUG_FILE)
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
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
        fast_how = "all blocks decoded back on the device; first/last 128 chunks: restated decoder"
        if O.have_ref_cli():  # the reference's own LZ4_decompress_safe (oracle/_ref/ref_cli, built from its liblz4.a)
            rb = O.ref_decompress_chunks(O.COMP_LZ4, CHUNK, f_blk, f_lens)
            fast_ok = fast_ok and len(rb) == len(src_h) and bool((rb == src_h).all())
            fast_how += " + the reference's LZ4_decompress_safe (ref_cli)"
        rs_b, rs_l = O.lz4_fast_chunks(src_h, CHUNK)  # the fast parse is deterministic: byte-equal to its restatement
        fast_ok = fast_ok and bool((f_lens == rs_l).all() and all((f_blk[i, : rs_l[i]] == rs_b[i, : rs_l[i]]).all()
                                                                 for i in range(len(sel))))
        fast_how += "; byte-identical to the restated fast parse"
        fast_ratio = float(lens.sum(dtype=torch.int64).item()) / float(n * CHUNK)
        barrier()
        eng.prof_enable(True)
        eng.prof_read()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            step_fast()
        f1.record()
        barrier()
        fk_ms, fk_n = eng.prof_read()
        eng.prof_enable(False)
        fms = max_over_ranks(f0.elapsed_time(f1))
        fast = {"value": float(n * CHUNK) * world * args.steps / (fms * 1e-3) / 1e9, "unit": "GB/s",
                "compression_ratio": fast_ratio, "ratio_vs_compat": fast_ratio / ratio,
                "roundtrip_all_blocks_and_sample_digests_ok": fast_ok, "verified_by": fast_how,
                "ms_per_step": fms / args.steps, "kernel_ms_per_launch": fk_ms / max(fk_n, 1)}
        step()  # leave compat results in the buffers
        torch.cuda.synchronize()
    except Exception as ex:  # noqa: BLE001
        fast = {"error": str(ex)}

    # ---- end to end through the host-buffer C ABI ----
    # headline: cw_hash_compress_packed (blocks back to back: the device-to-host leg carries the bytes that exist);
    # beside it the strided call, whose D2H moves whole strides (1.5 x the block bytes)
    e2e = None
    if not args.no_e2e:
        n_e = n
        try:
            import psutil
            avail = psutil.virtual_memory().available
            while n_e > 16384 and (n_e * (CHUNK + stride + DIGEST + 16)) * max(world, 1) > 0.5 * avail:
                n_e //= 2
        except Exception:
            pass
        h_src = torch.empty(n_e * CHUNK, dtype=torch.uint8).pin_memory()
        h_src.copy_(src[: n_e * CHUNK])
        h_blocks = torch.empty(n_e * stride, dtype=torch.uint8).pin_memory()  # strided output / packed output (upper bound)
        h_dig = torch.empty(n_e * DIGEST, dtype=torch.uint8).pin_memory()
        h_lens = torch.empty(n_e, dtype=torch.int32).pin_memory()
        h_off = torch.empty(n_e + 1, dtype=torch.int64).pin_memory()
        torch.cuda.synchronize()

        def e2e_strided():
            eng.process_block_ptr(cw.HASH_SKEIN512, cw.COMP_LZ4, cw.MODE_COMPAT, h_src.data_ptr(), CHUNK, n_e,
                                  h_dig.data_ptr(), h_blocks.data_ptr(), stride, h_lens.data_ptr())

        def e2e_packed():
            eng.process_block_packed_ptr(cw.HASH_SKEIN512, cw.COMP_LZ4, cw.MODE_COMPAT, h_src.data_ptr(), CHUNK, n_e,
                                         h_dig.data_ptr(), h_blocks.data_ptr(), n_e * stride, h_off.data_ptr(), h_lens.data_ptr())

        def timed_e2e(fn):
            fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                fn()
            t1 = time.perf_counter()
            barrier()
            return max_over_ranks(t1 - t0)

        e2e_strided()
        ok_e2e = bool((h_dig.view(n_e, DIGEST)[:128].numpy() == dig_h[:128]).all()
                      and (h_lens[:128].numpy().astype(np.uint32) == ol[:128]).all())
        sec_s = timed_e2e(e2e_strided)
        e2e_packed()
        offs = h_off.numpy()
        lens_e = h_lens.numpy().astype(np.int64)
        ok_e2e = ok_e2e and bool((np.diff(offs) == lens_e).all()) and int(offs[0]) == 0
        ok_e2e = ok_e2e and bool((h_dig.view(n_e, DIGEST)[:128].numpy() == dig_h[:128]).all())
        ok_e2e = ok_e2e and all((h_blocks.numpy()[offs[i]: offs[i] + ol[i]] == ob[i, : ol[i]]).all() for i in range(128))
        if not ok_e2e:
            raise SystemExit("bench.py: e2e outputs differ from the device-resident pass")
        packed_bytes = int(offs[n_e])
        sec_p = timed_e2e(e2e_packed)

        # the copy ceiling of this box for exactly these bytes: the same pinned buffers, 64 MiB pieces, copy-in and
        # copy-out at once on two streams, no kernel -- all ranks together, slowest rank counts
        d_a = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
        d_b = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
        cs1, cs2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

        def copy_only(out_bytes):
            piece = 64 << 20
            for off in range(0, n_e * CHUNK, piece):
                k = min(piece, n_e * CHUNK - off)
                with torch.cuda.stream(cs1):
                    d_a[:k].copy_(h_src[off:off + k], non_blocking=True)
            for off in range(0, out_bytes, piece):
                k = min(piece, out_bytes - off)
                with torch.cuda.stream(cs2):
                    h_blocks[off:off + k].copy_(d_b[:k], non_blocking=True)
            torch.cuda.synchronize()

        d2h_packed = int(packed_bytes + n_e * (DIGEST + 4 + 8) + 8)
        d2h_strided = int(n_e * (stride + DIGEST + 4))
        ceil_p = timed_e2e(lambda: copy_only(min(d2h_packed, n_e * stride)))
        ceil_s = timed_e2e(lambda: copy_only(min(d2h_strided, n_e * stride)))
        torch.cuda.set_stream(tstream)
        gb = float(n_e * CHUNK) * world * args.steps / 1e9
        e2e = {"value": gb / sec_p, "unit": "GB/s",
               "h2d_bytes_per_step": int(n_e * CHUNK) * world,
               "d2h_bytes_per_step": int(packed_bytes + n_e * (DIGEST + 4 + 8) + 8) * world,
               "chunks_per_gpu": n_e, "timer": "host wall clock around the blocking call, max over ranks",
               "api": "cw_hash_compress_packed (pinned host input; packed blocks + u64 offsets + lens + digests out)",
               "copy_ceiling": {"value": gb / ceil_p, "unit": "GB/s", "frac_of_ceiling": ceil_p / sec_p,
                                "how": "the same H2D + D2H bytes from / to the same pinned buffers with no kernel in between, "
                                       "all ranks at once (64 MiB pieces, two streams), slowest rank"},
               "strided": {"value": gb / sec_s, "unit": "GB/s", "api": "cw_hash_compress (strided output)",
                           "d2h_bytes_per_step": int(n_e * (stride + DIGEST + 4)) * world,
                           "copy_ceiling": {"value": gb / ceil_s, "unit": "GB/s", "frac_of_ceiling": ceil_s / sec_s}}}
        del h_src, h_blocks, h_dig, h_lens, h_off, d_a, d_b

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
        int_tops = int_ops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else None
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
            # the fused pass is bounded by the integer pipes, not by HBM; the HBM view of the same launch
            # (what the kernel moves against the measured copy peak) is kept beside it
            "roofline": {"bound": "int_pipe", "achieved": int_tops, "peak": int_peak, "unit": "Tops/s (int32)",
                         "frac": (int_tops / int_peak) if (int_tops and int_peak) else None,
                         "peak_source": "tools/ubench.cu on this pool's B200: ALU + FMA pipes dual-issued (profiles/ubench_r01.json)",
                         "algorithmic_ops_per_launch": int_ops,
                         "ops_counted": "Skein-512 only: 2180 int32 ops per 64-byte block x 129 blocks per chunk (SURVEY.md 8d); "
                                        "the LZ4 parse adds integer work that has no algorithmic count",
                         "kernel": "cw_fused_kernel<skein512, lz4 compat> (the one kernel of a cw_hash_compress_dev call)",
                         "kernel_ms_per_launch": k_ms,
                         "traffic": traffic,
                         "traffic_source": "static: dram__bytes_read + dram__bytes_write of ONE ncu --set full capture of this "
                                           "kernel on this shard (profiles/traffic.json), not measured in this run",
                         "hbm": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                                 "frac": (achieved / peak) if achieved else None, "peak_source": peak_src,
                                 "algorithmic_bytes_per_launch": alg_bytes}},
        }
        if fast and "value" in fast and fast.get("kernel_ms_per_launch"):
            f_ms = fast["kernel_ms_per_launch"]
            f_bytes = float(n * CHUNK) * (1.0 + fast["compression_ratio"]) + float(n * (DIGEST + 4))
            f_tops = int_ops / (f_ms * 1e-3) / 1e12
            fast["roofline"] = {"bound": "int_pipe", "achieved": f_tops, "peak": int_peak, "unit": "Tops/s (int32)",
                                "frac": f_tops / int_peak if int_peak else None,
                                "kernel": "cw_fused_kernel<skein512, lz4 fast> (teams of warps per chunk)",
                                "hbm": {"achieved": f_bytes / (f_ms * 1e-3) / 1e9, "peak": peak,
                                        "frac": f_bytes / (f_ms * 1e-3) / 1e9 / peak}}
        if world == 1 and not args.no_configs:
            try:
                line["configs"] = run_configs(eng, cw, O, torch, np, dev, stream, peaks, int_peak, max(3, min(args.steps, 5)))
            except Exception as ex:  # noqa: BLE001
                line["configs"] = {"error": f"{type(ex).__name__}: {ex}"}
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            ncpu = cpu_sample_chunks(threads)
            secs, kind, cratio = run_cpu_reference(ncpu, threads, 2)
            best = min(secs)
            line["cpu_baseline"] = {"value": ncpu * CHUNK / best / 1e9, "unit": "GB/s", "cores": threads, "kind": kind,
                                    "sample": f"{ncpu} chunks x {CHUNK} B prefix of the same synthetic stream, best of 2",
                                    "compression_ratio": cratio, "build": CPU_CHECK.get("build")}
            # the GPU's compat results over the same prefix against the reference run's own check values: every digest
            # byte (polynomial check value) and the total compressed size -- covers all chunks of the sample, not 256
            if CPU_CHECK.get("digest_fnv") and ncpu <= n:
                g_fnv = fnv_of_bytes(dig[: ncpu * DIGEST].cpu().numpy())
                g_bytes = int(lens[:ncpu].sum(dtype=torch.int64).item())
                same = (f"{g_fnv:016x}" == CPU_CHECK["digest_fnv"]) and (CPU_CHECK.get("comp_bytes") in (None, g_bytes))
                line["config"]["whole_sample_vs_reference_run"] = {
                    "chunks": ncpu, "digest_check_value_equal": f"{g_fnv:016x}" == CPU_CHECK["digest_fnv"],
                    "compressed_bytes_equal": CPU_CHECK.get("comp_bytes") == g_bytes, "compressed_bytes": g_bytes}
                if not same:
                    raise SystemExit("bench.py: digests / compressed size of the GPU pass differ from the reference run")
            s1, _, _ = run_cpu_reference(min(ncpu, 8192), 1, 2)
            line["cpu_baseline"]["one_thread"] = {"value": min(ncpu, 8192) * CHUNK / min(s1) / 1e9, "unit": "GB/s", "cores": 1}
            # for honesty: the same job with the reference's own prebuilt, unoptimised libskein.a
            shipped = os.path.join(ROOT, "oracle", "_ref", "ref_cli_shipped")
            if os.access(shipped, os.X_OK):
                try:
                    thr = line["cpu_baseline"]["cores"]
                    ns = min(ncpu, 32768)
                    out = subprocess.check_output([shipped, "bench", "skein512", "lz4", str(CHUNK), str(ns), str(thr), "0", hex(SEED), "2"],
                                                  text=True, timeout=300)
                    recs = [json.loads(l) for l in out.strip().splitlines() if l.startswith("{")]
                    line["cpu_baseline"]["as_shipped_libskein"] = {
                        "value": ns * CHUNK / min(r["seconds"] for r in recs) / 1e9, "unit": "GB/s", "cores": thr,
                        "what": "same driver linked against src/hashing_perf/skein/libskein.a as the reference ships it"}
                except Exception as ex:  # never fail the bench line over the side figure
                    line["cpu_baseline"]["as_shipped_libskein"] = {"error": str(ex)[:200]}
        if world == 1 and not args.no_configs:
            try:
                line["small_items"] = run_small_items(O, np)
            except Exception as ex:  # noqa: BLE001
                line["small_items"] = {"error": f"{type(ex).__name__}: {ex}"}
        print(json.dumps(line), flush=True)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
