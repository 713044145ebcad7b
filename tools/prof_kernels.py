"""Small driver for ncu: runs each kernel a few times over a 128 MiB synthetic shard (device resident).
usage: python tools/prof_kernels.py [lz4|lzf|skein512|skein256|sha256|fused|all] [n_chunks] [chunk]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import compute_war_b200 as cw

what = sys.argv[1] if len(sys.argv) > 1 else "all"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
eng = cw.Engine(0)
dev = torch.device("cuda:0")
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
s = st.cuda_stream
src = torch.empty(n * chunk, dtype=torch.uint8, device=dev)
stride = cw.compress_bound(cw.COMP_LZ4, chunk)
blocks = torch.empty(n * stride, dtype=torch.uint8, device=dev)
lens = torch.empty(n, dtype=torch.int32, device=dev)
dig = torch.empty(n * 64, dtype=torch.uint8, device=dev)
eng.synth_dev(src.data_ptr(), chunk, 0, n, 0xC0FFEE, s)
torch.cuda.synchronize()
runs = {
    "lz4": lambda: eng.compress_dev(cw.COMP_LZ4, 0, src.data_ptr(), chunk, n, blocks.data_ptr(), stride, lens.data_ptr(), s),
    "lzf": lambda: eng.compress_dev(cw.COMP_LZF, 0, src.data_ptr(), chunk, n, blocks.data_ptr(), stride, lens.data_ptr(), s),
    "skein512": lambda: eng.hash_dev(cw.HASH_SKEIN512, src.data_ptr(), chunk, n, dig.data_ptr(), s),
    "skein256": lambda: eng.hash_dev(cw.HASH_SKEIN256_128, src.data_ptr(), chunk, n, dig.data_ptr(), s),
    "sha256": lambda: eng.hash_dev(cw.HASH_SHA256, src.data_ptr(), chunk, n, dig.data_ptr(), s),
    "fused": lambda: eng.hash_compress_dev(cw.HASH_SKEIN512, cw.COMP_LZ4, 0, src.data_ptr(), chunk, n, dig.data_ptr(),
                                           blocks.data_ptr(), stride, lens.data_ptr(), s),
    "fusedlzf": lambda: eng.hash_compress_dev(cw.HASH_SKEIN512, cw.COMP_LZF, 0, src.data_ptr(), chunk, n, dig.data_ptr(),
                                              blocks.data_ptr(), stride, lens.data_ptr(), s),
    "fusedsha": lambda: eng.hash_compress_dev(cw.HASH_SHA256, cw.COMP_LZ4, 0, src.data_ptr(), chunk, n, dig.data_ptr(),
                                              blocks.data_ptr(), stride, lens.data_ptr(), s),
}
names = list(runs) if what == "all" else what.split(",")
for name in names:
    try:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        runs[name]()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            runs[name]()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{name}: {ms:.3f} ms per pass over {n * chunk / 2**20:.0f} MiB = {n * chunk / ms / 1e6:.1f} GB/s", flush=True)
    except Exception as ex:  # noqa: BLE001
        print(f"{name}: {ex}", flush=True)
eng.close()
