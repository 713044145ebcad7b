"""Per-data-class throughput of the compress / hash kernels (device resident).
Data sets: the four classes of the synthetic stream separately, the mixed stream, and (if staged) the Canterbury
corpus tiled.  usage: python tools/prof_classes.py [lz4|lzf|fused|fusedlzf] [chunk]   (CW_MODE=1: fast mode)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import compute_war_b200 as cw
import oracle_lib as O

what = sys.argv[1] if len(sys.argv) > 1 else "lz4"
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
N = 32768
MODE = int(os.environ.get("CW_MODE", "0"))
L = O.lib()
L.cwo_synth_class.argtypes = [C.c_uint64, C.c_uint64]
L.cwo_synth_class.restype = C.c_uint
sets = {}
idx_by_class = {c: [] for c in range(4)}
i = 0
while min(len(v) for v in idx_by_class.values()) < 512:
    idx_by_class[L.cwo_synth_class(0xC0FFEE, i)].append(i)
    i += 1
names = {0: "synthetic P (pattern)", 1: "synthetic T (text)", 2: "synthetic R (records)", 3: "synthetic X (random)"}
for c, idxs in idx_by_class.items():
    base = np.concatenate([O.synth(chunk, k, 1) for k in idxs[:512]])
    sets[names[c]] = np.tile(base, N // 512)
sets["synthetic mixed"] = O.synth(chunk, 0, N)
files = O.corpus_files()
if files:
    corp = np.concatenate([np.fromfile(f, np.uint8)[: os.path.getsize(f) // chunk * chunk] for f in files])
    reps = max(1, N * chunk // len(corp))
    sets[f"canterbury+large tiled x{reps}"] = np.tile(corp, reps)
    for f in files:
        if os.path.basename(f) in ("bible.txt", "kennedy.xls", "ptt5", "world192.txt", "alice29.txt"):
            d = np.fromfile(f, np.uint8)
            d = d[: len(d) // chunk * chunk]
            sets[os.path.basename(f)] = np.tile(d, max(1, N * chunk // len(d)))

eng = cw.Engine(0)
dev = torch.device("cuda:0")
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
s = st.cuda_stream
comp = cw.COMP_LZF if what in ("lzf", "fusedlzf") else cw.COMP_LZ4  # fusedlzf: Skein-512 + LZF in one pass
stride = cw.compress_bound(comp, chunk)
for name, data in sets.items():
    n = len(data) // chunk
    src = torch.from_numpy(data[: n * chunk]).to(dev)
    blocks = torch.empty(n * stride, dtype=torch.uint8, device=dev)
    lens = torch.empty(n, dtype=torch.int32, device=dev)
    dig = torch.empty(n * 64, dtype=torch.uint8, device=dev)

    def run():
        if what in ("fused", "fusedlzf"):
            eng.hash_compress_dev(cw.HASH_SKEIN512, comp, MODE, src.data_ptr(), chunk, n, dig.data_ptr(), blocks.data_ptr(),
                                  stride, lens.data_ptr(), s)
        else:
            eng.compress_dev(comp, MODE, src.data_ptr(), chunk, n, blocks.data_ptr(), stride, lens.data_ptr(), s)

    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    lens_h = lens.cpu().numpy().astype(np.int64)
    ratio = float(np.where(lens_h == 0, chunk, lens_h).sum()) / (n * chunk)
    print(f"{what:8s} chunk={chunk:5d} {name:38s} {n * chunk / 2**20:7.0f} MiB  {ms:8.3f} ms  {n * chunk / ms / 1e6:8.1f} GB/s  "
          f"ratio {ratio:.4f}", flush=True)
eng.close()
