"""ncu driver: LZ4/LZF/fused kernel over ONE data class.  usage: prof_one.py <lz4|lzf|fused> <P|T|R|X|mixed|file:NAME> [n_chunks] [chunk]"""
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

what, cls = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 8192
MODE = int(os.environ.get("CW_MODE", "0"))  # 1: fast mode
L = O.lib()
L.cwo_synth_class.argtypes = [C.c_uint64, C.c_uint64]
L.cwo_synth_class.restype = C.c_uint
if cls == "mixed":
    data = O.synth(chunk, 0, N)
elif cls.startswith("file:"):
    f = [p for p in O.corpus_files() if os.path.basename(p) == cls[5:]][0]
    d = np.fromfile(f, np.uint8)
    d = d[: len(d) // chunk * chunk]
    data = np.tile(d, max(1, N * chunk // len(d) + 1))[: N * chunk]
else:
    want = "PTRX".index(cls)
    idxs, i = [], 0
    while len(idxs) < 256:
        if L.cwo_synth_class(0xC0FFEE, i) == want:
            idxs.append(i)
        i += 1
    data = np.tile(np.concatenate([O.synth(chunk, k, 1) for k in idxs]), N // 256)
eng = cw.Engine(0)
dev = torch.device("cuda:0")
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
s = st.cuda_stream
comp = cw.COMP_LZF if what in ("lzf", "fusedlzf") else cw.COMP_LZ4
stride = cw.compress_bound(comp, chunk)
n = len(data) // chunk
src = torch.from_numpy(data).to(dev)
blocks = torch.empty(n * stride, dtype=torch.uint8, device=dev)
lens = torch.empty(n, dtype=torch.int32, device=dev)
dig = torch.empty(n * 64, dtype=torch.uint8, device=dev)
for _ in range(2):
    if what in ("fused", "fusedlzf"):
        eng.hash_compress_dev(cw.HASH_SKEIN512, comp, MODE, src.data_ptr(), chunk, n, dig.data_ptr(), blocks.data_ptr(), stride,
                              lens.data_ptr(), s)
    else:
        eng.compress_dev(comp, MODE, src.data_ptr(), chunk, n, blocks.data_ptr(), stride, lens.data_ptr(), s)
    torch.cuda.synchronize()
print("ok", what, cls, n)
eng.close()
