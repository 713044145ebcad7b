"""e2e (host pinned buffers -> cw_hash_compress) throughput for the pipeline knobs CW_SLICE_MB / CW_SLOTS."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import compute_war_b200 as cw
n, chunk = 262144, 8192
mode = int(os.environ.get("CW_MODE", "0"))
eng = cw.Engine(0)
dev = torch.device("cuda:0")
src = torch.empty(n * chunk, dtype=torch.uint8, device=dev)
eng.synth_dev(src.data_ptr(), chunk, 0, n); eng.sync()
stride = cw.compress_bound(cw.COMP_LZ4, chunk)
h_src = torch.empty(n * chunk, dtype=torch.uint8).pin_memory(); h_src.copy_(src)
h_blocks = torch.empty(n * stride, dtype=torch.uint8).pin_memory()
h_dig = torch.empty(n * 64, dtype=torch.uint8).pin_memory()
h_lens = torch.empty(n, dtype=torch.int32).pin_memory()
torch.cuda.synchronize()
def step():
    eng.process_block_ptr(cw.HASH_SKEIN512, cw.COMP_LZ4, mode, h_src.data_ptr(), chunk, n, h_dig.data_ptr(), h_blocks.data_ptr(), stride, h_lens.data_ptr())
step(); step()
t0 = time.perf_counter()
for _ in range(5): step()
dt = (time.perf_counter() - t0) / 5
print(f"slice_mb={os.environ.get("CW_SLICE_MB","64")} slots={os.environ.get("CW_SLOTS","4")} mode={mode}: {dt*1e3:.1f} ms/pass = {n*chunk/dt/1e9:.1f} GB/s e2e")
# plain copies for reference
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); src.copy_(h_src, non_blocking=True); e1.record(); torch.cuda.synchronize()
print(f"  H2D 2 GiB alone: {n*chunk/e0.elapsed_time(e1)/1e6:.1f} GB/s")
