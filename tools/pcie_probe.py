"""PCIe probe: H2D alone, D2H alone, both at once (pinned host memory, 64 MiB pieces on separate streams)."""
import time
import torch

dev = torch.device("cuda:0")
N = 2 << 30
piece = 64 << 20
h_in = torch.empty(N, dtype=torch.uint8).pin_memory()
h_out = torch.empty(N, dtype=torch.uint8).pin_memory()
d_a = torch.empty(N, dtype=torch.uint8, device=dev)
d_b = torch.empty(N, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, frac=1.0):
    torch.cuda.synchronize()
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
    return time.perf_counter() - t0


for name, a in (("h2d only", (True, False)), ("d2h only", (False, True)), ("both", (True, True)), ("both, d2h 70%", (True, True, 0.7)),
                ("both, d2h 1%", (True, True, 0.01))):
    run(*a)
    t = min(run(*a) for _ in range(3))
    print(f"{name:16s} {t * 1e3:8.2f} ms   {N / t / 1e9:6.1f} GB/s per direction (2 GiB)")
