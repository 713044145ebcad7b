"""Multi-GPU split of the chunked path: contiguous chunk ranges, one process (and one cw_ctx) per GPU.

The reference shards chunks across worker threads through a queue (HashAndCompress.cpp:263-272, 398-403); every chunk
is independent (fresh Skein context per block :127-129, fresh codec state per call), so there is no exchange step on
the data path.  Rank r of W owns chunks [r*n/W, (r+1)*n/W).  The only collective is the OPTIONAL gather of the small
per-chunk results (digest + length: 68 B per 8 KiB chunk) so that every rank can compute global packed offsets;
compressed blocks stay where they were produced.  Works with any torch.distributed backend (NCCL on GPUs, gloo on CPU).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_chunks: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of rank's contiguous chunk range; ranges tile [0, n_chunks) exactly, sizes differ by at most one"""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return n_chunks * rank // world, n_chunks * (rank + 1) // world


def packed_offsets(lens: np.ndarray, chunk: int) -> np.ndarray:
    """exclusive prefix sum of stored sizes: a chunk whose codec returned 0 (LZF 'does not fit') is stored raw"""
    stored = np.where(lens == 0, chunk, lens).astype(np.int64)
    out = np.zeros(len(stored) + 1, np.int64)
    np.cumsum(stored, out=out[1:])
    return out


def gather_results(digests, lens, n_total: int, digest_bytes: int):
    """all-gather per-chunk digests [n_local, D] (uint8) and lens [n_local] (int32) from every rank, in rank order.
    Tensors must live where the backend expects them (CUDA for NCCL, CPU for gloo).  Returns (digests[n_total, D],
    lens[n_total]) on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    rank = dist.get_rank()
    lo, hi = shard_range(n_total, rank, world)
    assert digests.shape[0] == hi - lo and lens.shape[0] == hi - lo
    per = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    pad_d = torch.zeros((per, digest_bytes), dtype=torch.uint8, device=digests.device)
    pad_l = torch.zeros(per, dtype=torch.int32, device=lens.device)
    pad_d[: hi - lo] = digests
    pad_l[: hi - lo] = lens
    all_d = [torch.empty_like(pad_d) for _ in range(world)]
    all_l = [torch.empty_like(pad_l) for _ in range(world)]
    dist.all_gather(all_d, pad_d)
    dist.all_gather(all_l, pad_l)
    ds, ls = [], []
    for r in range(world):
        a, b = shard_range(n_total, r, world)
        ds.append(all_d[r][: b - a])
        ls.append(all_l[r][: b - a])
    return torch.cat(ds), torch.cat(ls)
