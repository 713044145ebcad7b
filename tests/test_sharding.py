"""CPU tests of the multi-GPU split (world_size 2, gloo): contiguous chunk ranges tile the batch, per-rank results
gathered in rank order equal the single-process result, packed offsets follow the stored sizes.  The per-rank compute
is done by the oracle here (no GPU in this container); on the GPU box the same split is exercised by bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from compute_war_b200.sharding import gather_results, packed_offsets, shard_range


def test_shard_ranges_tile_exactly():
    for n in (0, 1, 7, 8, 262144, 2097152, 1133):
        for w in (1, 2, 3, 4, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    # BASELINE configs[4]: 16 GiB / 8 KiB over 8 GPUs
    assert shard_range(2097152, 3, 8) == (786432, 1048576)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_packed_offsets_store_raw_when_codec_returns_zero():
    lens = np.array([100, 0, 8191, 0, 1], np.uint32)
    off = packed_offsets(lens, 8192)
    assert off.tolist() == [0, 100, 8292, 16483, 24675, 24676]


def _worker(rank, world, port, n, chunk, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n, rank, world)
    data = O.synth(chunk, lo, hi - lo)  # every rank generates its own shard of the stream, as bench.py does on device
    dig = torch.from_numpy(O.hash_chunks(O.HASH_SKEIN512, data, chunk))
    _, lens = O.compress_chunks(O.COMP_LZ4, data, chunk)
    d_all, l_all = gather_results(dig, torch.from_numpy(lens.astype(np.int32)), n, 64)
    if rank == 0:
        q.put((d_all.numpy().copy(), l_all.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process():
    n, chunk, world = 37, 4096, 2  # odd count: ranks get 18 and 19 chunks
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, chunk, q)) for r in range(world)]
    for p in procs:
        p.start()
    d_all, l_all = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    data = O.synth(chunk, 0, n)
    assert (d_all == O.hash_chunks(O.HASH_SKEIN512, data, chunk)).all()
    _, lens = O.compress_chunks(O.COMP_LZ4, data, chunk)
    assert (l_all.astype(np.uint32) == lens).all()
