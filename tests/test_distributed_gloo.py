"""N > 1 host logic on CPU: tile partition + the single reduce(sum) of the frame, world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from source_b200.distributed import reduce_frame, tile_pixels


def test_tile_partition_is_a_disjoint_cover():
    for nx, ny, tile, ws in [(64, 48, 16, 2), (50, 37, 16, 3), (1024, 1024, 16, 8), (5, 7, 16, 4)]:
        seen = np.zeros((nx, ny), dtype=np.int32)
        sizes = []
        for r in range(ws):
            px = tile_pixels(nx, ny, tile, r, ws)
            assert px.dtype == np.int32 and px.shape[1] == 2
            seen[px[:, 0], px[:, 1]] += 1
            sizes.append(len(px))
        assert np.all(seen == 1)
        if nx * ny > 10000:
            assert max(sizes) - min(sizes) <= 2 * tile * tile * (nx // tile + 1)   # interleaved => balanced


def _worker(rank, world_size, port, nx, ny, bins, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        ix, iy, ib = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(bins), indexing="ij")
        full = np.stack([np.sin(ix * 0.37 + iy * 1.91 + ib * 0.11) * 1e3, np.cos(ix * 0.7 - iy * 0.3 + ib) ** 2])
        stats = torch.zeros((2, nx, ny, bins), dtype=torch.float64)
        px = tile_pixels(nx, ny, 16, rank, world_size)
        stats[:, px[:, 0], px[:, 1], :] = torch.from_numpy(full[:, px[:, 0], px[:, 1], :])
        out = torch.zeros_like(stats)
        for _ in range(2):      # twice: the rank's own buffer must stay zero outside its tiles between frames
            frame = reduce_frame(stats, out, rank, world_size)
        if rank == 0:
            q.put(bool(np.array_equal(frame.numpy(), full)))   # bit-exact: the sum only ever adds zeros
        else:
            assert frame is None
    finally:
        dist.destroy_process_group()


def test_reduce_assembles_frame_bit_exactly_world_size_2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 40, 33, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
