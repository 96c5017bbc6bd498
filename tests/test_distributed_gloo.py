"""N > 1 host logic on CPU: tile partition + the single collective (gather of owned tiles), world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from source_b200.distributed import TileGather, tile_pixels


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
        stats = torch.full((2, nx, ny, bins), 123.0, dtype=torch.float64)      # foreign rows hold leftovers, as after a previous frame
        px = tile_pixels(nx, ny, 16, rank, world_size)
        gather = TileGather(nx, ny, bins, 16, rank, world_size, torch.device("cpu"))
        ok = True
        for k in range(2):      # twice: buffers are reused from frame to frame
            stats[:, px[:, 0], px[:, 1], :] = torch.from_numpy(full[:, px[:, 0], px[:, 1], :]) + k
            frame = gather(stats)
            if rank == 0:
                ok = ok and bool(np.array_equal(frame.numpy(), full + k))   # bit-exact: rows are copied
            else:
                assert frame is None
        if rank == 0:
            q.put(ok)
    finally:
        dist.destroy_process_group()


def test_gather_assembles_frame_bit_exactly_world_size_2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 40, 50, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def _render_worker(rank, world_size, port, q):
    """each rank renders ITS tiles of a Cornell frame (host build of the device code, 2 accumulated passes) into an
    otherwise-zero frame; rank 0 checks the gathered frame against the single-process render of the whole frame"""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [os.path.dirname(here), here]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        import hostsim_api
        import scenes
        import source_b200 as api
        from source_b200 import _cabi as cabi
        from source_b200.engine import camera_desc, ray_config
        from source_b200.flatten import flatten_world
        nx, ny, bins, spp, passes, seed = 24, 20, 4, 2, 2, 77
        flat = flatten_world(scenes.cornell_box(api))
        be = hostsim_api.HostScene(flat)
        cam = camera_desc(nx, ny, spp, 45.0, 1.0, api.translate(0, 0, -3.3))
        cfg = ray_config(bins, 375.0, 740.0, 0.01, 3, 500, True, 0.25)
        sp = flat.spectral(375.0, 740.0, bins)
        px = tile_pixels(nx, ny, 8, rank, world_size)
        m, v, rays = be.render(cam, cfg, sp, cabi.RNG_MT19937_64, seed, px, passes=passes, seed_stride=nx * ny)
        stats = torch.from_numpy(np.stack([m, v]))
        frame = TileGather(nx, ny, bins, 8, rank, world_size, torch.device("cpu"))(stats)
        total = torch.tensor([rays], dtype=torch.int64)
        dist.all_reduce(total)
        if rank == 0:
            m1, v1, rays1 = be.render(cam, cfg, sp, cabi.RNG_MT19937_64, seed, None, passes=passes, seed_stride=nx * ny)
            q.put(bool(np.array_equal(frame[0].numpy(), m1) and np.array_equal(frame[1].numpy(), v1)
                       and int(total.item()) == rays1 and m1.sum() > 0))
        be.close()
    finally:
        dist.destroy_process_group()


def test_tile_partitioned_render_equals_single_process_render_world_size_2(lib):
    """pixel streams are keyed on (pass, pixel): the frame must not depend on how tiles are dealt to ranks"""
    import hostsim_api
    hostsim_api.build()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_render_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok
