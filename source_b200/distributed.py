"""One process per GPU: image tiles are partitioned across ranks, each rank renders every sample of its
own tiles, and ONE NCCL gather over NVLink assembles the spectral frame on rank 0 (exact: rows are copied).

The reference's only parallelism is the same pixel-task data parallelism over forked processes with
pickled (mean, variance) tuples (raysect/core/workflow.py:123-327); pixel random streams are keyed on the
pixel, so the frame does not depend on the number of ranks.
"""
import ctypes as C
import os

import numpy as np

from . import _cabi as cabi
from .engine import camera_desc, ray_config


def tile_pixels(nx, ny, tile, rank, world_size):
    """Pixels (x, y) of the tiles owned by `rank`: tile (i, j) of the tile grid -> rank (i + j) % world_size, the
    diagonals of the grid.  (Row-major dealing, t % world_size, hands a rank whole stripes of the image whenever the
    tile-grid width is a multiple of world_size -- 64 tiles across, 8 ranks: every rank one 16-pixel stripe in eight --
    and the stripes of a Cornell box do not cost the same.)"""
    tx, ty = (nx + tile - 1) // tile, (ny + tile - 1) // tile
    t = np.arange(tx * ty)
    deal = t if os.environ.get("RSB_TILE_DEAL") == "rowmajor" else t // ty + t % ty
    mine = t[deal % world_size == rank]
    ox, oy = (mine // ty) * tile, (mine % ty) * tile
    dx, dy = np.meshgrid(np.arange(tile), np.arange(tile), indexing="ij")
    x = (ox[:, None, None] + dx[None]).reshape(-1)
    y = (oy[:, None, None] + dy[None]).reshape(-1)
    keep = (x < nx) & (y < ny)
    return np.ascontiguousarray(np.stack([x[keep], y[keep]], axis=1).astype(np.int32))


class TileGather:
    """The single collective of the path.  Every rank packs the frame rows of ITS tiles (1/N of the frame) and rank 0
    gathers them over NCCL / NVLink and drops them into place in its own frame -- copies only, so the assembled frame
    holds bit for bit what the owning ranks wrote.  (Round 1 summed whole frames with reduce(sum): exact as well,
    because foreign entries were zero, but every rank shipped the full 2 x 1.07 GB frame, 7/8 of it zeros, after a
    2.15 GB scratch copy.)  ``stats`` is (2, nx, ny, bins): mean, variance."""

    def __init__(self, nx, ny, bins, tile, rank, world_size, device):
        import torch
        self.torch = torch
        self.rank, self.world_size = rank, world_size
        self.shape = (nx, ny, bins)
        self.index, self.counts = [], []
        for r in range(world_size):
            px = tile_pixels(nx, ny, tile, r, world_size)
            self.index.append(torch.from_numpy(px[:, 0].astype(np.int64) * ny + px[:, 1].astype(np.int64)).to(device))
            self.counts.append(len(px))
        n_max = max(self.counts)
        self.send = torch.zeros((2, n_max, bins), dtype=torch.float64, device=device)
        self.recv = [torch.zeros_like(self.send) for _ in range(world_size)] if rank == 0 else None
        self.bytes_sent = 0 if rank == 0 else 2 * self.counts[rank] * bins * 8

    def __call__(self, stats):
        """-> the assembled frame on rank 0 (assembled in place in rank 0's ``stats``: a render only ever touches the
        rows of its own pixels, so foreign rows left over from the previous frame are harmless), None elsewhere"""
        if self.world_size == 1:
            return stats
        import torch.distributed as dist
        torch = self.torch
        nx, ny, bins = self.shape
        flat = stats.view(2, nx * ny, bins)
        n_own = self.counts[self.rank]
        if self.rank != 0:
            torch.index_select(flat, 1, self.index[self.rank], out=self.send[:, :n_own]) if n_own == self.send.shape[1] \
                else self.send[:, :n_own].copy_(flat.index_select(1, self.index[self.rank]))
        dist.gather(self.send, self.recv, dst=0)
        if self.rank != 0:
            return None
        for r in range(1, self.world_size):
            flat.index_copy_(1, self.index[r], self.recv[r][:, :self.counts[r]])
        return stats


class FrameRenderer:
    """Renders one spectral slice of a PinholeCamera frame on `world_size` GPUs (this process = `rank`)."""

    def __init__(self, camera, accel, rank=0, world_size=1, tile=16, backend_reduce=None, passes=1):
        import torch
        self.torch = torch
        self.camera, self.accel = camera, accel
        self.rank, self.world_size = rank, world_size
        nx, ny = camera.pixels
        self.nx, self.ny, self.bins = nx, ny, camera.spectral_bins
        if camera.spectral_rays != 1:
            raise NotImplementedError("FrameRenderer handles one spectral slice per call")
        # ``passes`` accumulated observe() calls of pixel_samples / passes samples each, rendered concurrently
        # (rsb_render_passes_dev): the frame holds camera.pixel_samples samples per pixel either way
        self.passes = int(passes)
        if self.passes < 1 or camera.pixel_samples % self.passes:
            raise ValueError("pixel_samples must be a multiple of the number of passes")
        self.dev = torch.device("cuda", accel.device.index)
        self.cam = camera._camera_desc(camera.pixel_samples // self.passes)
        self.cfg = ray_config(self.bins, camera.min_wavelength, camera.max_wavelength, camera.ray_extinction_prob,
                              camera.ray_extinction_min_depth, camera.ray_max_depth, camera.ray_importance_sampling,
                              camera.ray_important_path_weight)
        self.spectral = accel.flat.spectral(camera.min_wavelength, camera.max_wavelength, self.bins)
        self.host_pixels = None
        self.dev_pixels = None
        if world_size > 1:
            px = tile_pixels(nx, ny, tile, rank, world_size)
            self.host_pixels = torch.from_numpy(px).pin_memory()
            self.dev_pixels = self.host_pixels.to(self.dev)
        # [0] mean, [1] variance: one tensor so that the frame crosses NVLink in a single reduce
        self.stats = torch.zeros((2, nx, ny, self.bins), dtype=torch.float64, device=self.dev)
        self.gather = backend_reduce or TileGather(nx, ny, self.bins, tile, rank, world_size, self.dev)
        self.host_stats = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _render(self, seed, pixels, count=False, time_trace=False):
        return self.accel.render_device(self.cam, self.cfg, self.spectral, self.camera.rng_mode, seed, pixels,
                                        self.stats[0], self.stats[1], count=count, time_trace=time_trace,
                                        passes=self.passes, seed_stride=self.nx * self.ny)[2]

    def _assemble(self):
        """the single collective: owned tiles gathered on rank 0"""
        return self.gather(self.stats)

    def step_device(self, seed, time_trace=False):
        """One frame with everything resident in HBM.  Returns this rank's ray counter (device tensor).
        With ``time_trace`` every launch of the dominant kernel is bracketed with CUDA events
        (Device.render_stats()['trace_ms'])."""
        rays = self._render(seed, self.dev_pixels, time_trace=time_trace)
        self._assemble()
        return rays

    def step_host(self, seed):
        """One frame through host buffers: pixel list and tables host->device, frame device->host (pinned),
        installed as the pipeline's StatsArray3D.  Returns this rank's ray count (int)."""
        torch = self.torch
        h2d = self.accel.flat.materials.__len__() * self.bins * 8
        pixels = None
        if self.host_pixels is not None:
            pixels = self.host_pixels.to(self.dev, non_blocking=True)
            h2d += self.host_pixels.numel() * 4
        rays = self._render(seed, pixels)
        frame = self._assemble()
        d2h = 8
        if self.rank == 0:
            # two pinned staging buffers: the device->host copy always lands in the one the pipeline's frame does NOT
            # view, so an accumulating pipeline combines (frame so far, new pass) and never (new, new)
            if self.host_stats is None:
                self.host_stats = [torch.zeros(self.stats.shape, dtype=torch.float64).pin_memory() for _ in range(2)]
                self._frame_buf = None
            k = 0 if self._frame_buf != 0 else 1
            self.host_stats[k].copy_(frame, non_blocking=True)
            d2h += self.host_stats[k].numel() * 8
        n = int(rays.item())    # device->host read of the step's ray counter; synchronises the stream
        if self.rank == 0:
            torch.cuda.current_stream(self.dev).synchronize()
            pipe = self.camera.pipelines[0]
            hs = self.host_stats[k].numpy()
            if pipe.frame is None or pipe.frame.shape != (self.nx, self.ny, self.bins) or not pipe.accumulate:
                from .observer import StatsArray3D
                f = StatsArray3D.__new__(StatsArray3D)
                f.nx, f.ny, f.nz = self.nx, self.ny, self.bins
                f.mean, f.variance = hs[0], hs[1]
                f.samples = np.full((self.nx, self.ny, self.bins), self.camera.pixel_samples, dtype=np.int32) \
                    if getattr(self, "_samples", None) is None or pipe.accumulate else self._samples
                if not pipe.accumulate:
                    self._samples = f.samples
                pipe.frame = f
                self._frame_buf = k
            else:
                pipe._samples = self.camera.pixel_samples
                pipe._spectral_slices = self.camera._slice_spectrum()
                pipe.update_slice(None, 0, hs[0], hs[1])
        self.h2d_bytes, self.d2h_bytes = h2d, d2h
        return n

    def count_pass(self, seed):
        """Untimed pass with the traversal counters switched on (roofline model inputs)."""
        self._render(seed, self.dev_pixels, count=True)
        c = self.accel.device.counters()
        self.local_counters = dict(c)      # this rank's share (per-GPU roofline)
        if self.world_size > 1:
            import torch.distributed as dist
            keys = sorted(c)
            t = self.torch.tensor([c[k] for k in keys], dtype=self.torch.int64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            c = {k: int(v) for k, v in zip(keys, t.tolist())}
        return c
