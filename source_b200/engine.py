"""Device context + scene accelerator: the Python face of the C ABI.

``Device``       one ``rsb_context`` (one per GPU / process).
``Accelerator``  one uploaded scene; mirrors ``raysect.core.acceleration.Accelerator``
                 (build / hit / contains, accelerator.pxd:37-41) and adds the batched forms the GPU wants.
"""
import ctypes as C
import math

import numpy as np

from . import _cabi as cabi
from .flatten import flatten_world

_default = None


def default_device():
    global _default
    if _default is None:
        _default = Device(0)
    return _default


class Device:
    def __init__(self, index=0):
        self.lib = cabi.load()
        h = C.c_uint64()
        cabi.check(self.lib.rsb_context_create(int(index), C.byref(h)))
        self.ctx = h.value
        self.index = int(index)
        sm, major, minor, mem = C.c_int32(), C.c_int32(), C.c_int32(), C.c_uint64()
        cabi.check(self.lib.rsb_device_info(self.ctx, C.byref(sm), C.byref(major), C.byref(minor), C.byref(mem)))
        self.sm_count, self.cc, self.total_mem = sm.value, (major.value, minor.value), mem.value

    def close(self):
        if self.ctx:
            self.lib.rsb_context_destroy(self.ctx)
            self.ctx = 0

    def set_query_reorder(self, on=True):
        """sort every hit_batch / sweep pass on its coherence key before the traversal (rsb_set_query_reorder; on by
        default): the answers are the same, incoherent batches run 1.2-1.75x faster, coherent ones pay the sort for nothing"""
        cabi.check(self.lib.rsb_set_query_reorder(self.ctx, int(bool(on))))

    def build(self, world, world_kdtree=None):
        """Accelerator.build(world.primitives)"""
        return Accelerator(self, flatten_world(world, world_kdtree))

    def rng_uniform(self, seed, n):
        """raysect.core.math.random: seed(seed); [uniform() for _ in range(n)] -- on the device"""
        out = np.zeros(int(n), dtype=np.float64)
        cabi.check(self.lib.rsb_rng_uniform(self.ctx, int(seed), int(n), cabi.ptr(out, C.c_double)))
        return out

    def counters(self):
        c = cabi.RsbCounters()
        cabi.check(self.lib.rsb_counters(self.ctx, C.byref(c)))
        return c.as_dict()

    def render_stats(self):
        r = cabi.RsbRenderStats()
        cabi.check(self.lib.rsb_render_stats(self.ctx, C.byref(r)))
        return r.as_dict()

    def last_kernel_ms(self):
        ms = C.c_float()
        cabi.check(self.lib.rsb_last_kernel_ms(self.ctx, C.byref(ms)))
        return ms.value


class HitBatch:
    """Arrays returned by ``hit_batch`` (one row per ray)."""
    __slots__ = ("primitive", "distance", "sub", "exiting", "node", "geometry", "uvw")

    def __init__(self, n, geometry):
        self.primitive = np.zeros(n, dtype=np.int32)
        self.distance = np.zeros(n, dtype=np.float64)
        self.sub = np.zeros(n, dtype=np.int32)
        self.exiting = np.zeros(n, dtype=np.uint8)
        self.node = np.zeros((n, 2), dtype=np.int32)
        self.geometry = np.zeros((n, 12), dtype=np.float64) if geometry else None
        self.uvw = np.zeros((n, 3), dtype=np.float32) if geometry else None


class Accelerator:
    def __init__(self, device, flat):
        self.device = device
        self.lib = device.lib
        self.flat = flat
        h = C.c_uint64()
        cabi.check(self.lib.rsb_scene_create(device.ctx, C.byref(flat.desc), C.byref(h)))
        self.scene = h.value

    def close(self):
        if self.scene and self.device.ctx:
            self.lib.rsb_scene_destroy(self.device.ctx, self.scene)
        self.scene = 0

    # ---- World.hit ------------------------------------------------------------------------------------
    def hit_batch(self, origins, directions, max_distance=None, geometry=False):
        o = cabi.as_f64(origins).reshape(-1, 3)
        d = cabi.as_f64(directions).reshape(-1, 3)
        if o.shape != d.shape:
            raise ValueError("origins and directions must have the same shape")
        n = o.shape[0]
        md = None if max_distance is None else cabi.as_f64(np.broadcast_to(max_distance, (n,)))
        out = HitBatch(n, geometry)
        cabi.check(self.lib.rsb_hit_batch(
            self.device.ctx, self.scene, n, cabi.ptr(o, C.c_double), cabi.ptr(d, C.c_double), cabi.ptr(md, C.c_double),
            cabi.ptr(out.primitive, C.c_int32), cabi.ptr(out.distance, C.c_double), cabi.ptr(out.sub, C.c_int32),
            cabi.ptr(out.exiting, C.c_uint8), cabi.ptr(out.node, C.c_int32), cabi.ptr(out.geometry, C.c_double),
            cabi.ptr(out.uvw, C.c_float)))
        return out

    def hit(self, ray, point_cls=None, vector_cls=None, intersection_cls=None):
        """Accelerator.hit(ray) -> Intersection or None (1-element batch; API parity, not the fast path)"""
        from .math3d import Point3D, Vector3D
        from .scenegraph import Intersection
        P = point_cls or Point3D
        V = vector_cls or Vector3D
        o, d = ray.origin, ray.direction
        r = self.hit_batch([[o.x, o.y, o.z]], [[d.x, d.y, d.z]], [ray.max_distance], geometry=True)
        if r.primitive[0] < 0:
            return None
        prim = self.flat.primitives[int(r.primitive[0])]
        g = r.geometry[0]
        make = intersection_cls or Intersection
        it = make(ray, float(r.distance[0]), prim, P(g[0], g[1], g[2]), P(g[3], g[4], g[5]), P(g[6], g[7], g[8]),
                  V(g[9], g[10], g[11]), bool(r.exiting[0]), prim.to_local(), prim.to_root())
        if self.flat.prim_type[int(r.primitive[0])] == cabi.PRIM_MESH:
            it.triangle = int(r.sub[0])
            it.u, it.v, it.w = (float(x) for x in r.uvw[0])
        return it

    # ---- World.contains -------------------------------------------------------------------------------
    def contains_batch(self, points, cap=8):
        p = cabi.as_f64(points).reshape(-1, 3)
        n = p.shape[0]
        count = np.zeros(n, dtype=np.int32)
        prims = np.full((n, cap), -1, dtype=np.int32)
        cabi.check(self.lib.rsb_contains_batch(self.device.ctx, self.scene, n, cabi.ptr(p, C.c_double), int(cap),
                                               cabi.ptr(count, C.c_int32), cabi.ptr(prims, C.c_int32)))
        return count, prims

    def contains(self, point):
        cap = 8
        while True:
            count, prims = self.contains_batch([[point.x, point.y, point.z]], cap)
            if count[0] <= cap:
                return [self.flat.primitives[int(i)] for i in prims[0, :count[0]]]
            cap = int(count[0])

    # ---- Observer._render_pixel over a pixel list ---------------------------------------------------------
    def render_device(self, camera, config, spectral, rng_mode, seed, pixels=None, mean=None, variance=None,
                      count=False, stream=None, time_trace=False, passes=1, seed_stride=0):
        """Device-resident form: ``mean``/``variance`` are torch CUDA float64 tensors of shape (nx, ny, bins)
        (allocated zero-filled when None), ``pixels`` an int32 CUDA tensor [n, 2] or None for the whole frame.
        Enqueues on torch's current stream and returns (mean, variance, ray_count_tensor) without synchronising
        (unless ``count``, which reads the traversal counters back).  ``passes`` > 1 renders that many accumulated
        observe() calls of ``camera.pixel_samples`` samples each concurrently (rsb_render_passes_dev): pass p
        draws from the streams seeded ``seed + p*seed_stride + y*nx + x``."""
        import torch
        dev = torch.device("cuda", self.device.index)
        nx, ny, bins = camera.nx, camera.ny, config.bins
        if mean is None:
            mean = torch.zeros((nx, ny, bins), dtype=torch.float64, device=dev)
        if variance is None:
            variance = torch.zeros((nx, ny, bins), dtype=torch.float64, device=dev)
        rays = torch.zeros(1, dtype=torch.int64, device=dev)
        n = nx * ny if pixels is None else int(pixels.shape[0])
        rng = cabi.RsbRngDesc(mode=int(rng_mode), seed=int(seed))
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        cabi.check(self.lib.rsb_render_passes_dev(
            self.device.ctx, self.scene, C.c_void_p(st.cuda_stream), C.byref(camera), C.byref(config), C.byref(spectral),
            C.byref(rng), int(passes), int(seed_stride), n, C.c_void_p(0 if pixels is None else pixels.data_ptr()), C.c_void_p(mean.data_ptr()),
            C.c_void_p(variance.data_ptr()), C.c_void_p(rays.data_ptr()),
            (cabi.RENDER_COUNT if count else 0) | (cabi.RENDER_TIME_TRACE if time_trace else 0)))
        return mean, variance, rays

    def render(self, camera, config, spectral, rng_mode, seed, pixels=None, mean=None, variance=None, passes=1,
               seed_stride=0):
        """Host-buffer form (the call the observer makes): camera: RsbCamera, config: RsbRayConfig, spectral:
        RsbSpectral (from FlatScene.spectral()).  Returns (mean, variance, ray_count); mean/variance are
        (nx, ny, bins) float64 numpy arrays; only the listed pixels are written.  ``passes``: see render_device."""
        nx, ny, bins = camera.nx, camera.ny, config.bins
        if mean is None:
            mean = np.zeros((nx, ny, bins), dtype=np.float64)
        if variance is None:
            variance = np.zeros((nx, ny, bins), dtype=np.float64)
        rng = cabi.RsbRngDesc(mode=int(rng_mode), seed=int(seed))
        rays = C.c_uint64(0)
        pix = None
        n = nx * ny
        if pixels is not None:
            pix = cabi.as_i32(pixels).reshape(-1, 2)
            n = pix.shape[0]
        cabi.check(self.lib.rsb_render_passes(self.device.ctx, self.scene, C.byref(camera), C.byref(config),
                                              C.byref(spectral), C.byref(rng), int(passes), int(seed_stride), n,
                                              cabi.ptr(pix, C.c_int32), cabi.ptr(mean, C.c_double),
                                              cabi.ptr(variance, C.c_double), C.byref(rays)))
        return mean, variance, rays.value


    # ---- RenderEngine.run + Pipeline.update for a whole slice (the drop-in engine's form) --------------------------
    def render_slice(self, camera, config, spectral, rng_mode, seed, pixels=None, passes=1, seed_stride=0):
        """Renders like ``render`` but keeps the slice on the device (rsb_render_slice); returns the ray count.  Follow
        with ``update_frame`` (once per pipeline) and / or ``read_slice``."""
        rng = cabi.RsbRngDesc(mode=int(rng_mode), seed=int(seed))
        rays = C.c_uint64(0)
        pix, n = None, camera.nx * camera.ny
        if pixels is not None:
            pix = cabi.as_i32(pixels).reshape(-1, 2)
            n = pix.shape[0]
        cabi.check(self.lib.rsb_render_slice(self.device.ctx, self.scene, C.byref(camera), C.byref(config), C.byref(spectral),
                                             C.byref(rng), int(passes), int(seed_stride), n, cabi.ptr(pix, C.c_int32), C.byref(rays)))
        self._slice_shape = (camera.nx, camera.ny, config.bins)
        return rays.value

    def render_slices(self, camera, config, spectrals, rng_mode, seed, pixels=None, passes=1, seed_stride=None, xyz=None,
                      keep_spectral=True):
        """Every spectral slice of an observe() in one call (rsb_render_slices): ``spectrals`` = one RsbSpectral per
        slice (all with ``config.bins`` bins); slice k of pass p draws from the streams seeded
        ``seed + (p*len(spectrals) + k)*seed_stride + y*nx + x`` (``seed_stride`` defaults to nx*ny: the per-slice seeds
        of the engine / mirror camera).  The frame with ``len(spectrals)*config.bins`` bins per pixel stays on the device
        as the held slice; returns the ray count.

        ``xyz`` = (curves, delta_wavelength[, modes]) -- ``(n_slices, bins, n_channels)`` curves resampled on every slice's
        range, every slice's Spectrum.delta_wavelength and one ``cabi.PROJ_*`` mode per channel (default: three PROJ_XYZ
        channels = colour.resample_ciexyz for an RGBPipeline2D) -- also keeps what the pipelines' pixel processors would
        (rsb_render_slices_proj), for ``update_xyz_frame`` / ``update_proj_frame``; ``keep_spectral=False`` then drops the
        per-bin frame."""
        rng = cabi.RsbRngDesc(mode=int(rng_mode), seed=int(seed))
        rays = C.c_uint64(0)
        pix, n = None, camera.nx * camera.ny
        if pixels is not None:
            pix = cabi.as_i32(pixels).reshape(-1, 2)
            n = pix.shape[0]
        arr = (cabi.RsbSpectral * len(spectrals))()
        for k, sp in enumerate(spectrals):
            C.memmove(C.byref(arr[k]), C.byref(sp), C.sizeof(cabi.RsbSpectral))
        stride = camera.nx * camera.ny if seed_stride is None else int(seed_stride)
        if xyz is None:
            if not keep_spectral:
                raise ValueError("nothing to render: no projection curves and no spectral frame")
            cabi.check(self.lib.rsb_render_slices(self.device.ctx, self.scene, C.byref(camera), C.byref(config), arr, C.byref(rng),
                                                  int(passes), len(spectrals), stride, n, cabi.ptr(pix, C.c_int32), C.byref(rays)))
        else:
            curves = np.ascontiguousarray(xyz[0], dtype=np.float64)
            delta = np.ascontiguousarray(xyz[1], dtype=np.float64).reshape(-1)
            if curves.ndim != 3 or curves.shape[:2] != (len(spectrals), config.bins) or delta.shape != (len(spectrals),):
                raise ValueError("xyz must be ((n_slices, bins, n_channels) curves, (n_slices,) delta_wavelength[, modes])")
            modes = np.ascontiguousarray(xyz[2] if len(xyz) > 2 else [cabi.PROJ_XYZ] * curves.shape[2], dtype=np.int32)
            if modes.shape != (curves.shape[2],):
                raise ValueError("one projection mode per channel")
            cabi.check(self.lib.rsb_render_slices_proj(self.device.ctx, self.scene, C.byref(camera), C.byref(config), arr, C.byref(rng),
                                                       int(passes), len(spectrals), stride, n, cabi.ptr(pix, C.c_int32),
                                                       int(curves.shape[2]), cabi.ptr(modes, C.c_int32), cabi.ptr(curves, C.c_double),
                                                       cabi.ptr(delta, C.c_double), int(bool(keep_spectral)), C.byref(rays)))
        self._keep = list(spectrals)      # the tables the descriptors point at must outlive the call
        self._slice_shape = (camera.nx, camera.ny, config.bins * len(spectrals))
        return rays.value

    def update_proj_frame(self, channel0, frame_mean, frame_variance, frame_samples, frame_is_empty=False):
        """Pipeline.update + finalise for every listed pixel of the render done last with ``xyz=``: merges projection channels
        ``channel0 ..`` into the HOST frame arrays -- (nx, ny, k) or, for one channel, (nx, ny) StatsArray buffers, modified in
        place -- on the device (rsb_slice_update_proj_frame)."""
        nx, ny = self._slice_shape[:2]
        k = 1 if frame_mean.ndim == 2 else frame_mean.shape[2]
        for a, dt in ((frame_mean, np.float64), (frame_variance, np.float64), (frame_samples, np.int32)):
            if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags.c_contiguous and a.flags.writeable
                    and a.shape in ((nx, ny, k), (nx, ny)) and a.size == nx * ny * k):
                raise TypeError("frame arrays must be writable C-contiguous (nx, ny[, channels]) float64 / int32 numpy arrays")
        cabi.check(self.lib.rsb_slice_update_proj_frame(self.device.ctx, int(channel0), int(k), int(bool(frame_is_empty)),
                                                        cabi.ptr(frame_mean, C.c_double), cabi.ptr(frame_variance, C.c_double),
                                                        cabi.ptr(frame_samples, C.c_int32)))

    def update_bayer_frame(self, channel0, frame_mean, frame_variance, frame_samples, frame_is_empty=False):
        """BayerPipeline2D: the (nx, ny) frame takes, per pixel, the one of the three filter channels ``channel0 ..
        channel0 + 2`` its mosaic position selects (rsb_slice_update_bayer_frame)."""
        nx, ny = self._slice_shape[:2]
        for a, dt in ((frame_mean, np.float64), (frame_variance, np.float64), (frame_samples, np.int32)):
            if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags.c_contiguous and a.flags.writeable and a.shape == (nx, ny)):
                raise TypeError("frame arrays must be writable C-contiguous (nx, ny) float64 / int32 numpy arrays")
        cabi.check(self.lib.rsb_slice_update_bayer_frame(self.device.ctx, int(channel0), int(bool(frame_is_empty)),
                                                         cabi.ptr(frame_mean, C.c_double), cabi.ptr(frame_variance, C.c_double),
                                                         cabi.ptr(frame_samples, C.c_int32)))

    def update_xyz_frame(self, xyz_mean, xyz_variance, xyz_samples, frame_is_empty=False):
        """RGBPipeline2D.update + finalise for every listed pixel of the render done last with ``xyz=``: merges its XYZ
        statistics into the HOST ``xyz_frame`` arrays ((nx, ny, 3) StatsArray3D buffers, modified in place) on the device
        (rsb_slice_update_xyz_frame)."""
        shape = self._slice_shape[:2] + (3,)
        for a, dt in ((xyz_mean, np.float64), (xyz_variance, np.float64), (xyz_samples, np.int32)):
            if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags.c_contiguous and a.flags.writeable and a.shape == shape):
                raise TypeError("xyz frame arrays must be writable C-contiguous (nx, ny, 3) float64 / int32 numpy arrays")
        cabi.check(self.lib.rsb_slice_update_xyz_frame(self.device.ctx, int(bool(frame_is_empty)), cabi.ptr(xyz_mean, C.c_double),
                                                       cabi.ptr(xyz_variance, C.c_double), cabi.ptr(xyz_samples, C.c_int32)))

    def gather_from(self, others):
        """make this accelerator's held slice the whole multi-device result: the rows of every other member's listed
        pixels are loaded out of its device's memory by a kernel on this device (rsb_comm_gather_slices)"""
        if not others:
            return
        ctxs = (C.c_uint64 * (1 + len(others)))(self.device.ctx, *[o.device.ctx for o in others])
        comm = C.c_uint64()
        cabi.check(self.lib.rsb_comm_create(len(ctxs), ctxs, C.byref(comm)))
        try:
            cabi.check(self.lib.rsb_comm_gather_slices(comm.value, 0))
        finally:
            self.lib.rsb_comm_destroy(comm.value)

    def pin(self, *arrays):
        """page-lock numpy buffers (rsb_host_pin); returns a callable that releases them"""
        ptrs = [(a.ctypes.data, a.nbytes) for a in arrays if a.nbytes]
        for ptr, n in ptrs:
            cabi.check(self.lib.rsb_host_pin(self.device.ctx, C.c_void_p(ptr), n))

        def release():
            for ptr, _ in ptrs:
                self.lib.rsb_host_unpin(self.device.ctx, C.c_void_p(ptr))
        return release

    def read_slice(self):
        """(mean, variance) of the slice rendered last, (nx, ny, slice_bins); unlisted pixels are zero"""
        mean = np.zeros(self._slice_shape, dtype=np.float64)
        variance = np.zeros(self._slice_shape, dtype=np.float64)
        cabi.check(self.lib.rsb_slice_read(self.device.ctx, cabi.ptr(mean, C.c_double), cabi.ptr(variance, C.c_double)))
        return mean, variance

    def update_frame(self, frame_mean, frame_variance, frame_samples, slice_offset, frame_is_empty=False):
        """SpectralPowerPipeline2D.update for every listed pixel of the slice rendered last: merges it into the HOST
        frame arrays (StatsArray3D.mean / .variance / .samples, modified in place) with combine_samples, on the device
        (rsb_slice_update_frame).  ``frame_is_empty``: the frame holds no samples yet (skips its upload)."""
        for a, dt in ((frame_mean, np.float64), (frame_variance, np.float64), (frame_samples, np.int32)):
            if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags.c_contiguous and a.flags.writeable and a.ndim == 3):
                raise TypeError("frame arrays must be writable C-contiguous (nx, ny, bins) float64 / int32 numpy arrays")
        cabi.check(self.lib.rsb_slice_update_frame(self.device.ctx, int(frame_mean.shape[2]), int(slice_offset), int(bool(frame_is_empty)),
                                                   cabi.ptr(frame_mean, C.c_double), cabi.ptr(frame_variance, C.c_double),
                                                   cabi.ptr(frame_samples, C.c_int32)))


class DeviceGroup:
    """Several GPUs driven from ONE process, behind the surface of a single ``Accelerator`` (render_slice(s) /
    update_frame / update_xyz_frame): the pixel tasks of a render are dealt to the members in 16 x 16 tiles
    (distributed.tile_pixels' diagonal dealing for a whole frame; sorted tile ids round-robin for a task list), every
    member renders its list on its own device from its own host thread (ctypes releases the GIL), and the first member
    gathers the others' rows over peer memory (rsb_comm_gather_slices) so that ONE update_frame merges the whole slice
    into the pipeline's frame.  Pixel streams are keyed on the pixel: the result does not depend on the number of
    members.  ``members`` need render_slices / update_frame / update_xyz_frame and the first one ``gather_from(others)``
    (the device Accelerator; the host build of the tests emulates it)."""

    tile = 16

    def __init__(self, members):
        if not members:
            raise ValueError("a device group needs at least one member")
        self.members = list(members)
        self.flat = self.members[0].flat
        self._rows_gathered = False

    def close(self):
        for m in self.members:
            m.close()

    def _deal(self, nx, ny, pixels):
        from .distributed import tile_pixels
        n = len(self.members)
        if pixels is None:
            return [tile_pixels(nx, ny, self.tile, k, n) for k in range(n)]
        pix = cabi.as_i32(pixels).reshape(-1, 2)
        tile_id = (pix[:, 0] // self.tile).astype(np.int64) * 65536 + (pix[:, 1] // self.tile)
        rank = np.searchsorted(np.unique(tile_id), tile_id) % n
        return [np.ascontiguousarray(pix[rank == k]) for k in range(n)]

    def render_slices(self, camera, config, spectrals, rng_mode, seed, pixels=None, passes=1, seed_stride=None, xyz=None,
                      keep_spectral=True):
        from concurrent.futures import ThreadPoolExecutor
        lists = self._deal(camera.nx, camera.ny, pixels)
        stride = camera.nx * camera.ny if seed_stride is None else int(seed_stride)
        kw = dict(xyz=xyz, keep_spectral=keep_spectral) if xyz is not None else {}

        def one(job):
            m, p = job
            return m.render_slices(camera, config, spectrals, rng_mode, seed, p, passes=passes, seed_stride=stride, **kw)
        with ThreadPoolExecutor(max_workers=len(self.members)) as pool:
            rays = sum(pool.map(one, zip(self.members, lists)))
        self._rows_gathered = False
        return rays

    def render_slice(self, camera, config, spectral, rng_mode, seed, pixels=None, passes=1, seed_stride=0):
        return self.render_slices(camera, config, [spectral], rng_mode, seed, pixels, passes=passes, seed_stride=seed_stride)

    def update_xyz_frame(self, xyz_mean, xyz_variance, xyz_samples, frame_is_empty=False):
        # XYZ statistics are per work item and stay with their owners: every member merges its own pixels into the
        # (nx, ny, 3) frame, one after the other (before the spectral rows are gathered)
        if self._rows_gathered:
            raise RuntimeError("update_xyz_frame must precede update_frame after a multi-device render")
        for k, m in enumerate(self.members):
            m.update_xyz_frame(xyz_mean, xyz_variance, xyz_samples, frame_is_empty=frame_is_empty and k == 0)

    def update_proj_frame(self, channel0, frame_mean, frame_variance, frame_samples, frame_is_empty=False):
        if self._rows_gathered:
            raise RuntimeError("update_proj_frame must precede update_frame after a multi-device render")
        for k, m in enumerate(self.members):
            m.update_proj_frame(channel0, frame_mean, frame_variance, frame_samples, frame_is_empty=frame_is_empty and k == 0)

    def update_bayer_frame(self, channel0, frame_mean, frame_variance, frame_samples, frame_is_empty=False):
        if self._rows_gathered:
            raise RuntimeError("update_bayer_frame must precede update_frame after a multi-device render")
        for k, m in enumerate(self.members):
            m.update_bayer_frame(channel0, frame_mean, frame_variance, frame_samples, frame_is_empty=frame_is_empty and k == 0)

    def update_frame(self, frame_mean, frame_variance, frame_samples, slice_offset, frame_is_empty=False):
        if not self._rows_gathered:
            self.members[0].gather_from(self.members[1:])
            self._rows_gathered = True
        self.members[0].update_frame(frame_mean, frame_variance, frame_samples, slice_offset, frame_is_empty=frame_is_empty)

    def pin(self, *arrays):
        return self.members[0].pin(*arrays)


def camera_desc(nx, ny, pixel_samples, fov, sensitivity, to_root, width=None, ccd=False, vector=None, pixel=None):
    """PinholeCamera._update_image_geometry (raysect/optical/observer/imaging/pinhole.pyx:148-160), or, with
    ``width`` (and ``fov`` None), OrthographicCamera._update_image_geometry (imaging/orthographic.pyx:132-137) /
    with ``ccd`` CCDArray._update_image_geometry (imaging/ccd.pyx:106-112; ``sensitivity`` = pixel area x 2 pi,
    ccd.pyx:150-151)"""
    cam = cabi.RsbCamera()
    if pixel is not None:
        # Pixel, a 0-D observer (nonimaging/pixel.pyx): ``pixel`` = (x_width, y_width); (nx, ny) = (number of tasks, 1)
        x_width, y_width = float(pixel[0]), float(pixel[1])
        if x_width <= 0 or y_width <= 0:
            raise RuntimeError("Pixel widths must be greater than zero.")
        cam.kind = cabi.CAMERA_PIXEL
        cam.nx, cam.ny, cam.pixel_samples = int(nx), int(ny), int(pixel_samples)
        cam.image_delta, cam.image_start_x, cam.image_start_y = x_width, y_width, 0.0
        cam.sensitivity = float(sensitivity)
        for k, v in enumerate(float(to_root[i, j]) for i in range(3) for j in range(4)):
            cam.to_root[k] = v
        if [float(to_root[3, j]) for j in range(3)] != [0.0, 0.0, 0.0]:
            raise NotImplementedError("the observer's transform is not affine")
        cam.to_root_w = float(to_root[3, 3])
        return cam
    if vector is not None:
        # VectorCamera (imaging/vector.pyx:44-156): ``vector`` = (pixel_origins, pixel_directions), (nx, ny, 3) float64 each
        origins = np.ascontiguousarray(vector[0], dtype=np.float64)
        directions = np.ascontiguousarray(vector[1], dtype=np.float64)
        if origins.shape != (nx, ny, 3) or directions.shape != (nx, ny, 3):
            raise ValueError("Pixel arrays must have equal shapes.")
        cam.kind = cabi.CAMERA_VECTOR
        cam.pixel_origins = cabi.ptr(origins, C.c_double)
        cam.pixel_directions = cabi.ptr(directions, C.c_double)
        cam._keep = (origins, directions)
        image_delta = 0.0
    elif width is not None:
        if width <= 0:
            raise ValueError("width can not be less than or equal to 0 meters.")
        image_delta = width / nx
        cam.kind = cabi.CAMERA_CCD if ccd else cabi.CAMERA_ORTHOGRAPHIC
    else:
        max_pixels = max(nx, ny)
        if max_pixels <= 1:
            raise RuntimeError("Number of Pinhole camera Pixels must be > 1.")
        image_max_width = 2 * math.tan(math.pi / 180 * 0.5 * fov)
        image_delta = image_max_width / max_pixels
        cam.kind = cabi.CAMERA_PINHOLE
    cam.nx, cam.ny, cam.pixel_samples = int(nx), int(ny), int(pixel_samples)
    cam.image_delta = image_delta
    cam.image_start_x = 0.5 * nx * image_delta
    cam.image_start_y = 0.5 * ny * image_delta
    cam.sensitivity = float(sensitivity)
    for k, v in enumerate(float(to_root[i, j]) for i in range(3) for j in range(4)):
        cam.to_root[k] = v
    if [float(to_root[3, j]) for j in range(3)] != [0.0, 0.0, 0.0]:
        raise NotImplementedError("the observer's transform is not affine (bottom row %r)" % ([float(to_root[3, j]) for j in range(4)],))
    cam.to_root_w = float(to_root[3, 3])
    return cam


def ray_config(bins, min_wavelength, max_wavelength, extinction_prob=0.1, extinction_min_depth=3, max_depth=100,
               importance_sampling=True, important_path_weight=0.25, max_distance=math.inf):
    """optical Ray defaults, raysect/optical/ray.pyx:85-96"""
    if bins < 1:
        raise ValueError("Number of bins cannot be less than 1.")
    if min_wavelength <= 0.0 or max_wavelength <= 0.0:
        raise ValueError("Wavelength must be greater than to zero.")
    if min_wavelength >= max_wavelength:
        raise ValueError("Minimum wavelength must be less than the maximum wavelength.")
    if important_path_weight < 0 or important_path_weight > 1.0:
        raise ValueError("Important path weight must be in the range [0, 1].")
    # optical Ray property setters, ray.pyx:274-292
    if extinction_min_depth < 1:
        raise ValueError("The minimum extinction depth cannot be less than 1.")
    if max_depth < extinction_min_depth:
        raise ValueError("The maximum depth cannot be less than the minimum depth.")
    c = cabi.RsbRayConfig()
    c.bins = int(bins)
    c.extinction_min_depth = int(extinction_min_depth)
    c.max_depth = int(max_depth)
    c.importance_sampling = int(bool(importance_sampling))
    c.min_wavelength, c.max_wavelength = float(min_wavelength), float(max_wavelength)
    c.extinction_prob = min(max(float(extinction_prob), 0.0), 1.0)    # clamp(extinction_prob, 0, 1), ray.pyx:262
    c.important_path_weight = float(important_path_weight)
    c.max_distance = float(max_distance)
    return c
