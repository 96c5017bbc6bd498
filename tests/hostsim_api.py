"""Loader for tests/_build/libhostsim.so: the product's device headers compiled as host C++ and driven
serially (tests/hostsim/hostsim.cpp).  TEST INFRASTRUCTURE ONLY -- lets the CPU test-suite pin the
arithmetic the CUDA kernels execute against the reference; never imported by source_b200."""
import ctypes as C
import os
import subprocess

import numpy as np

from source_b200 import _cabi as cabi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "_build", "libhostsim.so")
SOURCES = [os.path.join(HERE, "hostsim", "hostsim.cpp"),
           os.path.join(ROOT, "source_b200", "csrc", "scene_pack.cpp"),
           os.path.join(ROOT, "source_b200", "csrc", "kdtree_host.cpp")]
HEADERS = [os.path.join(ROOT, "source_b200", "csrc", h) for h in
           ("rsb_math.h", "rsb_scene.h", "rsb_geom.h", "rsb_trav.h", "rsb_rng.h", "rsb_path.h", "scene_pack.h", "kdtree_host.h")]


def build(force=False):
    newest = max(os.path.getmtime(p) for p in SOURCES + HEADERS)
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-pthread", "-fPIC", "-shared", "-o", SO] + SOURCES)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.hs_last_error.restype = C.c_char_p
        _lib.hs_scene_create.argtypes = [C.POINTER(cabi.RsbSceneDesc), C.POINTER(C.c_uint64)]
        _lib.hs_scene_destroy.argtypes = [C.c_uint64]
        _lib.hs_hit_batch.argtypes = [C.c_uint64, C.c_int64] + [C.c_void_p] * 11
        _lib.hs_hit_batch_mode.argtypes = [C.c_uint64, C.c_int32, C.c_int64] + [C.c_void_p] * 11
        _lib.hs_contains_batch.argtypes = [C.c_uint64, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        _lib.hs_rng_uniform.argtypes = [C.c_uint64, C.c_int64, C.c_void_p]
        _lib.hs_mt_state.argtypes = [C.c_uint64, C.c_int32, C.c_void_p]
        _lib.hs_render.argtypes = [C.c_uint64, C.POINTER(cabi.RsbCamera), C.POINTER(cabi.RsbRayConfig),
                                   C.POINTER(cabi.RsbSpectral), C.POINTER(cabi.RsbRngDesc), C.c_int64, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p, C.c_double,
                                   C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        _lib.hs_frame_combine.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                          C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


class HostScene:
    """Same surface as source_b200.engine.Accelerator, evaluated by the host build of the device code."""

    hit_mode = 0    # 0: nested loop, 1: split pipeline (serial), 2: single-visit walk (mesh-free scenes)

    def __init__(self, flat):
        self.flat = flat
        h = C.c_uint64()
        rc = lib().hs_scene_create(C.byref(flat.desc), C.byref(h))
        if rc:
            raise cabi.RsbError(rc, lib().hs_last_error().decode())
        self.scene = h.value
        self.counters = None

    def close(self):
        if self.scene:
            lib().hs_scene_destroy(self.scene)
            self.scene = 0

    def hit_batch(self, origins, directions, max_distance=None, geometry=True):
        from source_b200.engine import HitBatch
        o = cabi.as_f64(origins).reshape(-1, 3)
        d = cabi.as_f64(directions).reshape(-1, 3)
        n = o.shape[0]
        md = None if max_distance is None else cabi.as_f64(np.broadcast_to(max_distance, (n,)))
        out = HitBatch(n, True)
        counters = np.zeros(5, dtype=np.uint64)
        lib().hs_hit_batch_mode(self.scene, int(self.hit_mode), n, _p(o), _p(d), _p(md), _p(out.primitive), _p(out.distance),
                                _p(out.sub), _p(out.exiting), _p(out.node), _p(out.geometry), _p(out.uvw), _p(counters))
        self.counters = dict(zip(("branches", "leaves", "items", "prim_tests", "tri_tests"), (int(c) for c in counters)))
        return out

    def contains_batch(self, points, cap=8):
        p = cabi.as_f64(points).reshape(-1, 3)
        n = p.shape[0]
        count = np.zeros(n, dtype=np.int32)
        prims = np.full((n, cap), -1, dtype=np.int32)
        lib().hs_contains_batch(self.scene, n, _p(p), cap, _p(count), _p(prims))
        return count, prims

    def render(self, camera, config, spectral, rng_mode, seed, pixels=None, mean=None, variance=None, passes=1,
               seed_stride=0, xyz=None):
        """``xyz`` = ((bins, 3) curves, delta_wavelength): the per-task XYZ statistics of every pass are left in
        ``self._xyz_passes`` as [(mean (n, 3), variance (n, 3)), ...]"""
        if passes > 1:
            return self._render_passes(camera, config, spectral, rng_mode, seed, pixels, mean, variance, passes, seed_stride, xyz)
        nx, ny, bins = camera.nx, camera.ny, config.bins
        if mean is None:
            mean = np.zeros((nx, ny, bins))
        if variance is None:
            variance = np.zeros((nx, ny, bins))
        rng = cabi.RsbRngDesc(mode=int(rng_mode), seed=int(seed))
        rays = C.c_uint64(0)
        pix, n = None, nx * ny
        if pixels is not None:
            pix = cabi.as_i32(pixels).reshape(-1, 2)
            n = pix.shape[0]
        counters = np.zeros(7, dtype=np.uint64)
        curves = xm = xv = modes = None
        nch = 0
        if xyz is not None:
            curves = np.ascontiguousarray(xyz[0], dtype=np.float64)
            nch = curves.shape[1]
            assert curves.shape == (bins, nch)
            modes = np.ascontiguousarray(xyz[2] if len(xyz) > 2 else [cabi.PROJ_XYZ] * nch, dtype=np.int32)
            xm, xv = np.zeros((n, nch)), np.zeros((n, nch))
        rc = lib().hs_render(self.scene, C.byref(camera), C.byref(config), C.byref(spectral), C.byref(rng), n, _p(pix),
                             _p(mean), _p(variance), C.byref(rays), _p(counters), _p(curves), float(xyz[1]) if xyz is not None else 0.0,
                             _p(xm), _p(xv), nch, _p(modes))
        self._xyz_passes = [(xm, xv)] if xyz is not None else None
        if rc:
            raise cabi.RsbError(rc, lib().hs_last_error().decode())
        self.counters = dict(zip(("branches", "leaves", "items", "prim_tests", "tri_tests", "paths", "segments"),
                                 (int(c) for c in counters)))
        return mean, variance, rays.value


    # the drop-in engine's whole-slice form (Accelerator.render_slice / update_frame), restated with numpy
    def render_slice(self, camera, config, spectral, rng_mode, seed, pixels=None, passes=1, seed_stride=0):
        mean, variance, rays = self.render(camera, config, spectral, rng_mode, seed, pixels, passes=passes, seed_stride=seed_stride)
        pix = None if pixels is None else cabi.as_i32(pixels).reshape(-1, 2)
        self._slice = (mean, variance, pix, camera.pixel_samples * passes)
        return rays

    def render_slices(self, camera, config, spectrals, rng_mode, seed, pixels=None, passes=1, seed_stride=None, xyz=None,
                      keep_spectral=True):
        """rsb_render_slices(_xyz) restated with the sequential pieces: one render per slice with the slice's own seed base"""
        nx, ny, bins = camera.nx, camera.ny, config.bins
        stride = nx * ny if seed_stride is None else seed_stride
        n = len(spectrals)
        mean, variance, total = np.zeros((nx, ny, bins * n)), np.zeros((nx, ny, bins * n)), 0
        xyz_slices = []
        for k, sp in enumerate(spectrals):
            m, v, rays = self.render(camera, config, sp, rng_mode, seed + k * stride, pixels, passes=passes, seed_stride=n * stride,
                                     xyz=None if xyz is None else (np.asarray(xyz[0])[k], np.asarray(xyz[1]).reshape(-1)[k]) + tuple(xyz[2:]))
            xyz_slices.append(self._xyz_passes)
            mean[:, :, k * bins:(k + 1) * bins] = m
            variance[:, :, k * bins:(k + 1) * bins] = v
            total += rays
        pix = None if pixels is None else cabi.as_i32(pixels).reshape(-1, 2)
        self._slice = (mean, variance, pix, camera.pixel_samples * passes) if keep_spectral else None
        self._xyz = None if xyz is None else (xyz_slices, pix, camera.pixel_samples, (nx, ny))
        return total

    def update_xyz_frame(self, xyz_mean, xyz_variance, xyz_samples, frame_is_empty=False):
        self.update_proj_frame(0, xyz_mean, xyz_variance, xyz_samples, frame_is_empty)

    def update_bayer_frame(self, channel0, frame_mean, frame_variance, frame_samples, frame_is_empty=False):
        """rsb_slice_update_bayer_frame restated: the three filter channels merged into a scratch (nx, ny, 3) copy of the
        frame, then every pixel keeps the channel its mosaic position selects"""
        nx, ny = frame_mean.shape
        m3, v3, s3 = (np.repeat(a[:, :, None], 3, axis=2).copy() for a in (frame_mean, frame_variance, frame_samples))
        self.update_proj_frame(channel0, m3, v3, s3, frame_is_empty)
        xs, ys = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
        sel = np.array([0, 1, 1, 2])[(xs % 2) + 2 * (ys % 2)]
        for dst, src in ((frame_mean, m3), (frame_variance, v3), (frame_samples, s3)):
            dst[:, :] = np.take_along_axis(src, sel[:, :, None], axis=2)[:, :, 0]

    def update_proj_frame(self, channel0, xyz_mean, xyz_variance, xyz_samples, frame_is_empty=False):
        """rsb_slice_update_proj_frame restated with numpy: per pass, the slices' statistics summed in slice order, merged
        with combine_samples"""
        from source_b200.observer import combine_samples
        slices, pix, samples, (nx, ny) = self._xyz
        flat2d = xyz_mean.ndim == 2
        if flat2d:
            xyz_mean, xyz_variance, xyz_samples = (a.reshape(nx, ny, 1) for a in (xyz_mean, xyz_variance, xyz_samples))
        nc = xyz_mean.shape[2]
        if pix is None:
            xs, ys = (a.reshape(-1) for a in np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij"))
        else:
            xs, ys = pix[:, 0], pix[:, 1]
        if frame_is_empty:
            assert not xyz_samples.any()
        for p in range(len(slices[0])):
            wm, wv = np.zeros((len(xs), nc)), np.zeros((len(xs), nc))
            for parts in slices:
                wm = wm + parts[p][0][:, channel0:channel0 + nc]
                wv = wv + parts[p][1][:, channel0:channel0 + nc]
            mt, vt, nt = combine_samples(xyz_mean[xs, ys], xyz_variance[xs, ys], xyz_samples[xs, ys], wm, np.maximum(wv, 0.0), samples)
            xyz_mean[xs, ys] = mt
            xyz_variance[xs, ys] = vt
            xyz_samples[xs, ys] = nt

    def read_slice(self):
        return self._slice[0].copy(), self._slice[1].copy()

    def gather_from(self, others):
        """rsb_comm_gather_slices restated with numpy: the rows of every other member's listed pixels are copied in"""
        mean, variance, pix, samples = self._slice
        lists = [] if pix is None else [pix]
        assert pix is not None or not others or all(len(o._slice[2]) == 0 for o in others)
        for o in others:
            om, ov, op, osamples = o._slice
            assert op is not None and osamples == samples and om.shape == mean.shape
            mean[op[:, 0], op[:, 1]] = om[op[:, 0], op[:, 1]]
            variance[op[:, 0], op[:, 1]] = ov[op[:, 0], op[:, 1]]
            lists.append(op)
        union = np.concatenate(lists) if lists else None
        if union is not None:
            assert len(np.unique(union[:, 0].astype(np.int64) * 65536 + union[:, 1])) == len(union), "pixel lists overlap"
            if len(union) == mean.shape[0] * mean.shape[1]:
                union = None
        self._slice = (mean, variance, union, samples)

    def update_frame(self, frame_mean, frame_variance, frame_samples, slice_offset, frame_is_empty=False):
        from source_b200.observer import combine_samples
        mean, variance, pix, samples = self._slice
        sl = slice(slice_offset, slice_offset + mean.shape[2])
        if pix is None:
            xs, ys = (a.reshape(-1) for a in np.meshgrid(np.arange(mean.shape[0]), np.arange(mean.shape[1]), indexing="ij"))
        else:
            xs, ys = pix[:, 0], pix[:, 1]
        if frame_is_empty:
            assert not frame_samples[:, :, sl].any()
        mt, vt, nt = combine_samples(frame_mean[xs, ys, sl], frame_variance[xs, ys, sl], frame_samples[xs, ys, sl],
                                     mean[xs, ys], np.maximum(variance[xs, ys], 0.0), samples)
        frame_mean[xs, ys, sl] = mt
        frame_variance[xs, ys, sl] = vt
        frame_samples[xs, ys, sl] = nt

    def _render_passes(self, camera, config, spectral, rng_mode, seed, pixels, mean, variance, passes, seed_stride, xyz=None):
        """rsb_render_passes restated with the sequential pieces: one render per pass, merged in pass order into an
        empty frame with StatsArray3D.combine_samples (hs_frame_combine)."""
        nx, ny, bins = camera.nx, camera.ny, config.bins
        fm, fv = np.zeros((nx, ny, bins)), np.zeros((nx, ny, bins))
        fs = np.zeros((nx, ny, bins), dtype=np.int32)
        total = 0
        xyz_passes = []
        for p in range(passes):
            m, v, rays = self.render(camera, config, spectral, rng_mode, seed + p * seed_stride, pixels, xyz=xyz)
            xyz_passes.extend(self._xyz_passes or [])
            total += rays
            lib().hs_frame_combine(nx * ny, bins, 0, bins, _p(m), _p(v), camera.pixel_samples, _p(fm), _p(fv), _p(fs))
        if mean is None:
            mean = np.zeros((nx, ny, bins))
        if variance is None:
            variance = np.zeros((nx, ny, bins))
        listed = np.ones((nx, ny), dtype=bool)
        if pixels is not None:
            pix = cabi.as_i32(pixels).reshape(-1, 2)
            listed[:] = False
            listed[pix[:, 0], pix[:, 1]] = True
        mean[listed] = fm[listed]
        variance[listed] = fv[listed]
        self._xyz_passes = xyz_passes if xyz is not None else None
        return mean, variance, total


def rng_uniform(seed, n):
    out = np.zeros(n)
    lib().hs_rng_uniform(int(seed), n, _p(out))
    return out


def mt_state(seed, fast):
    out = np.zeros(312, dtype=np.uint64)
    lib().hs_mt_state(int(seed), int(fast), _p(out))
    return out
