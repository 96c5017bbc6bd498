"""ctypes loader for oracle/rs_oracle.c -- TEST INFRASTRUCTURE ONLY (see the header of that file).
Builds oracle/_build/librs_oracle.so with gcc (-O2 -ffp-contract=off: no FMA, like the reference build)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "rs_oracle.c")
SO = os.path.join(HERE, "_build", "librs_oracle.so")


def build(force=False):
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["gcc", "-std=gnu11", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", SO, SRC, "-lm"])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def uniform(seed, n):
    out = np.zeros(n)
    lib().ro_uniform(C.c_uint64(seed), C.c_int64(n), _p(out))
    return out


class OracleScene:
    """Same surface as source_b200.engine.Accelerator (hit_batch / contains_batch / render), evaluated by the
    independent C restatement.  `flat` is a source_b200.flatten.FlatScene (only its C-ABI descriptor is read)."""

    def __init__(self, flat):
        self.flat = flat

    def close(self):
        pass

    def hit_batch(self, origins, directions, max_distance=None, geometry=True):
        from source_b200.engine import HitBatch
        o = np.ascontiguousarray(origins, dtype=np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(directions, dtype=np.float64).reshape(-1, 3)
        n = o.shape[0]
        md = None if max_distance is None else np.ascontiguousarray(np.broadcast_to(max_distance, (n,)), dtype=np.float64)
        out = HitBatch(n, True)
        lib().ro_hit(C.byref(self.flat.desc), C.c_int64(n), _p(o), _p(d), _p(md), _p(out.primitive), _p(out.distance), _p(out.sub),
                     _p(out.exiting), _p(out.geometry), _p(out.uvw), _p(out.node))
        return out

    def contains_batch(self, points, cap=8):
        p = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        n = p.shape[0]
        count = np.zeros(n, dtype=np.int32)
        prims = np.full((n, cap), -1, dtype=np.int32)
        lib().ro_contains(C.byref(self.flat.desc), C.c_int64(n), _p(p), C.c_int32(cap), _p(count), _p(prims))
        return count, prims

    def render(self, camera, config, spectral, rng_mode, seed, pixels=None, mean=None, variance=None, passes=1,
               seed_stride=0):
        if rng_mode != 0:
            raise NotImplementedError("the C oracle implements the reference generator (MT19937-64) only")
        nx, ny, bins = camera.nx, camera.ny, config.bins
        if passes > 1:
            # `passes` observe() calls into an accumulating pipeline, starting from an empty frame
            fm, fv = np.zeros((nx, ny, bins)), np.zeros((nx, ny, bins))
            fs = np.zeros((nx, ny, bins), dtype=np.int32)
            total = 0
            for p in range(passes):
                m, v, r = self.render(camera, config, spectral, rng_mode, seed + p * seed_stride, pixels)
                total += r
                lib().ro_combine(C.c_int64(fm.size), _p(m), _p(v), C.c_int32(camera.pixel_samples), _p(fm), _p(fv), _p(fs))
            mean = np.zeros((nx, ny, bins)) if mean is None else mean
            variance = np.zeros((nx, ny, bins)) if variance is None else variance
            listed = np.ones((nx, ny), dtype=bool)
            if pixels is not None:
                pix = np.ascontiguousarray(pixels, dtype=np.int32).reshape(-1, 2)
                listed[:] = False
                listed[pix[:, 0], pix[:, 1]] = True
            mean[listed], variance[listed] = fm[listed], fv[listed]
            return mean, variance, total
        if mean is None:
            mean = np.zeros((nx, ny, bins))
        if variance is None:
            variance = np.zeros((nx, ny, bins))
        pix, n = None, nx * ny
        if pixels is not None:
            pix = np.ascontiguousarray(pixels, dtype=np.int32).reshape(-1, 2)
            n = pix.shape[0]
        rays = C.c_uint64(0)
        lib().ro_render(C.byref(self.flat.desc), C.byref(camera), C.byref(config), C.byref(spectral), C.c_uint64(seed), C.c_int64(n),
                        _p(pix), _p(mean), _p(variance), C.byref(rays))
        return mean, variance, rays.value
