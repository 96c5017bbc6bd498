"""Drivers for the compiled reference in ``oracle/_ref`` -- TEST INFRASTRUCTURE ONLY.

Nothing under ``source_b200/`` imports this module.  It is used by tests (when ``oracle/_ref`` is
present), by ``tests/golden/make_golden.py`` to generate the committed fixtures, and by ``bench.py``
for the CPU baseline / reference arm.
"""
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
STUB = os.path.join(HERE, "mpl_stub")


def available():
    return os.path.exists(os.path.join(REF, "raysect", "_built_ok"))


def _activate():
    if not available():
        raise ImportError("oracle/_ref is not built (run `python oracle/build_ref.py` where /root/reference exists)")
    for p in (STUB, REF):
        if p not in sys.path:
            sys.path.insert(0, p)


def ref_api():
    """Namespace with the reference's classes under the names tests/scenes.py expects."""
    _activate()
    from raysect.core import AffineMatrix3D, Point3D, Vector3D, rotate, rotate_x, rotate_y, rotate_z, translate
    from raysect.core.ray import Ray as CoreRay
    from raysect.optical import ConstantSF, InterpolatedSF, Node, World
    from raysect.optical.library import schott
    from raysect.optical.material import (AbsorbingSurface, Conductor, Dielectric, Lambert, RoughConductor, Sellmeier,
                                          UniformSurfaceEmitter, UniformVolumeEmitter, UnitySurfaceEmitter, UnityVolumeEmitter)
    from raysect.optical.observer import FullFrameSampler2D, OrthographicCamera, PinholeCamera, SpectralPowerPipeline2D
    from raysect.primitive import Box, Cone, Cylinder, Intersect, Mesh, Parabola, Sphere, Subtract, Union
    ns = types.SimpleNamespace(**{k: v for k, v in locals().items() if k != "ns"})
    ns.Ray = CoreRay
    return ns


def world_kdtree_stream(world):
    """Serialised _PrimitiveKDTree(world.primitives) (raysect/core/acceleration/kdtree.pyx:41-58)"""
    _activate()
    from raysect.core.acceleration.kdtree import _PrimitiveKDTree
    buf = io.BytesIO()
    _PrimitiveKDTree(list(world.primitives)).save(buf)
    return buf.getvalue()


def oracle_hit(world, origins, directions, max_distance=None):
    """World.hit per ray -> dict of arrays (primitive id in world.primitives, ray_distance, exiting,
    hit/inside/outside/normal in primitive-local space, mesh triangle + barycentrics)."""
    _activate()
    from raysect.core import Point3D, Vector3D
    from raysect.core.ray import Ray as CoreRay
    n = len(origins)
    index = {id(p): i for i, p in enumerate(world.primitives)}
    prim = np.full(n, -1, dtype=np.int32)
    t = np.full(n, np.inf)
    exiting = np.zeros(n, dtype=np.uint8)
    geom = np.zeros((n, 12))
    tri = np.full(n, -1, dtype=np.int32)
    uvw = np.zeros((n, 3), dtype=np.float32)
    world.build_accelerator()
    for i in range(n):
        md = float("inf") if max_distance is None else float(max_distance[i])
        it = world.hit(CoreRay(Point3D(*origins[i]), Vector3D(*directions[i]), md))
        if it is None:
            continue
        prim[i] = index[id(it.primitive)]
        t[i] = it.ray_distance
        exiting[i] = 1 if it.exiting else 0
        geom[i] = [it.hit_point.x, it.hit_point.y, it.hit_point.z, it.inside_point.x, it.inside_point.y, it.inside_point.z,
                   it.outside_point.x, it.outside_point.y, it.outside_point.z, it.normal.x, it.normal.y, it.normal.z]
        if hasattr(it, "triangle"):
            tri[i] = it.triangle
            uvw[i] = [it.u, it.v, it.w]
    return dict(primitive=prim, distance=t, exiting=exiting, geometry=geom, triangle=tri, uvw=uvw)


def oracle_contains(world, points, cap=8):
    _activate()
    from raysect.core import Point3D
    index = {id(p): i for i, p in enumerate(world.primitives)}
    n = len(points)
    count = np.zeros(n, dtype=np.int32)
    prims = np.full((n, cap), -1, dtype=np.int32)
    for i in range(n):
        lst = world.contains(Point3D(*points[i]))
        count[i] = len(lst)
        for k, p in enumerate(lst[:cap]):
            prims[i, k] = index[id(p)]
    return count, prims


def oracle_uniform(seed_value, n):
    """seed(seed_value); n x uniform() (raysect/core/math/random.pyx:215-265)"""
    _activate()
    from raysect.core.math.random import seed, uniform
    seed(seed_value)
    return np.array([uniform() for _ in range(n)])


def make_reseeding_engine(seed_base, nx):
    """A RenderEngine (raysect/core/workflow.py:35-97) that re-seeds the reference's global MT19937-64
    stream with seed(seed_base + y*nx + x) before rendering each pixel task -- the per-pixel stream
    definition the device path implements (include/raysect_b200.h, RSB_RNG_MT19937_64)."""
    _activate()
    from raysect.core.math.random import seed
    from raysect.core.workflow import RenderEngine

    class ReseedingEngine(RenderEngine):
        def run(self, tasks, render, update, render_args=(), render_kwargs={}, update_args=(), update_kwargs={}):
            for task in tasks:
                x, y = task
                seed(seed_base + y * nx + x)
                update(render(task, *render_args, **render_kwargs), *update_args, **update_kwargs)

        def worker_count(self):
            return 1

    return ReseedingEngine()


def oracle_render(camera, pipeline, seed_base, passes=1):
    """camera.observe() with per-pixel re-seeding.  One pass per spectral slice; slice k uses
    seed_base + k*nx*ny, matching source_b200.observer.PinholeCamera.observe.  ``passes`` > 1 calls observe()
    that many times into the accumulating pipeline (the reference's progressive-render loop,
    demos/cornell_box.py:160-174); call p offsets the seed by p*n_slices*nx*ny."""
    _activate()
    nx, ny = camera.pixels
    slices = camera.spectral_rays
    state = {"pass": 0}

    class SliceAwareEngine(type(make_reseeding_engine(seed_base, nx))):
        def run(self, tasks, render, update, render_args=(), render_kwargs={}, update_args=(), update_kwargs={}):
            from raysect.core.math.random import seed
            slice_id = render_args[0] if render_args else 0
            for task in tasks:
                x, y = task
                seed(seed_base + (state["pass"] * slices + slice_id) * nx * ny + y * nx + x)
                update(render(task, *render_args, **render_kwargs), *update_args, **update_kwargs)

    camera.render_engine = SliceAwareEngine()
    if passes > 1:
        pipeline.accumulate = True
    for p in range(passes):
        state["pass"] = p
        camera.observe()
    f = pipeline.frame
    return np.array(f.mean), np.array(f.variance), np.array(f.samples)
