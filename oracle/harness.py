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


# ---- kd-node indices (north-star: "hit primitive IDs / kd-node indices bit-exact") ---------------------------------
def parse_kdtree_stream(stream):
    """KDTree3DCore.save() byte stream (kdtree3d.pyx:864-912) -> dict(bounds, type[], split[], count[], items[list])
    with the reference's node numbering (node id = position in the stream, lower child = id + 1, upper child = count)."""
    import struct
    off = 0
    max_depth, min_items = struct.unpack_from("<ii", stream, off); off += 8
    hit_cost, empty_bonus = struct.unpack_from("<dd", stream, off); off += 16
    bounds = struct.unpack_from("<6d", stream, off); off += 48
    (n,) = struct.unpack_from("<i", stream, off); off += 4
    types, splits, counts, items = [0] * n, [0.0] * n, [0] * n, [None] * n
    for i in range(n):
        (t,) = struct.unpack_from("<i", stream, off); off += 4
        types[i] = t
        if t == -1:     # LEAF
            (c,) = struct.unpack_from("<i", stream, off); off += 4
            counts[i] = c
            items[i] = list(struct.unpack_from("<%di" % c, stream, off)); off += 4 * c
        else:
            (s,) = struct.unpack_from("<d", stream, off); off += 8
            (c,) = struct.unpack_from("<i", stream, off); off += 4
            splits[i], counts[i] = s, c
    assert off == len(stream)
    return dict(bounds=bounds, type=types, split=splits, count=counts, items=items, max_depth=max_depth)


def _slab(origin, direction, lower, upper, front, back):
    """BoundingBox3D._slab (boundingbox.pyx:200-245)"""
    if direction != 0.0:
        reciprocal = 1.0 / direction
        if direction > 0:
            tmin, tmax = (lower - origin) * reciprocal, (upper - origin) * reciprocal
        else:
            tmin, tmax = (upper - origin) * reciprocal, (lower - origin) * reciprocal
    else:
        if origin < lower:
            tmin = tmax = -np.inf
        elif origin > upper:
            tmin = tmax = np.inf
        else:
            tmin, tmax = -np.inf, np.inf
    return max(front, tmin) if tmin > front else front, min(back, tmax) if tmax < back else back


def kd_leaf_sequence(tree, o, d):
    """KDTree3DCore._trace / _trace_branch (kdtree3d.pyx:589-700) restated over a parsed stream, leaf tests left out:
    the leaves the walk would visit if no leaf ever reported a hit, in order, as (node id, max_range).  The real walk
    stops at the first leaf that reports a hit, i.e. it visits a prefix of this list."""
    lo, hi = tree["bounds"][:3], tree["bounds"][3:]
    front, back = -np.inf, np.inf
    for k in range(3):
        front, back = _slab(o[k], d[k], lo[k], hi[k], front, back)
    if front > back or (front < 0.0 and back < 0.0):
        return []
    out = []
    stack = [(0, front, back)]
    types, splits, counts = tree["type"], tree["split"], tree["count"]
    while stack:
        node, min_range, max_range = stack.pop()
        while types[node] != -1:
            axis, split = types[node], splits[node]
            lower_id, upper_id = node + 1, counts[node]
            origin, direction = o[axis], d[axis]
            if direction == 0:
                node = lower_id if origin < split else upper_id
                continue
            plane_distance = (split - origin) / direction
            below = origin < split or (origin == split and direction < 0)
            near_id, far_id = (lower_id, upper_id) if below else (upper_id, lower_id)
            if plane_distance > max_range or plane_distance <= 0:
                node = near_id
            elif plane_distance < min_range:
                node = far_id
            else:
                stack.append((far_id, plane_distance, max_range))
                node, max_range = near_id, plane_distance
        out.append((node, max_range))
    return out


def oracle_leaf_visits(world, origins, directions, max_distance=None):
    """Leaf visits of World.hit observed on the compiled reference: a Python subclass of KDTree3D built
    from the world's primitive boxes serialises byte-identically to _PrimitiveKDTree(world.primitives)
    (acceleration/kdtree.pyx:41-58) and its _trace_items override is called once per visited leaf, in traversal
    order; inside the hook the leaf rule of kdtree.pyx:103-122 is replayed on the reference's own BoundPrimitive /
    Primitive.hit.  Returns per ray the list of (item ids, max_range) visited and the primitive index / distance
    of the hit (-1, inf: miss)."""
    _activate()
    from raysect.core import Point3D, Vector3D
    from raysect.core.acceleration.boundprimitive import BoundPrimitive
    from raysect.core.math.spatial.kdtree3d import Item3D, KDTree3D
    from raysect.core.ray import Ray as CoreRay
    prims = list(world.primitives)
    bound = [BoundPrimitive(p) for p in prims]   # .box / .primitive are readable; .hit is cdef: replayed below

    class Recorder(KDTree3D):
        def __init__(self):
            items = [Item3D(i, b.box) for i, b in enumerate(bound)]
            super().__init__(items, max_depth=0, min_items=1, hit_cost=80.0, empty_bonus=0.2)
            self.visits, self.hit = [], None

        def _trace_items(self, item_ids, ray, max_range):
            self.visits.append((list(item_ids), max_range))
            distance = min(ray.max_distance, max_range)
            closest = None
            for item in item_ids:
                # BoundPrimitive.hit (boundprimitive.pyx:42-51)
                it = bound[item].primitive.hit(ray) if bound[item].box.hit(ray) else None
                if it is not None and it.ray_distance <= distance:
                    distance = it.ray_distance
                    closest = (item, it)
            if closest is None:
                return False
            self.hit = closest
            return True

    tree = Recorder()
    buf = io.BytesIO()
    tree.save(buf)
    assert buf.getvalue() == world_kdtree_stream(world), "the Python-subclassed tree is not the world's tree"
    out = []
    for i in range(len(origins)):
        md = float("inf") if max_distance is None else float(max_distance[i])
        tree.visits, tree.hit = [], None
        hit = tree.trace(CoreRay(Point3D(*origins[i]), Vector3D(*directions[i]), md))
        out.append((tree.visits, (tree.hit[0], tree.hit[1].ray_distance, tree.hit[1]) if hit else (-1, np.inf, None)))
    return buf.getvalue(), out


def oracle_hit_nodes(world, origins, directions, max_distance=None, hits=None):
    """kd-node indices of World.hit: for every ray the world kd leaf (reference node id) in which the hit was accepted
    and, for mesh hits, the mesh kd leaf that produced the triangle; -1 where there is none.

    World leaf: the reference's own leaf-visit sequence (oracle_leaf_visits) is matched, visit by visit, with the walk
    of the serialised tree (kd_leaf_sequence) -- item lists and max_range must agree bit for bit, which pins node ids
    to visits; the hit leaf is the last one visited.  Mesh leaf: MeshData is a cdef subclass (no Python hook), so the
    same stream walk -- validated on the world tree just before -- is run on the mesh's own tree with the mesh-local
    ray, and the accepting leaf is the first one visited that lists the reference's hit triangle with
    float32 t < min(ray.max_distance, max_range) (MeshData._trace_leaf, mesh.pyx:520-563)."""
    _activate()
    from raysect.core import Point3D, Vector3D
    stream, visits = oracle_leaf_visits(world, origins, directions, max_distance)
    tree = parse_kdtree_stream(stream)
    prims = list(world.primitives)
    mesh_trees = {}
    n = len(origins)
    leaf = np.full(n, -1, dtype=np.int32)
    mesh_leaf = np.full(n, -1, dtype=np.int32)
    n_visits = 0
    for i in range(n):
        seq, (pid, t, it) = visits[i]
        walk = kd_leaf_sequence(tree, [float(x) for x in origins[i]], [float(x) for x in directions[i]])
        assert len(seq) <= len(walk), "ray %d: the reference visited more leaves than the stream walk" % i
        for (ids, mr), (node, mr2) in zip(seq, walk):
            assert ids == tree["items"][node] and mr == mr2, "ray %d: leaf visit mismatch" % i
        n_visits += len(seq)
        if hits is not None:
            assert pid == int(hits["primitive"][i]) and (pid < 0 or t == hits["distance"][i]), "ray %d: hook hit != World.hit" % i
        if pid < 0:
            assert len(seq) == len(walk)
            continue
        leaf[i] = walk[len(seq) - 1][0]
        if hasattr(it, "triangle"):
            p = prims[pid]
            if id(p.data) not in mesh_trees:
                buf = io.BytesIO()
                p.data.save(buf)
                blob = buf.getvalue()
                from source_b200.flatten import rsm_kdtree_stream
                mesh_trees[id(p.data)] = parse_kdtree_stream(blob[rsm_kdtree_stream(blob):])
            mt = mesh_trees[id(p.data)]
            lo = Point3D(*origins[i]).transform(p.to_local())
            ld = Vector3D(*directions[i]).transform(p.to_local())
            md = float("inf") if max_distance is None else float(max_distance[i])
            t32 = float(np.float32(it.ray_distance))
            assert t32 == it.ray_distance
            for node, mr in kd_leaf_sequence(mt, [lo.x, lo.y, lo.z], [ld.x, ld.y, ld.z]):
                if it.triangle in mt["items"][node] and t32 < min(md, mr):
                    mesh_leaf[i] = node
                    break
            assert mesh_leaf[i] >= 0, "ray %d: no mesh leaf accepts the reference's triangle" % i
    return leaf, mesh_leaf, n_visits
