#!/usr/bin/env python
"""Generates the committed golden vectors in tests/golden/ from the UNMODIFIED compiled reference
(oracle/_ref, built by oracle/build_ref.py from /root/reference).  Run in the build container:

    python tests/golden/make_golden.py

Inputs are regenerated deterministically by tests/scenes.py on both sides, so the fixtures hold only the
reference's OUTPUTS (plus digests of the reference-built kd-trees)."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from oracle import harness  # noqa: E402
import scenes  # noqa: E402

api = harness.ref_api()


def digest(b):
    return np.frombuffer(hashlib.sha256(b).digest(), dtype=np.uint8)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-28s %7.1f KB" % (name + ".npz", os.path.getsize(path) / 1024))


def assert_affine(world):
    """The device path carries rows 0..2 and m33 of every transform (RsbSceneDesc); the first three elements of the
    bottom row must be zero (projective matrices are rejected by the flattener)."""
    for p in world.primitives:
        for m in (p.to_local(), p.to_root()):
            assert [m[3, j] for j in range(3)] == [0.0, 0.0, 0.0], (type(p).__name__, [m[3, j] for j in range(4)])


def hits(world, o, d, md=None):
    assert_affine(world)
    r = harness.oracle_hit(world, o, d, md)
    stream = harness.world_kdtree_stream(world)
    return dict(primitive=r["primitive"], distance=r["distance"], exiting=r["exiting"], geometry=r["geometry"],
                triangle=r["triangle"], uvw=r["uvw"], tree_sha256=digest(stream), tree_bytes=np.int64(len(stream)))


def render(world, camera_kwargs, seed):
    cam, pipe = scenes.cornell_camera(api, world, **camera_kwargs)
    mean, var, n = harness.oracle_render(cam, pipe, seed)
    return cam, dict(mean=mean, variance=var, samples=n)


def passes_goldens():
    """Accumulated observe() passes (the reference's progressive-render loop, demos/cornell_box.py:160-174):
    the reference merges each pass into the frame with StatsArray3D.combine_samples."""
    world = scenes.cornell_box(api)
    cam, pipe = scenes.cornell_camera(api, world, pixels=(16, 12), samples=2, bins=16, spectral_rays=4, path_weight=0.5)
    mean, var, n = harness.oracle_render(cam, pipe, 999, passes=3)
    save("cornell_16x12_s2_p3_b16_r4", mean=mean, variance=var, samples=n)
    # one sample per pass walks the n in {0, 1} special cases of _combine_samples (statsarray.pyx:822-857)
    cam, pipe = scenes.cornell_camera(api, world, pixels=(12, 12), samples=1, bins=8)
    mean, var, n = harness.oracle_render(cam, pipe, 555, passes=4)
    save("cornell_12x12_s1_p4_b8", mean=mean, variance=var, samples=n)


def edge_goldens():
    """hand-placed rays and points on faces, edges, corners, tips, tangents, shared faces; empty and one-primitive worlds"""
    world = scenes.edge_scene(api)
    o, d, md = scenes.edge_rays()
    g = hits(world, o, d, md)
    cc, cp = harness.oracle_contains(world, scenes.edge_points())
    empty = api.World()
    ge = harness.oracle_hit(empty, o[:40], d[:40])
    ce, pe = harness.oracle_contains(empty, scenes.edge_points())
    one = api.World()
    api.Sphere(0.5, one, api.translate(3, 0, 0), api.AbsorbingSurface())
    g1 = harness.oracle_hit(one, o, d, md)
    c1, p1 = harness.oracle_contains(one, scenes.edge_points())
    save("edge_hits", **g, contains_count=cc, contains_prims=cp, empty_primitive=ge["primitive"], empty_contains=ce,
         one_primitive=g1["primitive"], one_distance=g1["distance"], one_contains=c1,
         empty_tree_sha256=digest(harness.world_kdtree_stream(empty)), one_tree_sha256=digest(harness.world_kdtree_stream(one)))
    print("edge rays: %d, hits %d; one-sphere hits %d" % (len(o), (g["primitive"] >= 0).sum(), (g1["primitive"] >= 0).sum()))


def metal_goldens():
    """Conductor (specular, complex index) on an analytic sphere and on a smooth-shaded mesh, UnitySurfaceEmitter"""
    world = scenes.metal_scene(api)
    cam, r = render(world, dict(pixels=(28, 24), samples=4, bins=12, path_weight=0.3), 2024)
    save("metal_28x24_s4_b12", **r)
    cam, r = render(world, dict(pixels=(16, 16), samples=3, bins=8, spectral_rays=4, importance=False), 77)
    save("metal_16x16_s3_b8_r4_noimp", **r)


def volume_goldens():
    """UniformVolumeEmitter / UnityVolumeEmitter: NullSurface transits (keep_alive, depth not counted) + emission * length"""
    world = scenes.volume_scene(api, fog=True)
    cam, r = render(world, dict(pixels=(24, 20), samples=4, bins=10, path_weight=0.3), 31)
    save("volume_fog_24x20_s4_b10", **r)
    world = scenes.volume_scene(api, fog=False)
    cam, r = render(world, dict(pixels=(16, 16), samples=3, bins=8, spectral_rays=2, min_depth=1, extinction=0.3, max_depth=4), 32)
    save("volume_16x16_s3_b8_r2_shallow", **r)


def ortho_goldens():
    """OrthographicCamera observer (parallel rays from the image plane, projection weight 1)"""
    world = scenes.cornell_box(api)
    cam, pipe = scenes.orthographic_camera(api, world, pixels=(20, 16), width=2.4, samples=3, bins=8, spectral_rays=2)
    mean, var, n = harness.oracle_render(cam, pipe, 808)
    save("ortho_20x16_s3_b8_r2", mean=mean, variance=var, samples=n)


def rough_goldens():
    """RoughConductor (GGX / Smith / conductor Fresnel) under multiple importance sampling and without it"""
    world = scenes.rough_metal_scene(api)
    cam, r = render(world, dict(pixels=(24, 24), samples=4, bins=10, path_weight=0.3), 606)
    save("rough_24x24_s4_b10", **r)
    cam, r = render(world, dict(pixels=(16, 12), samples=3, bins=6, spectral_rays=2, importance=False), 607)
    save("rough_16x12_s3_b6_r2_noimp", **r)


def scaled_goldens():
    """anisotropic scale + shear on every primitive type, a CSG tree and a mesh: hits, contains, a rendered frame"""
    world = scenes.scaled_scene(api)
    o, d = scenes.zoo_rays(4000, seed=21)
    g = hits(world, o, d)
    pts = np.random.default_rng(8).uniform([-2.6, -1.5, -1.2], [2.6, 2.9, 1.2], (3000, 3))
    cc, cp = harness.oracle_contains(world, pts)
    cam, r = render(world, dict(pixels=(20, 16), samples=3, bins=6, path_weight=0.3), 5150)
    save("scaled_hits_and_frame", **g, contains_count=cc, contains_prims=cp, **r)
    print("scaled scene: hits %d of %d, points inside %d" % ((g["primitive"] >= 0).sum(), len(o), (cc > 0).sum()))


def extremes_goldens():
    """corners of the ray / camera configuration space"""
    out = {}
    # A: roulette from the first daughter on (min depth 1, the smallest legal value), hard depth limit 3, wide fov,
    #    sensitivity != 1, tall frame
    world = scenes.cornell_box(api)
    cam, r = render(world, dict(pixels=(10, 14), samples=4, bins=5, min_depth=1, max_depth=3, extinction=0.3, fov=70.0,
                                sensitivity=2.5), 1)
    out.update({"a_" + k: v for k, v in r.items()})
    # B: only important paths (weight 1.0) with unequal importances (light 5, glass box 2, glass sphere 0)
    world = scenes.cornell_box(api)
    for p, imp in zip(world.primitives[-3:], (5.0, 2.0, 0.0)):
        p.material.importance = imp
    cam, r = render(world, dict(pixels=(12, 10), samples=4, bins=5, path_weight=1.0), 2)
    out.update({"b_" + k: v for k, v in r.items()})
    # C: never the important path (weight 0.0, MIS pdf still evaluated), camera INSIDE the glass sphere
    world = scenes.cornell_box(api)
    cam, pipe = scenes.cornell_camera(api, world, pixels=(10, 10), samples=4, bins=5, path_weight=0.0)
    cam.transform = api.translate(-0.4, -0.6, -0.45) * api.rotate(20, 10, 0)
    mean, var, n = harness.oracle_render(cam, pipe, 3)
    out.update(c_mean=mean, c_variance=var, c_samples=n)
    save("cornell_extremes", **out)
    print("light is primitive", [type(p.material).__name__ for p in world.primitives[-3:]])


def parabola_goldens():
    world = scenes.parabola_scene(api)
    o, d = scenes.parabola_rays()
    g = hits(world, o, d)
    pts = np.random.default_rng(9).uniform([-2.0, -1.3, -1.0], [2.0, 1.9, 1.0], (3000, 3))
    cc, cp = harness.oracle_contains(world, pts)
    cam, r = render(world, dict(pixels=(20, 16), samples=3, bins=5, path_weight=0.3), 8080)
    save("parabola_hits_and_frame", **g, contains_count=cc, contains_prims=cp, **r)
    print("parabola scene: hits %d of %d, points inside %d" % ((g["primitive"] >= 0).sum(), len(o), (cc > 0).sum()))


def nodes_goldens():
    """kd-node indices (north-star "kd-node indices bit-exact"): for every hit golden above, the world kd leaf in which
    the reference accepted the hit and the mesh kd leaf that produced the triangle (oracle/harness.oracle_hit_nodes:
    the reference's own leaf-visit sequence through the Python-subclassable KDTree3D, kdtree3d.pyx:993-1098)."""
    out = {}

    def add(tag, world, o, d, md=None):
        h = harness.oracle_hit(world, o, d, md)
        leaf, mesh_leaf, visits = harness.oracle_hit_nodes(world, o, d, md, hits=h)
        out[tag + "_leaf"], out[tag + "_mesh_leaf"] = leaf, mesh_leaf
        print("%-10s rays %5d hits %5d mesh hits %5d leaf visits %6d" % (tag, len(o), (leaf >= 0).sum(), (mesh_leaf >= 0).sum(), visits))
    o, d = scenes.zoo_rays(6000)
    add("zoo", scenes.primitive_zoo(api), o, d)
    add("zoo_md", scenes.primitive_zoo(api), o, d, np.random.default_rng(4).uniform(0.5, 7.0, len(o)))
    o, d, md = scenes.edge_rays()
    add("edge", scenes.edge_scene(api), o, d, md)
    rng = np.random.default_rng(1)
    o = np.tile(np.array([0, 0, -4.0]), (5000, 1))
    tgt = np.c_[rng.uniform(-1, 1, 5000), rng.uniform(-1, 1, 5000), np.zeros(5000)]
    d = tgt - o
    d /= np.linalg.norm(d, axis=1)[:, None]
    add("spheres", scenes.random_spheres(api, 2000, seed=7), np.ascontiguousarray(o), np.ascontiguousarray(d))
    o, d = scenes.mesh_rays(5000)
    add("mesh_smooth", scenes.mesh_scene(api, True), o, d)
    add("mesh_flat", scenes.mesh_scene(api, False), o, d)
    o, d = scenes.zoo_rays(4000, seed=21)
    add("scaled", scenes.scaled_scene(api), o, d)
    o, d = scenes.parabola_rays()
    add("parabola", scenes.parabola_scene(api), o, d)
    save("kd_nodes", **out)


def with_nodes(world, o, d, md=None):
    g = hits(world, o, d, md)
    leaf, mesh_leaf, visits = harness.oracle_hit_nodes(world, o, d, md, hits=g)
    g.update(leaf=leaf, mesh_leaf=mesh_leaf)
    return g


def bunny_goldens():
    """(1) the reference's own mesh fixture demos/resources/stanford_bunny.rsm (144,046 triangles, the depth-24 tree stored
    in the file); (2) BASELINE config 4: Cornell box + the bunny refined to 1,000,000 triangles -- mesh and tree written
    by the mirror's .rsm writer, LOADED BY THE REFERENCE (Mesh.from_file), hit by the reference"""
    import time
    t0 = time.time()
    world = scenes.bunny_rsm_scene(api)
    print("reference loaded stanford_bunny.rsm in %.1f s" % (time.time() - t0))
    o, d = scenes.bunny_rsm_rays()
    g = with_nodes(world, o, d)
    cc, cp = harness.oracle_contains(world, scenes.bunny_rsm_points())
    save("bunny_rsm_hits", **g, contains_count=cc, contains_prims=cp)
    print("bunny.rsm: hits %d of %d (mesh %d), points inside %d" % ((g["primitive"] >= 0).sum(), len(o), (g["triangle"] >= 0).sum(), (cc > 0).sum()))
    path = scenes.refined_bunny_rsm(1000000)
    t0 = time.time()
    world = scenes.cornell_mesh_scene(api, path)
    print("reference loaded %s in %.1f s" % (os.path.basename(path), time.time() - t0))
    o, d = scenes.cornell_mesh_rays()
    g = with_nodes(world, o, d)
    pts = np.random.default_rng(15).uniform([-0.5, -1.0, -0.4], [0.7, 0.05, 0.6], (1500, 3))
    cc, cp = harness.oracle_contains(world, pts)
    save("cornell_bunny_1m_hits", **g, contains_count=cc, contains_prims=cp, rsm_sha256=digest(open(path, "rb").read()))
    print("cornell + 1M bunny: hits %d of %d (mesh %d), points inside %d" % ((g["primitive"] >= 0).sum(), len(o), (g["triangle"] >= 0).sum(), (cc > 0).sum()))


def w_goldens():
    """transforms whose inverse has m33 = 1 - 1 ulp: Point3D.transform's division by w is live (point.pyx:272-281)"""
    world = scenes.w_scene(api)
    off = [p.to_local()[3, 3] - 1.0 for p in world.primitives]
    assert sum(1 for x in off if x != 0.0) >= 6, off
    o, d = scenes.zoo_rays(4000, seed=33)
    r = harness.oracle_hit(world, o, d)
    leaf, mesh_leaf, _ = harness.oracle_hit_nodes(world, o, d, hits=r)
    stream = harness.world_kdtree_stream(world)
    pts = np.random.default_rng(34).uniform([-2.6, -1.5, -1.2], [2.6, 2.9, 1.2], (3000, 3))
    cc, cp = harness.oracle_contains(world, pts)
    cam, pipe = scenes.cornell_camera(api, world, pixels=(20, 16), samples=3, bins=6, path_weight=0.3)
    cam.transform = api.translate(0, 0.3, -4.5) * scenes.w_transforms(api, 1, seed=5)[0]       # the observer's to_root as well
    mean, var, n = harness.oracle_render(cam, pipe, 6160)
    save("w_hits_and_frame", primitive=r["primitive"], distance=r["distance"], exiting=r["exiting"], geometry=r["geometry"],
         triangle=r["triangle"], uvw=r["uvw"], leaf=leaf, mesh_leaf=mesh_leaf, tree_sha256=digest(stream), tree_bytes=np.int64(len(stream)),
         contains_count=cc, contains_prims=cp, mean=mean, variance=var, samples=n, m33_minus_1=np.array(off))
    print("w scene: hits %d of %d, points inside %d, m33 - 1: %s" % ((r["primitive"] >= 0).sum(), len(o), (cc > 0).sum(), off))


def prism_512_golden():
    """BASELINE config 3 shape on a small frame: 512 spectral bins traced as 512 spectral rays (one bin per slice)"""
    world = scenes.prism_scene(api)
    cam, pipe = scenes.cornell_camera(api, world, pixels=(12, 10), samples=2, bins=512, spectral_rays=512, path_weight=0.75)
    cam.transform = api.translate(0.3, 0.2, -2.2) * api.rotate(5, -3, 0)
    mean, var, n = harness.oracle_render(cam, pipe, 2718)
    save("prism_12x10_s2_b512_r512", mean=mean, variance=var, samples=n)


def sweep_goldens():
    """BASELINE config 5 at size: 10,000 spheres drawn from the reference generator after seed(7); the first 20,000 rays
    of the device sweep (seed 2024; incoherent order and Morton order) restated in numpy, hit by the reference"""
    from raysect.core.math.random import seed, uniform
    seed(7)
    world = scenes.sweep_spheres(api, uniform)
    out = {}
    for tag, order in (("random", 0), ("morton", 7)):
        o, d, idx = scenes.sweep_rays(2024, 0, 20000, scenes.SWEEP_ORIGIN, scenes.SWEEP_TARGET, scenes.SWEEP_HALF, order)
        g = with_nodes(world, o, d)
        out.update({tag + "_" + k: g[k] for k in ("primitive", "distance", "leaf")})
        out["tree_sha256"], out["tree_bytes"] = g["tree_sha256"], g["tree_bytes"]
        print("sweep %s: hits %d of %d" % (tag, (g["primitive"] >= 0).sum(), len(o)))
    save("spheres10k_sweep", **out)


def main():
    if "--bunny-only" in sys.argv:
        return bunny_goldens()
    if "--w-only" in sys.argv:
        return w_goldens()
    if "--prism512-only" in sys.argv:
        return prism_512_golden()
    if "--sweep-only" in sys.argv:
        return sweep_goldens()
    if "--nodes-only" in sys.argv:
        return nodes_goldens()
    if "--parabola-only" in sys.argv:
        return parabola_goldens()
    if "--extremes-only" in sys.argv:
        return extremes_goldens()
    if "--scaled-only" in sys.argv:
        return scaled_goldens()
    if "--rough-only" in sys.argv:
        return rough_goldens()
    if "--ortho-only" in sys.argv:
        return ortho_goldens()
    if "--volume-only" in sys.argv:
        return volume_goldens()
    if "--metal-only" in sys.argv:
        return metal_goldens()
    if "--passes-only" in sys.argv:
        return passes_goldens()
    if "--edge-only" in sys.argv:
        return edge_goldens()
    # 1. RNG known answers: the reference's own test vector (raysect/core/math/tests/test_random.py:37-253)
    from raysect.core.math.tests.test_random import _random_reference
    kat = np.array(_random_reference)
    assert np.array_equal(harness.oracle_uniform(1234567890, len(kat)), kat)
    save("rng_kat", seed=np.uint64(1234567890), uniform=kat, seed77=harness.oracle_uniform(77, 700))

    # 2. analytic primitives + CSG
    world = scenes.primitive_zoo(api)
    o, d = scenes.zoo_rays(6000)
    g = hits(world, o, d)
    md = np.random.default_rng(4).uniform(0.5, 7.0, len(o))
    g2 = harness.oracle_hit(world, o, d, md)
    pts = np.random.default_rng(5).uniform([-2.8, -2.6, -0.6], [2.8, 2.2, 0.8], (4000, 3))
    cc, cp = harness.oracle_contains(world, pts)
    save("zoo_hits", **g, md_primitive=g2["primitive"], md_distance=g2["distance"], contains_count=cc, contains_prims=cp)

    # 3. config-5 style sphere field (2,000 spheres for the fixture; the bench uses 10,000)
    world = scenes.random_spheres(api, 2000, seed=7)
    rng = np.random.default_rng(1)
    o = np.tile(np.array([0, 0, -4.0]), (5000, 1))
    tgt = np.c_[rng.uniform(-1, 1, 5000), rng.uniform(-1, 1, 5000), np.zeros(5000)]
    d = tgt - o
    d /= np.linalg.norm(d, axis=1)[:, None]
    save("spheres_hits", **hits(world, np.ascontiguousarray(o), np.ascontiguousarray(d)))

    # 4. triangle mesh (float32 watertight test, smoothing normals, instancing, contains)
    for smoothing in (True, False):
        world = scenes.mesh_scene(api, smoothing)
        o, d = scenes.mesh_rays(5000)
        g = hits(world, o, d)
        pts = np.random.default_rng(6).uniform([-1.0, -0.8, -0.5], [1.1, 0.6, 0.8], (3000, 3))
        cc, cp = harness.oracle_contains(world, pts)
        import io
        buf = io.BytesIO()
        world.primitives[0].data.save(buf)
        from source_b200.flatten import rsm_kdtree_stream
        blob = buf.getvalue()
        stream = blob[rsm_kdtree_stream(blob):]
        save("mesh_hits_smooth" if smoothing else "mesh_hits_flat", **g, contains_count=cc, contains_prims=cp,
             mesh_tree_sha256=digest(stream), mesh_tree_bytes=np.int64(len(stream)),
             face_normals=np.array(world.primitives[0].data.face_normals))

    # 5. Cornell box renders (SerialEngine semantics, per-pixel re-seeding; seed base 1000)
    world = scenes.cornell_box(api)
    cam, r = render(world, dict(pixels=(32, 32), samples=4, bins=15), 1000)
    save("cornell_32x32_s4_b15", **r)
    cam, r = render(world, dict(pixels=(16, 12), samples=3, bins=16, spectral_rays=4, path_weight=0.5), 4242)
    save("cornell_16x12_s3_b16_r4", **r)
    world = scenes.cornell_box(api, glass=False)
    cam, r = render(world, dict(pixels=(24, 24), samples=2, bins=8, importance=False, min_depth=2, max_depth=6, extinction=0.2), 77)
    save("cornell_noglass_noimp_24", **r)

    passes_goldens()
    edge_goldens()
    metal_goldens()
    volume_goldens()
    ortho_goldens()
    rough_goldens()
    scaled_goldens()
    extremes_goldens()
    parabola_goldens()
    nodes_goldens()
    bunny_goldens()
    sweep_goldens()
    w_goldens()
    prism_512_golden()

    # 6. dispersive CSG prism: one spectral ray per bin
    world = scenes.prism_scene(api)
    cam, pipe = scenes.cornell_camera(api, world, pixels=(24, 24), samples=6, bins=6, spectral_rays=6, path_weight=0.75)
    cam.transform = api.translate(0.3, 0.2, -2.2) * api.rotate(5, -3, 0)
    mean, var, n = harness.oracle_render(cam, pipe, 31337)
    save("prism_24x24_s6_b6_r6", mean=mean, variance=var, samples=n)

    # 7. host object model: transforms, bounding volumes, spectral resampling
    world = scenes.primitive_zoo(api)
    rows = []
    for p in world.primitives:
        b, s = p.bounding_box(), p.bounding_sphere()
        rows.append([p.to_local()[i, j] for i in range(3) for j in range(4)] + [p.to_root()[i, j] for i in range(3) for j in range(4)]
                    + [b.lower.x, b.lower.y, b.lower.z, b.upper.x, b.upper.y, b.upper.z, s.centre.x, s.centre.y, s.centre.z, s.radius])
    glass = api.schott("N-BK7")
    sf11 = api.schott("SF11")
    white = api.InterpolatedSF(scenes.CB_WAVELENGTHS, scenes.CB_WHITE)
    light = api.InterpolatedSF(*scenes.CB_LIGHT)
    spec = {}
    for tag, (lo, hi, bins) in dict(a=(375.0, 740.0, 15), b=(400.0, 700.0, 64), c=(520.5, 530.25, 3), d=(300.0, 420.0, 7)).items():
        spec["white_" + tag] = np.array(white.sample(lo, hi, bins))
        spec["light_" + tag] = np.array(light.sample(lo, hi, bins))
        spec["bk7_t_" + tag] = np.array(glass.transmission.sample(lo, hi, bins))
        spec["bk7_n_" + tag] = np.float64(glass.index.average(lo, hi))
        spec["sf11_t_" + tag] = np.array(sf11.transmission.sample(lo, hi, bins))
        spec["sf11_n_" + tag] = np.float64(sf11.index.average(lo, hi))
    save("object_model", zoo_rows=np.array(rows), **spec)


if __name__ == "__main__":
    main()
