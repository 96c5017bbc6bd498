"""GPU suite: the CUDA kernels, called through the C ABI, against the golden vectors of the compiled
reference.  ids, kd trees, t and local-space geometry are bit-exact; rendered frames are held to the
north-star tolerance (1e-6 relative) with the path-divergence rate reported."""
import os

import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def make_backend(device):
    from source_b200.engine import Accelerator
    return lambda flat: Accelerator(device, flat)


def test_device_is_b200(device):
    assert device.cc[0] == 10
    assert device.sm_count >= 100


def test_rng_known_answers(device):
    g = parity.golden("rng_kat")
    np.testing.assert_array_equal(device.rng_uniform(int(g["seed"]), len(g["uniform"])), g["uniform"])
    np.testing.assert_array_equal(device.rng_uniform(77, 700), g["seed77"])


def test_zoo_hits_contains(make_backend):
    parity.zoo(make_backend)


def test_edge_cases(make_backend):
    parity.edge(make_backend)


def test_non_rigid_transforms(make_backend):
    parity.scaled(make_backend, exact=False, rtol=1e-6, max_divergent_fraction=0.0)


def test_configuration_extremes(make_backend):
    parity.extremes(make_backend, exact=False, rtol=1e-6, max_divergent_fraction=0.0)


def test_parabola_primitive(make_backend):
    parity.parabola(make_backend, exact=False, rtol=1e-6, max_divergent_fraction=0.0)


def test_transforms_with_m33_off_one(make_backend):
    fr = parity.w_matrices(make_backend, exact=False, rtol=1e-6, max_divergent_fraction=0.0)
    print("divergent pixel fraction:", fr)


def test_prism_512_spectral_slices_rendered_concurrently(make_backend):
    fr = parity.prism_512(make_backend, exact=False, rtol=1e-6, max_divergent_fraction=0.0)
    print("divergent pixel fraction:", fr)


def test_sphere_field(make_backend):
    parity.spheres(make_backend)


@pytest.mark.parametrize("smoothing", [True, False])
def test_mesh(make_backend, smoothing):
    parity.mesh(make_backend, smoothing)


def test_cornell_frames(make_backend):
    fr = parity.cornell(make_backend, exact=False, rtol=1e-6, max_divergent_fraction=0.0)
    print("divergent pixel fractions:", fr)


def test_accumulated_passes_vs_reference(make_backend):
    """rsb_render_passes: P observe() passes rendered concurrently vs the reference calling observe() P times"""
    fr = parity.cornell_passes(make_backend, exact=False, rtol=1e-6, max_divergent_fraction=0.0)
    print("divergent pixel fractions:", fr)


@pytest.mark.parametrize("rng", ["mt", "philox"])
def test_concurrent_passes_equal_sequential_passes_bit_for_bit(device, make_backend, monkeypatch, rng):
    """Concurrent passes (one wavefront over (pass, pixel) streams, several chunks, a masked task list) must equal
    the same passes rendered one rsb_render call at a time and merged on the host with combine_samples -- exactly:
    both sides run the same device code, so no libm tolerance is involved."""
    import scenes
    import source_b200 as api
    from source_b200.engine import Accelerator, Device, camera_desc, ray_config
    from source_b200.observer import StatsArray3D
    monkeypatch.setenv("RSB_CHUNK_ITEMS", "5500")   # 5 passes x 40 x 30 pixels = 6000 work items -> 2 chunks, split inside pass 4
    monkeypatch.setenv("RSB_SLOTS_PER_SM", "32")    # 4,736 slots < 5,500 items: slots pick up items of later passes
    dev = Device(device.index)
    world = scenes.cornell_box(api)
    accel = Accelerator(dev, parity.flatten_world(world))
    nx, ny, bins, spp, passes, seed = 40, 30, 12, 3, 5, 4321
    cam_t = api.translate(0, 0, -3.3)
    cam = camera_desc(nx, ny, spp, 45.0, 1.0, cam_t)
    cfg = ray_config(bins, 375.0, 740.0, 0.01, 3, 500, True, 0.25)
    spectral = accel.flat.spectral(375.0, 740.0, bins)
    from source_b200 import _cabi as cabi
    mode = cabi.RNG_MT19937_64 if rng == "mt" else cabi.RNG_PHILOX
    for pixels in (None, np.argwhere(np.random.default_rng(0).random((nx, ny)) < 0.4).astype(np.int32)):
        m, v, rays = accel.render(cam, cfg, spectral, mode, seed, pixels, passes=passes, seed_stride=nx * ny)
        frame = StatsArray3D(nx, ny, bins)
        total = 0
        for p in range(passes):
            mp, vp, rp = accel.render(cam, cfg, spectral, mode, seed + p * nx * ny, pixels)
            frame.combine_slice(pixels, 0, mp, vp, spp)
            total += rp
        assert rays == total
        np.testing.assert_array_equal(m, frame.mean)
        np.testing.assert_array_equal(v, frame.variance)
        if pixels is not None:
            untouched = np.ones((nx, ny), dtype=bool)
            untouched[pixels[:, 0], pixels[:, 1]] = False
            assert np.all(m[untouched] == 0) and np.all(v[untouched] == 0)
    accel.close()
    dev.close()


def test_conductor_and_unity_emitter(make_backend):
    fr = parity.metal(make_backend, exact=False, rtol=1e-6, max_divergent_fraction=0.0)
    print("divergent pixel fractions:", fr)


def test_volume_emitters(make_backend):
    fr = parity.volumes(make_backend, exact=False, rtol=1e-6, max_divergent_fraction=0.0)
    print("divergent pixel fractions:", fr)


def test_orthographic_camera(make_backend):
    fr = parity.orthographic(make_backend, exact=False, rtol=1e-6, max_divergent_fraction=0.0)
    print("divergent pixel fraction:", fr)


def test_rough_conductor(make_backend):
    fr = parity.rough_metal(make_backend, exact=False, rtol=1e-6, max_divergent_fraction=0.0)
    print("divergent pixel fractions:", fr)


def test_prism_csg_dispersion(make_backend):
    fr = parity.prism(make_backend, exact=False, rtol=1e-6, max_divergent_fraction=0.0)
    print("divergent pixel fraction:", fr)


def test_gpu_matches_host_build_of_same_source(make_backend, lib):
    """Same source, two compilers: the CUDA build and the host build must agree on a seeded frame to
    the same tolerance (differences come only from libm: sin/cos/asin/pow)."""
    import hostsim_api
    import scenes
    import source_b200 as api
    world = scenes.cornell_box(api)
    _, f_gpu = parity.observe(make_backend, world, 99, pixels=(48, 40), samples=8, bins=32)
    _, f_cpu = parity.observe(hostsim_api.HostScene, world, 99, pixels=(48, 40), samples=8, bins=32)
    g = dict(mean=f_cpu.mean, variance=f_cpu.variance, samples=f_cpu.samples)
    fr = parity.compare_frame(f_gpu, g, exact=False, rtol=1e-6, max_divergent_fraction=0.0)
    print("divergent pixel fraction vs host build:", fr)


def test_philox_mode_statistics(make_backend):
    import scenes
    import source_b200 as api
    from source_b200 import _cabi as cabi
    g = parity.golden("cornell_32x32_s4_b15")
    world = scenes.cornell_box(api)
    cam, frame = parity.observe(make_backend, world, 5, rng_mode=cabi.RNG_PHILOX, pixels=(32, 32), samples=4, bins=15)
    total, total_ref = frame.mean.sum(), g["mean"].sum()
    sigma = np.sqrt((g["variance"] / 4).sum())
    assert abs(total - total_ref) < 5 * np.sqrt(2) * sigma


def test_partial_pixel_list_leaves_other_pixels_untouched(make_backend):
    import scenes
    import source_b200 as api
    world = scenes.cornell_box(api)
    mask = np.zeros((16, 16), dtype=bool)
    mask[3:9, 5:11] = True
    cam, pipe = scenes.cornell_camera(api, world, pixels=(16, 16), samples=2, bins=4)
    cam.frame_sampler = api.FullFrameSampler2D(mask)
    world._accel = parity._Accel(make_backend(parity.flatten_world(world)))
    world._rebuild = False
    cam.observe()
    f = pipe.frame
    assert np.all(f.samples[~mask] == 0) and np.all(f.mean[~mask] == 0)
    assert np.all(f.samples[mask] == 2)


def test_errors_are_loud(device):
    from source_b200 import RsbError
    with pytest.raises(RsbError):
        device.rng_uniform(0, 4)          # seed(0) means "reseed from urandom" in the reference: rejected


def test_plugin_on_device_against_live_reference(device, reference):
    """Real Raysect objects + CudaAccelerator / CudaRenderEngine on the B200 vs Raysect's own KDTree + serial render"""
    import scenes
    api = reference.ref_api()      # puts oracle/_ref on sys.path
    from raysect.core import Point3D, Vector3D
    from source_b200.plugin import CudaAccelerator, CudaRenderEngine
    world = scenes.primitive_zoo(api)
    o, d = scenes.zoo_rays(400)
    ref = reference.oracle_hit(world, o, d)
    acc = CudaAccelerator(device=device)
    world.accelerator = acc
    world.build_accelerator(force=True)
    r = acc.hit_batch(o, d, geometry=True)
    parity.check_hits(r, ref)
    it = world.hit(api.Ray(Point3D(*o[0]), Vector3D(*d[0])))
    assert (it is None) == (ref["primitive"][0] < 0)
    kw = dict(pixels=(20, 16), samples=4, bins=12, spectral_rays=2)
    w1 = scenes.cornell_box(api)
    cam, pipe = scenes.cornell_camera(api, w1, **kw)
    m_ref, v_ref, n_ref = reference.oracle_render(cam, pipe, 999)
    w2 = scenes.cornell_box(api)
    cam2, pipe2 = scenes.cornell_camera(api, w2, **kw)
    cam2.render_engine = CudaRenderEngine(seed=999, rng="mt", device=device)
    cam2.observe()
    class F:  # noqa: E701
        mean, variance, samples = np.array(pipe2.frame.mean), np.array(pipe2.frame.variance), np.array(pipe2.frame.samples)
    fr = parity.compare_frame(F, dict(mean=m_ref, variance=v_ref, samples=n_ref), exact=False, rtol=1e-6, max_divergent_fraction=0.0)
    print("plugin render divergent fraction", fr)


@pytest.mark.parametrize("passes,rgb_only", [(1, False), (2, False), (2, True)])
def test_rgb_pipeline_on_device_against_live_reference(device, reference, passes, rgb_only):
    """RGBPipeline2D (+ a spectral pipeline) through CudaRenderEngine on the B200 vs the reference's serial render feeding
    the same pipelines: the XYZ frame within 1e-6 relative (no divergent pixel), sample counts exact; with the RGB pipeline
    alone the device keeps no spectral frame at all."""
    import scenes
    api = reference.ref_api()
    from raysect.optical.observer import RGBPipeline2D
    from source_b200.plugin import CudaRenderEngine, WholeFrameSampler2D
    kw = dict(pixels=(20, 16), bins=12, spectral_rays=2)
    w1 = scenes.cornell_box(api)
    cam, pipe = scenes.cornell_camera(api, w1, samples=3, sensitivity=1.7, **kw)
    rgb = RGBPipeline2D(display_progress=False, accumulate=passes > 1)
    cam.pipelines = [pipe, rgb]
    m_ref, v_ref, n_ref = reference.oracle_render(cam, pipe, 31337, passes=passes)
    x_ref = dict(mean=np.array(rgb.xyz_frame.mean), variance=np.array(rgb.xyz_frame.variance), samples=np.array(rgb.xyz_frame.samples))
    assert x_ref["mean"].max() > 0
    w2 = scenes.cornell_box(api)
    cam2, pipe2 = scenes.cornell_camera(api, w2, samples=3 * passes, sensitivity=1.7, **kw)
    rgb2 = RGBPipeline2D(display_progress=False)
    cam2.pipelines = [rgb2] if rgb_only else [pipe2, rgb2]
    cam2.frame_sampler = WholeFrameSampler2D()
    cam2.render_engine = CudaRenderEngine(seed=31337, rng="mt", device=device, passes=passes)
    cam2.observe()

    class X:  # noqa: E701
        mean, variance, samples = np.array(rgb2.xyz_frame.mean), np.array(rgb2.xyz_frame.variance), np.array(rgb2.xyz_frame.samples)
    print("xyz divergent fraction", parity.compare_frame(X, x_ref, exact=False, rtol=1e-6, max_divergent_fraction=0.0))
    if not rgb_only:
        class F:  # noqa: E701
            mean, variance, samples = np.array(pipe2.frame.mean), np.array(pipe2.frame.variance), np.array(pipe2.frame.samples)
        parity.compare_frame(F, dict(mean=m_ref, variance=v_ref, samples=n_ref), exact=False, rtol=1e-6, max_divergent_fraction=0.0)


def test_two_devices_in_one_process_equal_one_device(device, reference):
    """CudaRenderEngine(devices=[0, 1]) (engine.DeviceGroup: tiles dealt to both GPUs, one host thread each, the spectral
    rows of device 1 pulled into device 0's slice by k_gather_peer_rows over peer memory, XYZ frames merged member by
    member) gives bit for bit the frames of one device -- whole frame, task mask, passes, RGB + spectral pipelines."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in this process (gpurun --gpus 2)")
    import scenes
    api = reference.ref_api()
    from raysect.optical.observer import RGBPipeline2D
    from source_b200.engine import Device, DeviceGroup
    from source_b200.plugin import CudaRenderEngine, WholeFrameSampler2D
    second = Device(1)
    kw = dict(pixels=(70, 52), bins=12, spectral_rays=2, samples=4)
    mask = np.ones((70, 52), dtype=bool)
    mask[20:31, 5:40] = False
    frames = []
    for devices, sampler in ((None, "whole"), ([device, second], "whole"), (None, "mask"), ([device, second], "mask")):
        w = scenes.cornell_box(api)
        cam, pipe = scenes.cornell_camera(api, w, sensitivity=1.3, **kw)
        rgb = RGBPipeline2D(display_progress=False)
        cam.pipelines = [pipe, rgb]
        cam.frame_sampler = WholeFrameSampler2D() if sampler == "whole" else api.FullFrameSampler2D(mask)
        eng = CudaRenderEngine(seed=2718, rng="mt", passes=2, **(dict(devices=devices) if devices else dict(device=device)))
        cam.render_engine = eng
        cam.observe()
        assert isinstance(eng._accel, DeviceGroup) == bool(devices)
        frames.append([np.array(a) for f in (pipe.frame, rgb.xyz_frame) for a in (f.mean, f.variance, f.samples)] + [eng.ray_count])
    for one, two in ((frames[0], frames[1]), (frames[2], frames[3])):
        for a, b in zip(one[:-1], two[:-1]):
            np.testing.assert_array_equal(a, b)
        assert one[-1] == two[-1] > 0
    assert frames[0][2].min() == 4 and frames[2][2][~mask].max() == 0 and frames[2][5][mask].min() == 4
    second.close()


def test_mono_pipelines_on_device_against_live_reference(device, reference):
    """PowerPipeline2D (filtered) + RadiancePipeline2D + RGBPipeline2D + BayerPipeline2D + a spectral pipeline from ONE device
    render (8 projection channels) vs the reference's serial render feeding the same pipelines: 1e-6 relative, no divergent
    pixel."""
    import scenes
    api = reference.ref_api()
    from raysect.optical.observer import BayerPipeline2D, PowerPipeline2D, RadiancePipeline2D, RGBPipeline2D
    from source_b200.plugin import CudaRenderEngine, WholeFrameSampler2D
    kw = dict(pixels=(20, 16), bins=12, spectral_rays=2)
    filt = api.InterpolatedSF([300, 450, 600, 800], [0.1, 1.0, 0.6, 0.2])
    rgb_filters = [api.InterpolatedSF([300, 550, 800], v) for v in ([0.0, 0.2, 1.0], [0.1, 1.0, 0.1], [1.0, 0.3, 0.0])]

    def camera(samples, accumulate):
        cam, pipe = scenes.cornell_camera(api, scenes.cornell_box(api), samples=samples, sensitivity=2.2, **kw)
        extra = [PowerPipeline2D(filter=filt, display_progress=False, accumulate=accumulate),
                 RadiancePipeline2D(display_progress=False, accumulate=accumulate), RGBPipeline2D(display_progress=False, accumulate=accumulate),
                 BayerPipeline2D(*rgb_filters, display_progress=False, accumulate=accumulate)]
        cam.pipelines = [extra[0], pipe, extra[1], extra[2], extra[3]]
        return cam, pipe, extra
    cam, pipe, extra = camera(3, True)
    reference.oracle_render(cam, pipe, 777, passes=2)
    cam2, pipe2, extra2 = camera(6, False)
    cam2.frame_sampler = WholeFrameSampler2D()
    cam2.render_engine = CudaRenderEngine(seed=777, rng="mt", device=device, passes=2)
    cam2.observe()
    for ref_p, our_p in zip([pipe] + extra, [pipe2] + extra2):
        fr, fo = (getattr(q, "xyz_frame", None) or q.frame for q in (ref_p, our_p))
        shape3 = lambda a: np.array(a).reshape(20, 16, -1)     # noqa: E731

        class F:  # noqa: E701
            mean, variance, samples = shape3(fo.mean), shape3(fo.variance), shape3(fo.samples)
        assert shape3(fr.mean).max() > 0
        parity.compare_frame(F, dict(mean=shape3(fr.mean), variance=shape3(fr.variance), samples=shape3(fr.samples)), exact=False,
                             rtol=1e-6, max_divergent_fraction=0.0)


def test_vector_camera_on_device_against_live_reference(device, reference):
    """VectorCamera (per-pixel origins / directions, slerp sub-sampling off the edge, no draws at the edge) with an RGB and a
    spectral pipeline on the B200 vs the reference's serial render: 1e-6 relative, no divergent pixel."""
    import scenes
    api = reference.ref_api()
    from raysect.core import Point3D, Vector3D
    from raysect.optical.observer import RGBPipeline2D, VectorCamera
    from source_b200.plugin import CudaRenderEngine
    nx, ny = 14, 11
    origins = np.empty((nx, ny), dtype=object)
    directions = np.empty((nx, ny), dtype=object)
    for x in range(nx):
        for y in range(ny):
            origins[x, y] = Point3D(0.02 * (x - nx / 2), 0.02 * (y - ny / 2), 0.0)
            directions[x, y] = Vector3D(-0.9 * (x + 0.5 - nx / 2) / nx, -0.9 * (y + 0.5 - ny / 2) / ny, 1.0 + 0.01 * x * y)

    def camera(world):
        pipe, rgb = api.SpectralPowerPipeline2D(), RGBPipeline2D(display_progress=False)
        cam = VectorCamera(origins, directions, frame_sampler=api.FullFrameSampler2D(), pipelines=[pipe, rgb], sensitivity=1.4,
                           parent=world, transform=api.translate(0.05, 0.0, -3.1) * api.rotate(3, -2, 1))
        cam.spectral_rays = 1
        cam.spectral_bins = 12
        cam.spectral_rays = 2
        cam.pixel_samples = 4
        cam.quiet = True
        return cam, pipe, rgb
    cam, pipe, rgb = camera(scenes.cornell_box(api))
    m_ref, v_ref, n_ref = reference.oracle_render(cam, pipe, 3030)
    x_ref = dict(mean=np.array(rgb.xyz_frame.mean), variance=np.array(rgb.xyz_frame.variance), samples=np.array(rgb.xyz_frame.samples))
    cam2, pipe2, rgb2 = camera(scenes.cornell_box(api))
    cam2.render_engine = CudaRenderEngine(seed=3030, rng="mt", device=device)
    cam2.observe()

    class F:  # noqa: E701
        mean, variance, samples = np.array(pipe2.frame.mean), np.array(pipe2.frame.variance), np.array(pipe2.frame.samples)

    class X:  # noqa: E701
        mean, variance, samples = np.array(rgb2.xyz_frame.mean), np.array(rgb2.xyz_frame.variance), np.array(rgb2.xyz_frame.samples)
    assert m_ref[0].max() > 0 and m_ref[1:-1, 1:-1].max() > 0      # edge and interior pixels both see light
    parity.compare_frame(F, dict(mean=m_ref, variance=v_ref, samples=n_ref), exact=False, rtol=1e-6, max_divergent_fraction=0.0)
    parity.compare_frame(X, x_ref, exact=False, rtol=1e-6, max_divergent_fraction=0.0)


def test_torus_on_device_against_live_reference(device, reference):
    """Torus on the B200 vs the reference's KDTree + serial render.  The quartic solver goes through cbrt / acos / cos, which
    differ from glibc's in the last place on the device: primitive ids and exiting flags equal, distances and local geometry
    within 1e-9 relative on 3,000 rays.  In a rendered frame a few paths do branch differently -- a daughter ray leaving the
    glass torus from its 1e-9 offset point sees the surface it just left as a root of the quartic a hair above or below zero,
    and the last bit of cbrt / cos decides which (3 % of the pixels of this frame: measured) -- so the frame is held to 1e-6
    on at least 94 % of the pixels; the host build of the same source (glibc) reproduces the reference's frame bit for bit
    (tests/test_plugin_host.py::test_torus_matches_reference)."""
    import scenes
    api = reference.ref_api()
    from raysect.primitive import Torus
    from source_b200.plugin import CudaAccelerator, CudaRenderEngine

    def scene():
        world = api.World()
        Torus(1.0, 0.35, parent=world, transform=api.translate(0.1, -0.2, 0.3) * api.rotate(25, 40, 10),
              material=api.UniformSurfaceEmitter(api.ConstantSF(1.0), 0.8))
        Torus(0.6, 0.6, parent=world, transform=api.translate(-0.4, 0.9, 1.4) * api.rotate(-70, 15, 0), material=api.schott("N-BK7"))
        api.Sphere(0.3, parent=world, transform=api.translate(0.1, -0.2, 0.3), material=api.Lambert(api.ConstantSF(0.7)))
        return world
    rng = np.random.default_rng(11)
    n = 3000
    o = rng.uniform(-2.5, 2.5, (n, 3))
    d = rng.normal(size=(n, 3))
    md = np.where(rng.uniform(size=n) < 0.3, rng.uniform(0.2, 3.0, n), np.inf)
    world = scene()
    ref = reference.oracle_hit(world, o, d, md)
    acc = CudaAccelerator(device=device)
    world.accelerator = acc
    world.build_accelerator(force=True)
    r = acc.hit_batch(o, d, md, geometry=True)
    hit = ref["primitive"] >= 0
    assert (ref["primitive"] == 0).sum() > 150 and (ref["primitive"] == 1).sum() > 60
    np.testing.assert_array_equal(r.primitive, ref["primitive"])
    np.testing.assert_array_equal(r.exiting[hit], ref["exiting"][hit])
    np.testing.assert_allclose(r.distance[hit], ref["distance"][hit], rtol=1e-9)
    np.testing.assert_allclose(r.geometry[hit], ref["geometry"][hit], rtol=1e-9, atol=1e-12)
    kw = dict(pixels=(24, 20), samples=3, bins=8, spectral_rays=1)
    cam, pipe = scenes.cornell_camera(api, scene(), **kw)
    cam.transform = api.translate(0, 0, -3.5)
    m_ref, v_ref, n_ref = reference.oracle_render(cam, pipe, 404)
    cam2, pipe2 = scenes.cornell_camera(api, scene(), **kw)
    cam2.transform = api.translate(0, 0, -3.5)
    cam2.render_engine = CudaRenderEngine(seed=404, rng="mt", device=device)
    cam2.observe()

    class F:  # noqa: E701
        mean, variance, samples = np.array(pipe2.frame.mean), np.array(pipe2.frame.variance), np.array(pipe2.frame.samples)
    print("torus frame divergent fraction", parity.compare_frame(F, dict(mean=m_ref, variance=v_ref, samples=n_ref), exact=False,
                                                                 rtol=1e-6, max_divergent_fraction=0.06))


def test_lens_library_on_device_against_live_reference(device, reference):
    """Lenses (EncapsulatedPrimitive around CSG trees) on the B200: hits bit-exact against the reference's KDTree (only
    + - * / sqrt involved), a frame through the lenses within 1e-6 with no divergent pixel."""
    import scenes
    api = reference.ref_api()
    from raysect.primitive.lens.spherical import BiConvex, Meniscus
    from source_b200.plugin import CudaAccelerator, CudaRenderEngine

    def scene():
        world = api.World()
        glass = api.schott("N-BK7")
        BiConvex(0.8, 0.25, 1.2, 1.5, parent=world, transform=api.translate(-0.5, 0.0, 0.0) * api.rotate(10, 5, 0), material=glass)
        Meniscus(0.7, 0.12, 0.9, 1.4, parent=world, transform=api.translate(0.5, 0.1, 0.2) * api.rotate(-8, 12, 3), material=glass)
        api.Box(api.Point3D(-2, -2, 2.0), api.Point3D(2, 2, 2.1), parent=world,
                material=api.UniformSurfaceEmitter(api.InterpolatedSF([300, 550, 800], [0.3, 1.0, 0.5])))
        return world
    rng = np.random.default_rng(8)
    n = 3000
    o = np.c_[rng.uniform(-1.2, 1.2, (n, 2)), rng.uniform(-2.0, -1.0, n)]
    d = np.c_[rng.normal(scale=0.25, size=(n, 2)), np.ones(n)]
    world = scene()
    ref = reference.oracle_hit(world, o, d)
    acc = CudaAccelerator(device=device)
    world.accelerator = acc
    world.build_accelerator(force=True)
    assert (ref["primitive"] == 0).sum() > 100 and (ref["primitive"] == 1).sum() > 100
    parity.check_hits(acc.hit_batch(o, d, geometry=True), ref)
    kw = dict(pixels=(20, 16), samples=3, bins=8, spectral_rays=2)
    cam, pipe = scenes.cornell_camera(api, scene(), **kw)
    cam.transform = api.translate(0, 0, -2.5)
    m_ref, v_ref, n_ref = reference.oracle_render(cam, pipe, 606)
    cam2, pipe2 = scenes.cornell_camera(api, scene(), **kw)
    cam2.transform = api.translate(0, 0, -2.5)
    cam2.render_engine = CudaRenderEngine(seed=606, rng="mt", device=device)
    cam2.observe()

    class F:  # noqa: E701
        mean, variance, samples = np.array(pipe2.frame.mean), np.array(pipe2.frame.variance), np.array(pipe2.frame.samples)
    parity.compare_frame(F, dict(mean=m_ref, variance=v_ref, samples=n_ref), exact=False, rtol=1e-6, max_divergent_fraction=0.0)


@pytest.mark.parametrize("kind", ["pixel", "sightline"])
def test_pixel_observer_on_device_against_live_reference(device, reference, kind):
    """Pixel, a 0-D observer (tasks as the pixels of an (n_tasks, 1) frame) with a spectral and two mono 0-D pipelines on the
    B200 vs the reference driven by an engine that re-seeds per (slice, task): accumulated statistics within 1e-6 relative."""
    import scenes
    api = reference.ref_api()
    from raysect.core.math.random import seed as reseed
    from raysect.core.workflow import RenderEngine
    from raysect.optical.observer import Pixel, PowerPipeline0D, RadiancePipeline0D, SightLine, SpectralPowerPipeline0D
    from source_b200.plugin import CudaRenderEngine
    filt = api.InterpolatedSF([300, 450, 600, 800], [0.1, 1.0, 0.6, 0.2])

    def observer(world):
        pipes = [SpectralPowerPipeline0D(display_progress=False), PowerPipeline0D(filter=filt), RadiancePipeline0D()]
        kw = dict(parent=world, transform=api.translate(0.1, -0.1, -0.9) * api.rotate(6, -4, 2), pixel_samples=60, samples_per_task=10,
                  spectral_bins=12, spectral_rays=2, quiet=True)
        px = Pixel(pipes, x_width=0.3, y_width=0.2, **kw) if kind == "pixel" else SightLine(pipelines=pipes, sensitivity=2.5, **kw)
        px.ray_extinction_min_depth = 2
        px.ray_extinction_prob = 0.1
        return px, pipes

    class Reseeding(RenderEngine):
        def run(self, tasks, render, update, render_args=(), render_kwargs={}, update_args=(), update_kwargs={}):
            slice_id = render_args[0]
            for k, task in enumerate(tasks):
                reseed(5150 + slice_id * len(tasks) + k)
                update(render(task, *render_args, **render_kwargs), *update_args, **update_kwargs)

        def worker_count(self):
            return 1
    px, pipes = observer(scenes.cornell_box(api))
    px.render_engine = Reseeding()
    px.observe()
    px2, pipes2 = observer(scenes.cornell_box(api))
    px2.render_engine = CudaRenderEngine(seed=5150, rng="mt", device=device)
    px2.observe()
    np.testing.assert_array_equal(np.array(pipes2[0].samples.samples), np.array(pipes[0].samples.samples))
    np.testing.assert_allclose(np.array(pipes2[0].samples.mean), np.array(pipes[0].samples.mean), rtol=1e-6)
    np.testing.assert_allclose(np.array(pipes2[0].samples.variance), np.array(pipes[0].samples.variance), rtol=1e-6)
    for a, b in zip(pipes2[1:], pipes[1:]):
        assert a.value.samples == b.value.samples == 60 and b.value.mean > 0
        assert abs(a.value.mean - b.value.mean) <= 1e-6 * b.value.mean and abs(a.value.variance - b.value.variance) <= 1e-6 * b.value.variance


def test_ccd_array_on_device_against_live_reference(device, reference):
    """CCDArray (a bare sensor inside the Cornell box) with its default RGB pipeline and a spectral one through
    CudaRenderEngine on the B200 vs the reference's serial render: 1e-6 relative, no divergent pixel."""
    import scenes
    api = reference.ref_api()
    from raysect.optical.observer import CCDArray, RGBPipeline2D
    from source_b200.plugin import CudaRenderEngine

    def camera(world, samples):
        pipe, rgb = api.SpectralPowerPipeline2D(), RGBPipeline2D(display_progress=False, accumulate=True)
        cam = CCDArray((18, 14), width=0.4, parent=world, transform=api.translate(0.1, -0.05, -0.9) * api.rotate(8, -5, 3),
                       pipelines=[pipe, rgb])
        cam.spectral_rays = 1
        cam.spectral_bins = 12
        cam.spectral_rays = 2
        cam.pixel_samples = samples
        cam.ray_extinction_min_depth = 2
        cam.ray_extinction_prob = 0.1
        cam.quiet = True
        return cam, pipe, rgb
    cam, pipe, rgb = camera(scenes.cornell_box(api), 3)
    m_ref, v_ref, n_ref = reference.oracle_render(cam, pipe, 8086, passes=2)
    x_ref = dict(mean=np.array(rgb.xyz_frame.mean), variance=np.array(rgb.xyz_frame.variance), samples=np.array(rgb.xyz_frame.samples))
    cam2, pipe2, rgb2 = camera(scenes.cornell_box(api), 6)
    cam2.render_engine = CudaRenderEngine(seed=8086, rng="mt", device=device, passes=2)
    cam2.observe()

    class F:  # noqa: E701
        mean, variance, samples = np.array(pipe2.frame.mean), np.array(pipe2.frame.variance), np.array(pipe2.frame.samples)

    class X:  # noqa: E701
        mean, variance, samples = np.array(rgb2.xyz_frame.mean), np.array(rgb2.xyz_frame.variance), np.array(rgb2.xyz_frame.samples)
    assert m_ref.max() > 0
    print("ccd divergent fractions", parity.compare_frame(F, dict(mean=m_ref, variance=v_ref, samples=n_ref), exact=False, rtol=1e-6,
                                                           max_divergent_fraction=0.0),
          parity.compare_frame(X, x_ref, exact=False, rtol=1e-6, max_divergent_fraction=0.0))


def test_checkerboard_emitter_on_device_against_live_reference(device, reference):
    """Checkerboard emitter (two emission spectra picked by the parity of the local hit point's cell) on the B200 vs the
    reference's serial render: 1e-6 relative, no divergent pixel."""
    import scenes
    api = reference.ref_api()
    from raysect.optical.material import Checkerboard
    from source_b200.plugin import CudaRenderEngine

    def scene():
        world = api.World()
        api.Box(api.Point3D(-1.5, -0.05, -1.5), api.Point3D(1.5, 0.0, 1.5), parent=world,
                transform=api.translate(0.1, -0.6, 0.3) * api.rotate(20, 5, -8),
                material=Checkerboard(0.23, api.ConstantSF(1.0), api.InterpolatedSF([300, 500, 800], [0.2, 1.5, 0.4]), 0.3, 2.0))
        api.Sphere(0.35, parent=world, transform=api.translate(-0.2, 0.0, 0.2), material=api.Lambert(api.ConstantSF(0.8)))
        return world
    kw = dict(pixels=(28, 22), samples=4, bins=9, spectral_rays=1)
    cam, pipe = scenes.cornell_camera(api, scene(), **kw)
    cam.transform = api.translate(0, 0.3, -2.6) * api.rotate(0, -8, 0)
    m_ref, v_ref, n_ref = reference.oracle_render(cam, pipe, 1212)
    cam2, pipe2 = scenes.cornell_camera(api, scene(), **kw)
    cam2.transform = api.translate(0, 0.3, -2.6) * api.rotate(0, -8, 0)
    cam2.render_engine = CudaRenderEngine(seed=1212, rng="mt", device=device)
    cam2.observe()

    class F:  # noqa: E701
        mean, variance, samples = np.array(pipe2.frame.mean), np.array(pipe2.frame.variance), np.array(pipe2.frame.samples)
    assert (m_ref.sum(axis=2) > 0).sum() > 100
    parity.compare_frame(F, dict(mean=m_ref, variance=v_ref, samples=n_ref), exact=False, rtol=1e-6, max_divergent_fraction=0.0)


def test_hit_sweep_device_generated_rays(device):
    """config-5 style sweep: rays generated on device; hits/sum(t) must agree with the batched API on the same rays"""
    import ctypes as C
    import torch
    import scenes
    import source_b200 as api
    from source_b200 import _cabi as cabi
    world = scenes.random_spheres(api, 3000, seed=7)
    acc = device.build(world)
    n = 200000
    hits = torch.zeros(1, dtype=torch.int64, device="cuda")
    sum_t = torch.zeros(1, dtype=torch.float64, device="cuda")
    xr = torch.zeros(1, dtype=torch.int64, device="cuda")
    origin = (C.c_double * 3)(0, 0, -4.0)
    target = (C.c_double * 3)(0, 0, 0)
    st = torch.cuda.current_stream().cuda_stream
    cabi.check(device.lib.rsb_hit_sweep_dev(device.ctx, acc.scene, C.c_void_p(st), n, 0, 12345, origin, target, 0.9, 0,
                                            C.c_void_p(hits.data_ptr()), C.c_void_p(sum_t.data_ptr()), C.c_void_p(xr.data_ptr()), 1))
    torch.cuda.synchronize()
    c = device.counters()
    assert c["rays"] == n and c["branches"] > n and c["prim_tests"] > 0
    frac = hits.item() / n
    assert 0.5 < frac <= 1.0
    # same answer when the sweep is split in two halves (index-keyed rays => order independent)
    h2 = torch.zeros(1, dtype=torch.int64, device="cuda"); s2 = torch.zeros(1, dtype=torch.float64, device="cuda"); x2 = torch.zeros(1, dtype=torch.int64, device="cuda")
    for first, cnt in ((0, n // 2), (n // 2, n - n // 2)):
        cabi.check(device.lib.rsb_hit_sweep_dev(device.ctx, acc.scene, C.c_void_p(st), cnt, first, 12345, origin, target, 0.9, 0,
                                                C.c_void_p(h2.data_ptr()), C.c_void_p(s2.data_ptr()), C.c_void_p(x2.data_ptr()), 0))
    torch.cuda.synchronize()
    assert h2.item() == hits.item() and x2.item() == xr.item()
    assert abs(s2.item() - sum_t.item()) <= 1e-9 * abs(sum_t.item())


def test_frame_combine_kernel_matches_numpy(device):
    import ctypes as C
    import torch
    from source_b200 import _cabi as cabi
    from source_b200.observer import combine_samples
    rng = np.random.default_rng(3)
    n_pix, fb, sb, off = 500, 7, 3, 2
    fm = rng.normal(size=(n_pix, fb)); fv = rng.uniform(0, 2, (n_pix, fb)); fs = rng.integers(0, 4, (n_pix, fb)).astype(np.int32)
    fm[fs == 0] = 0; fv[fs <= 1] = 0
    m = rng.normal(size=(n_pix, sb)); v = rng.uniform(0, 2, (n_pix, sb))
    t = lambda a: torch.from_numpy(a.copy()).cuda()
    gm, gv, gs, dm, dv = t(fm), t(fv), t(fs), t(m), t(v)
    st = torch.cuda.current_stream().cuda_stream
    cabi.check(device.lib.rsb_frame_combine_dev(device.ctx, C.c_void_p(st), n_pix, fb, off, sb, n_pix, None, 1, C.c_void_p(dm.data_ptr()),
                                                C.c_void_p(dv.data_ptr()), 9, C.c_void_p(gm.data_ptr()), C.c_void_p(gv.data_ptr()), C.c_void_p(gs.data_ptr())))
    torch.cuda.synchronize()
    mt, vt, nt = combine_samples(fm[:, off:off + sb], fv[:, off:off + sb], fs[:, off:off + sb], m, v, 9)
    np.testing.assert_array_equal(gs.cpu().numpy()[:, off:off + sb], nt)
    np.testing.assert_array_equal(gm.cpu().numpy()[:, off:off + sb], mt)
    np.testing.assert_array_equal(gv.cpu().numpy()[:, off:off + sb], vt)


def test_million_triangle_mesh_properties(make_backend):
    """config-4 geometry at full size: Cornell box + a 1.3M-triangle closed mesh (tree built by this package's SAH
    builder).  The oracle cannot finish this size in seconds, so: (1) the CUDA traversal must equal the host build of
    the same source on a sample of rays, bit for bit; (2) size-independent properties: every hit point lies on the mesh
    surface shell, rays started inside the mesh report exiting hits, contains() agrees with the radial shell."""
    import time
    import hostsim_api
    import scenes
    import source_b200 as api
    from source_b200.flatten import flatten_world
    verts, tris, normals = scenes.icosphere(8, radius=0.45, bumps=0.1)      # 20 * 4^8 = 1,310,720 triangles
    assert len(tris) > 1_000_000
    t0 = time.time()

    def extra(a, w):
        a.Mesh(verts, tris, normals, smoothing=True, closed=True, parent=w, transform=a.translate(0.1, -0.5, 0.1),
               material=a.Lambert(a.ConstantSF(0.7)))
    world = scenes.cornell_box(api, glass=False, extra=extra)
    flat = flatten_world(world)
    print("1.3M-triangle kd build + flatten: %.1f s" % (time.time() - t0))
    be = make_backend(flat)
    rng = np.random.default_rng(12)
    n = 60000
    o = np.tile(np.array([0.0, 0.0, -3.3]), (n, 1))
    tgt = np.array([0.1, -0.5, 0.1]) + rng.uniform(-0.6, 0.6, (n, 3))
    d = tgt - o
    d /= np.linalg.norm(d, axis=1)[:, None]
    r = be.hit_batch(o, d, geometry=True)
    mesh_id = len(world.primitives) - 1
    on_mesh = r.primitive == mesh_id
    assert on_mesh.sum() > 10000
    # hit points (mesh-local == translated world) lie in the radial shell of the bumpy sphere
    rad = np.linalg.norm(r.geometry[on_mesh, 0:3], axis=1)
    assert rad.min() > 0.45 * 0.88 and rad.max() < 0.45 * 1.12
    assert np.all(r.exiting[on_mesh] == 0)
    np.testing.assert_allclose(r.uvw[on_mesh].sum(axis=1), 1.0, atol=1e-5)
    # same source compiled for the host: bit-exact on a sub-sample
    k = 4000
    h = hostsim_api.HostScene(flat).hit_batch(o[:k], d[:k])
    for f in ("primitive", "distance", "sub", "exiting", "geometry", "uvw"):
        np.testing.assert_array_equal(getattr(r, f)[:k], getattr(h, f))
    # rays from the mesh centre leave through the surface: exiting hits at the shell radius
    c = np.array([0.1, -0.5, 0.1])
    dirs = rng.normal(size=(5000, 3))
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    r2 = be.hit_batch(np.tile(c, (5000, 1)), dirs, geometry=True)
    assert np.all(r2.primitive == mesh_id) and np.all(r2.exiting == 1)
    assert r2.distance.min() > 0.45 * 0.88 and r2.distance.max() < 0.45 * 1.12
    # contains: centre region inside, far points outside
    pts = np.r_[c + rng.uniform(-0.2, 0.2, (2000, 3)), c + np.array([0.0, 0.0, 0.7]) + rng.uniform(-0.1, 0.1, (2000, 3))]
    cnt, prims = be.contains_batch(pts, 4)
    assert np.all(cnt[:2000] == 1) and np.all(prims[:2000, 0] == mesh_id) and np.all(cnt[2000:] == 0)
    be.close()


@pytest.mark.gpu
@pytest.mark.parametrize("order", [0, 7])
@pytest.mark.parametrize("kind", ["spheres", "mesh"])
def test_hit_sweep_equals_hit_batch_on_the_same_rays(device, kind, order):
    """The sweep (rays generated on the device, incoherent or along the Morton curve, several pipeline chunks) must report
    exactly the hits rsb_hit_batch reports for the same rays restated in numpy (scenes.sweep_rays)."""
    import ctypes as C
    import torch
    import scenes
    import source_b200 as api
    from source_b200 import _cabi as cabi
    if kind == "spheres":
        world = scenes.random_spheres(api, 3000, seed=7)
        origin, target, half = (0.0, 0.0, -4.0), (0.0, 0.0, 0.0), 0.9
    else:
        verts, tris, normals = scenes.icosphere(5, radius=0.45, bumps=0.1)     # 20,480 triangles

        def extra(a, w):
            a.Mesh(verts, tris, normals, smoothing=True, closed=True, parent=w, transform=a.translate(0.1, -0.5, 0.1) * a.rotate(20, 10, 0),
                   material=a.Lambert(a.ConstantSF(0.7)))
        world = scenes.cornell_box(api, glass=False, extra=extra)
        origin, target, half = (0.0, 0.0, -3.3), (0.1, -0.5, 0.1), 0.6
    acc = device.build(world)
    n, first, seed = 150000, 1000, 2024
    hits = torch.zeros(1, dtype=torch.int64, device="cuda")
    sum_t = torch.zeros(1, dtype=torch.float64, device="cuda")
    xr = torch.zeros(1, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    cabi.check(device.lib.rsb_hit_sweep_dev(device.ctx, acc.scene, C.c_void_p(st), n, first, seed, (C.c_double * 3)(*origin),
                                            (C.c_double * 3)(*target), half, order, C.c_void_p(hits.data_ptr()), C.c_void_p(sum_t.data_ptr()),
                                            C.c_void_p(xr.data_ptr()), 0))
    torch.cuda.synchronize()
    o, d, idx = scenes.sweep_rays(seed, first, n, origin, target, half, order)
    r = acc.hit_batch(o, d)
    hit = r.primitive >= 0
    assert hit.sum() > n // 2
    assert hits.item() == int(hit.sum())
    key = (r.primitive[hit].astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + idx[hit])
    assert np.uint64(xr.item() & 0xFFFFFFFFFFFFFFFF) == np.bitwise_xor.reduce(key)
    assert abs(sum_t.item() - r.distance[hit].sum()) <= 1e-9 * r.distance[hit].sum()
    acc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["spheres", "mesh"])
def test_query_reordering_changes_no_answer(kind):
    """rsb_set_query_reorder: every pipeline pass is sorted on its coherence key, traversed as a permuted copy and written
    back at the caller's index.  Same answers, bit for bit, as the unsorted run -- ids, distances, kd nodes, geometry,
    barycentrics -- for a shared-origin batch (all 20 key bits on the direction), for rays with scattered origins,
    per-ray max_distance, NaN / zero directions, over several pipeline passes; and the sweep's checksums agree."""
    import ctypes as C
    import torch
    import scenes
    import source_b200 as api
    from source_b200 import _cabi as cabi
    from source_b200.engine import Device
    old = os.environ.get("RSB_RQ_CHUNK")
    os.environ["RSB_RQ_CHUNK"] = "40000"          # several passes per batch
    try:
        dev = Device(0)
    finally:
        if old is None:
            del os.environ["RSB_RQ_CHUNK"]
        else:
            os.environ["RSB_RQ_CHUNK"] = old
    if kind == "spheres":
        world = scenes.random_spheres(api, 3000, seed=7)
        origin, target, half = (0.0, 0.0, -4.0), (0.0, 0.0, 0.0), 0.9
    else:
        verts, tris, normals = scenes.icosphere(5, radius=0.45, bumps=0.1)

        def extra(a, w):
            a.Mesh(verts, tris, normals, smoothing=True, closed=True, parent=w, transform=a.translate(0.1, -0.5, 0.1) * a.rotate(20, 10, 0),
                   material=a.Lambert(a.ConstantSF(0.7)))
        world = scenes.cornell_box(api, glass=False, extra=extra)
        origin, target, half = (0.0, 0.0, -3.3), (0.1, -0.5, 0.1), 0.6
    acc = dev.build(world)
    n = 150000
    o1, d1, _ = scenes.sweep_rays(2024, 0, n, origin, target, half, 0)
    rng = np.random.default_rng(5)
    o2 = rng.uniform(-1.2, 1.2, (n, 3))
    d2 = rng.normal(size=(n, 3))
    d2[:50] = 0.0
    d2[50:100, 1] = np.nan
    md2 = np.where(rng.uniform(size=n) < 0.3, rng.uniform(0.0, 2.0, n), np.inf)
    answers = []
    for on in (False, True):
        dev.set_query_reorder(on)
        ra = acc.hit_batch(o1, d1, geometry=True)
        rb = acc.hit_batch(o2, d2, md2, geometry=True)
        hits = torch.zeros(1, dtype=torch.int64, device="cuda")
        sum_t = torch.zeros(1, dtype=torch.float64, device="cuda")
        xr = torch.zeros(1, dtype=torch.int64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        cabi.check(dev.lib.rsb_hit_sweep_dev(dev.ctx, acc.scene, C.c_void_p(st), n, 77, 2024, (C.c_double * 3)(*origin),
                                             (C.c_double * 3)(*target), half, 0, C.c_void_p(hits.data_ptr()), C.c_void_p(sum_t.data_ptr()),
                                             C.c_void_p(xr.data_ptr()), 0))
        torch.cuda.synchronize()
        answers.append((ra, rb, hits.item(), xr.item(), sum_t.item()))
    (a0, b0, h0, x0, s0), (a1, b1, h1, x1, s1) = answers
    assert (a0.primitive >= 0).sum() > n // 2 and (b0.primitive >= 0).sum() > n // 20
    for r0, r1 in ((a0, a1), (b0, b1)):
        for name in ("primitive", "distance", "sub", "exiting", "node", "geometry", "uvw"):
            np.testing.assert_array_equal(getattr(r0, name), getattr(r1, name), err_msg=name)
    assert h0 == h1 and x0 == x1 and abs(s0 - s1) <= 1e-9 * abs(s0)
    acc.close()
    dev.close()


@pytest.mark.gpu
def test_mesh_frame_gpu_matches_host_build(make_backend):
    """Cornell box holding a 20k-triangle mesh, rendered by the wavefront kernels (two-level traversal loop with pooled
    triangle tests in k_wf_trace) against the host build of the same source running the sequential traversal."""
    import hostsim_api
    import scenes
    import source_b200 as api
    verts, tris, normals = scenes.icosphere(5, radius=0.45, bumps=0.1)

    def extra(a, w):
        a.Mesh(verts, tris, normals, smoothing=True, closed=True, parent=w, transform=a.translate(0.1, -0.5, 0.1) * a.rotate(20, 10, 0),
               material=a.Lambert(a.ConstantSF(0.7)))
    world = scenes.cornell_box(api, glass=True, extra=extra)
    kw = dict(pixels=(40, 36), samples=6, bins=16)
    _, f_gpu = parity.observe(make_backend, world, 31, **kw)
    _, f_cpu = parity.observe(hostsim_api.HostScene, world, 31, **kw)
    g = dict(mean=f_cpu.mean, variance=f_cpu.variance, samples=f_cpu.samples)
    fr = parity.compare_frame(f_gpu, g, exact=False, rtol=1e-6, max_divergent_fraction=0.0)
    print("mesh frame: divergent pixel fraction vs host build:", fr)


def test_config5_sphere_field_at_size_and_its_sweep(device, make_backend):
    """BASELINE config 5 at size, against the REFERENCE: 10,000 spheres from the reference generator after seed(7) (the
    device's MT19937-64 reproduces the stream), the first 20,000 rays of the device sweep in both orders -- ids, t and
    world kd leaves of rsb_hit_batch bit-exact, and the sweep's own reduction (hits, xor of ids, sum of t) equal to the
    reference's answers for the rays it generates on the device."""
    import ctypes as C
    import torch
    import scenes
    from source_b200 import _cabi as cabi
    stream = device.rng_uniform(7, 40000)
    acc, g = parity.sweep10k(make_backend, stream)
    st = torch.cuda.current_stream().cuda_stream
    for tag, order in (("random", 0), ("morton", 7)):
        hits = torch.zeros(1, dtype=torch.int64, device="cuda")
        sum_t = torch.zeros(1, dtype=torch.float64, device="cuda")
        xr = torch.zeros(1, dtype=torch.int64, device="cuda")
        cabi.check(device.lib.rsb_hit_sweep_dev(device.ctx, acc.scene, C.c_void_p(st), 20000, 0, 2024, (C.c_double * 3)(*scenes.SWEEP_ORIGIN),
                                                (C.c_double * 3)(*scenes.SWEEP_TARGET), scenes.SWEEP_HALF, order, C.c_void_p(hits.data_ptr()),
                                                C.c_void_p(sum_t.data_ptr()), C.c_void_p(xr.data_ptr()), 0))
        torch.cuda.synchronize()
        prim, dist = g[tag + "_primitive"], g[tag + "_distance"]
        hit = prim >= 0
        assert hits.item() == int(hit.sum())
        key = prim[hit].astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.arange(20000, dtype=np.uint64)[hit]
        assert np.uint64(xr.item() & 0xFFFFFFFFFFFFFFFF) == np.bitwise_xor.reduce(key)
        assert abs(sum_t.item() - dist[hit].sum()) <= 1e-12 * dist[hit].sum()
    acc.close()


def test_reference_bunny_fixture(make_backend):
    """demos/resources/stanford_bunny.rsm (the reference's own mesh fixture, 144,046 triangles, tree from the file)"""
    import os
    import scenes
    if not os.path.exists(scenes.BUNNY_RSM):
        pytest.skip("demos/resources/stanford_bunny.rsm did not travel with this snapshot (oracle/_ref/resources)")
    parity.bunny_rsm(make_backend)


def test_config4_million_triangle_bunny(device, make_backend):
    """BASELINE config 4 geometry at size against the REFERENCE (golden = the reference hitting the 1,000,000-triangle mesh
    it loaded from the .rsm this package wrote), then the device sweep on the same scene against rsb_hit_batch"""
    import ctypes as C
    import os
    import torch
    import scenes
    from source_b200 import _cabi as cabi
    if not os.path.exists(scenes.BUNNY_OBJ) and not os.path.exists(os.path.join(scenes.MESH_CACHE, "bunny_1000000.rsm")):
        pytest.skip("demos/resources/stanford_bunny.obj did not travel with this snapshot (oracle/_ref/resources)")
    acc, world = parity.cornell_bunny_1m(make_backend)
    n, first, seed = 300000, 5, 99
    origin, target, half = (0.0, 0.0, -3.3), (0.1, -0.5, 0.1), 0.6
    st = torch.cuda.current_stream().cuda_stream
    for order in (0, 9):
        hits = torch.zeros(1, dtype=torch.int64, device="cuda")
        sum_t = torch.zeros(1, dtype=torch.float64, device="cuda")
        xr = torch.zeros(1, dtype=torch.int64, device="cuda")
        cabi.check(device.lib.rsb_hit_sweep_dev(device.ctx, acc.scene, C.c_void_p(st), n, first, seed, (C.c_double * 3)(*origin),
                                                (C.c_double * 3)(*target), half, order, C.c_void_p(hits.data_ptr()),
                                                C.c_void_p(sum_t.data_ptr()), C.c_void_p(xr.data_ptr()), 0))
        torch.cuda.synchronize()
        o, d, idx = scenes.sweep_rays(seed, first, n, origin, target, half, order)
        r = acc.hit_batch(o, d)
        hit = r.primitive >= 0
        assert hits.item() == int(hit.sum()) and hit.sum() > n // 2
        key = r.primitive[hit].astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + idx[hit]
        assert np.uint64(xr.item() & 0xFFFFFFFFFFFFFFFF) == np.bitwise_xor.reduce(key)
    acc.close()


def test_frame_renderer_step_host_accumulates_without_aliasing(device):
    """FrameRenderer.step_host into an ACCUMULATING pipeline: the second frame must be combined with the first
    (StatsArray3D.combine_samples), not with itself -- the pinned staging buffer the frame views is never the one the
    next device->host copy lands in."""
    import scenes
    import source_b200 as api
    from source_b200.distributed import FrameRenderer
    from source_b200.observer import combine_samples
    world = scenes.cornell_box(api)
    cam, pipe = scenes.cornell_camera(api, world, pixels=(24, 20), samples=2, bins=8)
    pipe.accumulate = True
    world._device = device
    accel = world.build_accelerator()
    r = FrameRenderer(cam, accel)
    r.step_host(seed=5)
    m1, v1, n1 = pipe.frame.mean.copy(), pipe.frame.variance.copy(), pipe.frame.samples.copy()
    m2, v2, _ = accel.render(r.cam, r.cfg, r.spectral, cam.rng_mode, 6)
    r.step_host(seed=6)
    mt, vt, nt = combine_samples(m1, v1, n1, m2, np.maximum(v2, 0.0), 2)
    np.testing.assert_array_equal(pipe.frame.samples, nt)
    np.testing.assert_array_equal(pipe.frame.mean, mt)
    np.testing.assert_array_equal(pipe.frame.variance, vt)
    assert not np.array_equal(m1, m2)
    r.step_host(seed=7)
    assert int(pipe.frame.samples.max()) == 6
