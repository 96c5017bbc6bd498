"""CPU suite, part 2: the Raysect plugin classes (CudaAccelerator / CudaRenderEngine) driven on REAL Raysect
objects from the compiled reference, with the device replaced by the host build of the device code -- this
pins the host-side logic (flattening of live Raysect scenegraphs, Intersection reconstruction, slice
handling, pipeline updates) bit-for-bit against the reference's own KDTree accelerator / SerialEngine."""
import numpy as np
import pytest

import hostsim_api
import parity
import scenes

pytestmark = pytest.mark.ref


@pytest.fixture(scope="module")
def api(reference, lib):
    hostsim_api.build()
    return reference.ref_api()


def test_accelerator_scalar_api_matches_reference_kdtree(api, reference):
    from raysect.core import Point3D, Vector3D
    from source_b200.plugin import CudaAccelerator
    world = scenes.primitive_zoo(api)
    o, d = scenes.zoo_rays(300)
    ref = reference.oracle_hit(world, o, d)
    pts = np.random.default_rng(5).uniform([-2.8, -2.6, -0.6], [2.8, 2.2, 0.8], (200, 3))
    c_ref, p_ref = reference.oracle_contains(world, pts)
    world.accelerator = CudaAccelerator(backend=hostsim_api.HostScene)     # World.accelerator seam, world.pyx:67-70
    index = {id(p): i for i, p in enumerate(world.primitives)}
    for i in range(len(o)):
        it = world.hit(api.Ray(Point3D(*o[i]), Vector3D(*d[i])))
        if ref["primitive"][i] < 0:
            assert it is None
            continue
        assert index[id(it.primitive)] == ref["primitive"][i]
        assert it.ray_distance == ref["distance"][i]
        assert bool(it.exiting) == bool(ref["exiting"][i])
        g = ref["geometry"][i]
        assert (it.hit_point.x, it.hit_point.y, it.hit_point.z) == tuple(g[0:3])
        assert (it.inside_point.x, it.inside_point.y, it.inside_point.z) == tuple(g[3:6])
        assert (it.outside_point.x, it.outside_point.y, it.outside_point.z) == tuple(g[6:9])
        assert (it.normal.x, it.normal.y, it.normal.z) == tuple(g[9:12])
        assert it.primitive_to_world is not None and it.world_to_primitive is not None
    for i in range(len(pts)):
        got = [index[id(p)] for p in world.contains(Point3D(*pts[i]))]
        assert got == list(p_ref[i, :c_ref[i]])


def test_accelerator_inside_optical_trace(api, reference):
    """The reference's own Ray.trace recursion running on top of CudaAccelerator.hit/contains must reproduce the
    reference render exactly (every World.hit / World.contains of the render goes through the plugin)."""
    from source_b200.plugin import CudaAccelerator
    world = scenes.cornell_box(api)
    cam, pipe = scenes.cornell_camera(api, world, pixels=(8, 8), samples=2, bins=4)
    m_ref, v_ref, _ = reference.oracle_render(cam, pipe, 321)
    world2 = scenes.cornell_box(api)
    world2.accelerator = CudaAccelerator(backend=hostsim_api.HostScene)
    cam2, pipe2 = scenes.cornell_camera(api, world2, pixels=(8, 8), samples=2, bins=4)
    m, v, _ = reference.oracle_render(cam2, pipe2, 321)
    np.testing.assert_array_equal(m, m_ref)
    np.testing.assert_array_equal(v, v_ref)


@pytest.mark.parametrize("bulk", [True, False])
def test_render_engine_matches_serial_reference(api, reference, bulk):
    from source_b200.plugin import CudaRenderEngine
    kw = dict(pixels=(12, 10), samples=3, bins=9, spectral_rays=3)
    world = scenes.cornell_box(api)
    cam, pipe = scenes.cornell_camera(api, world, **kw)
    m_ref, v_ref, n_ref = reference.oracle_render(cam, pipe, 555)
    world2 = scenes.cornell_box(api)
    cam2, pipe2 = scenes.cornell_camera(api, world2, **kw)
    cam2.render_engine = CudaRenderEngine(seed=555, rng="mt", bulk_update=bulk, backend=hostsim_api.HostScene)
    cam2.observe()                                          # observer.render_engine seam, observer.pyx:106
    f = pipe2.frame
    np.testing.assert_array_equal(np.array(f.samples), n_ref)
    np.testing.assert_array_equal(np.array(f.mean), m_ref)
    np.testing.assert_array_equal(np.array(f.variance), v_ref)
    # accumulate=True: a second observe() on the SAME engine draws from fresh streams (the engine's seed has moved past
    # the ones the first call used) and merges into the frame with the reference's combine rule -- the progressive
    # loop of demos/cornell_box.py:160-174.  Reference: its second pass re-seeded with the advanced base.
    m1 = np.array(pipe2.frame.mean)
    pipe.accumulate = True
    reference.oracle_render(cam, pipe, 555 + 3 * 12 * 10)
    cam2.observe()
    assert not np.array_equal(np.array(pipe2.frame.mean), m1)
    np.testing.assert_array_equal(np.array(pipe2.frame.samples), np.array(pipe.frame.samples))
    np.testing.assert_array_equal(np.array(pipe2.frame.mean), np.array(pipe.frame.mean))
    np.testing.assert_array_equal(np.array(pipe2.frame.variance), np.array(pipe.frame.variance))


def test_render_engine_concurrent_passes_match_repeated_observe(api, reference):
    """CudaRenderEngine(passes=3) rendering pixel_samples=6 in one observe() == the reference calling observe()
    three times with pixel_samples=2 into an accumulating pipeline (re-seeded per pass, slice and pixel)."""
    from source_b200.plugin import CudaRenderEngine
    kw = dict(pixels=(10, 8), bins=8, spectral_rays=2)
    world = scenes.cornell_box(api)
    cam, pipe = scenes.cornell_camera(api, world, samples=2, **kw)
    m_ref, v_ref, n_ref = reference.oracle_render(cam, pipe, 808, passes=3)
    world2 = scenes.cornell_box(api)
    cam2, pipe2 = scenes.cornell_camera(api, world2, samples=6, **kw)
    cam2.render_engine = CudaRenderEngine(seed=808, rng="mt", backend=hostsim_api.HostScene, passes=3)
    cam2.observe()
    f = pipe2.frame
    np.testing.assert_array_equal(np.array(f.samples), n_ref)
    np.testing.assert_array_equal(np.array(f.mean), m_ref)
    np.testing.assert_array_equal(np.array(f.variance), v_ref)
    cam2.pixel_samples = 7
    with pytest.raises(ValueError):
        cam2.observe()


def test_frame_buffers_are_prepared_while_the_device_renders(api, reference):
    """The helper threads that fault in / page-lock the pipeline's frame buffers during the render call (large frames
    only, so the threshold is lowered here) change nothing: an accumulating pipeline observed twice gives the frame of
    the same run without them, an empty frame is zero-filled, a frame holding samples is locked once and released when
    another frame replaces it."""
    from source_b200.plugin import CudaRenderEngine
    kw = dict(pixels=(10, 8), bins=8, spectral_rays=2)
    frames, locks = [], []

    class Backend(hostsim_api.HostScene):
        def pin(self, *arrays):
            locks.append(("lock", [a.ctypes.data for a in arrays]))
            return lambda: locks.append(("release",))

    for threshold in (1 << 60, 0):
        world = scenes.cornell_box(api)
        cam, pipe = scenes.cornell_camera(api, world, samples=3, **kw)
        pipe.accumulate = True
        eng = CudaRenderEngine(seed=31, rng="mt", backend=Backend)
        eng._READY_MIN_BYTES = threshold
        cam.render_engine = eng
        cam.observe()
        assert not locks                        # an empty frame is only touched, never locked
        cam.observe()
        f = pipe.frame
        frames.append((np.array(f.mean), np.array(f.variance), np.array(f.samples)))
        if threshold == 0:
            assert [l[0] for l in locks] == ["lock"] * 3
            pipe.accumulate = False             # initialise() now replaces the frame: the old locks are released
            cam.observe()
            assert [l[0] for l in locks].count("release") == 3
    for a, b in zip(*frames):
        np.testing.assert_array_equal(a, b)
    assert frames[0][2].max() == 6


@pytest.mark.parametrize("bulk", [True, False])
def test_whole_frame_sampler_equals_full_frame_sampler(api, reference, bulk):
    """WholeFrameSampler2D ("every pixel once" as one task) gives the frame of the stock FullFrameSampler2D, on the fast
    path and on the per-pixel update path; a task made for another frame size is refused."""
    from source_b200.plugin import CudaRenderEngine, WholeFrameSampler2D, WholeFrameTask
    kw = dict(pixels=(9, 7), bins=8, spectral_rays=2)
    frames = []
    for sampler in (None, WholeFrameSampler2D()):
        world = scenes.cornell_box(api)
        cam, pipe = scenes.cornell_camera(api, world, samples=3, **kw)
        if sampler is not None:
            cam.frame_sampler = sampler
        cam.render_engine = CudaRenderEngine(seed=77, rng="mt", backend=hostsim_api.HostScene, bulk_update=bulk)
        cam.observe()
        f = pipe.frame
        frames.append((np.array(f.mean), np.array(f.variance), np.array(f.samples)))
    for a, b in zip(*frames):
        np.testing.assert_array_equal(a, b)
    assert frames[0][2].min() == 3

    class Wrong(WholeFrameSampler2D):
        def generate_tasks(self, pixels):
            return [WholeFrameTask((4, 4))]
    cam.frame_sampler = Wrong()
    with pytest.raises(ValueError):
        cam.observe()


def _rgb_reference(api, reference, passes, seed, sensitivity, **kw):
    from raysect.optical.observer import RGBPipeline2D
    world = scenes.cornell_box(api)
    cam, pipe = scenes.cornell_camera(api, world, samples=2, sensitivity=sensitivity, **kw)
    rgb = RGBPipeline2D(display_progress=False, accumulate=passes > 1)
    cam.pipelines = [pipe, rgb]
    spectral = reference.oracle_render(cam, pipe, seed, passes=passes)
    f = rgb.xyz_frame
    return spectral, (np.array(f.mean), np.array(f.variance), np.array(f.samples))


@pytest.mark.parametrize("passes,spectral_rays", [(1, 1), (1, 2), (3, 2)])
def test_rgb_pipeline_matches_serial_reference(api, reference, passes, spectral_rays):
    """RGBPipeline2D fed by CudaRenderEngine == the reference's SerialEngine feeding it through XYZPixelProcessor
    (rgb.pyx:534-562) / update / finalise (rgb.pyx:249-290), bit for bit: next to a spectral pipeline (one render keeps
    both statistics), on its own (no spectral frame is kept), with several slices (their XYZ means are summed before the
    frame sees them) and with concurrent passes (== repeated observe() into an accumulating pipeline)."""
    from raysect.optical.observer import RGBPipeline2D
    from source_b200.plugin import CudaRenderEngine
    kw = dict(pixels=(9, 7), bins=8, spectral_rays=spectral_rays)
    (m_ref, v_ref, n_ref), xyz_ref = _rgb_reference(api, reference, passes, 4242, 2.5, **kw)
    assert xyz_ref[2].max() == 2 * passes and xyz_ref[0].max() > 0
    for with_spectral in (True, False):
        world = scenes.cornell_box(api)
        cam, pipe = scenes.cornell_camera(api, world, samples=2 * passes, sensitivity=2.5, **kw)
        rgb = RGBPipeline2D(display_progress=False)
        cam.pipelines = [pipe, rgb] if with_spectral else [rgb]
        cam.render_engine = CudaRenderEngine(seed=4242, rng="mt", backend=hostsim_api.HostScene, passes=passes)
        cam.observe()
        f = rgb.xyz_frame
        for ours, ref in zip((f.mean, f.variance, f.samples), xyz_ref):
            np.testing.assert_array_equal(np.array(ours), ref)
        if with_spectral:
            np.testing.assert_array_equal(np.array(pipe.frame.mean), m_ref)
            np.testing.assert_array_equal(np.array(pipe.frame.variance), v_ref)
            np.testing.assert_array_equal(np.array(pipe.frame.samples), n_ref)


@pytest.mark.parametrize("passes", [1, 2])
def test_mono_power_and_radiance_pipelines_match_serial_reference(api, reference, passes):
    """PowerPipeline2D (sample * filter * sensitivity * delta summed over the bins, mono/power.pyx:768-779) and
    RadiancePipeline2D (sample * filter * delta, mono/radiance.pyx:184-195), each with its own filter, next to an RGB and a
    spectral pipeline: ONE device render feeds all four (5 projection channels), every frame bit for bit the reference's."""
    from raysect.optical.observer import PowerPipeline2D, RadiancePipeline2D, RGBPipeline2D
    from source_b200.plugin import CudaRenderEngine
    kw = dict(pixels=(9, 7), bins=8, spectral_rays=2)
    filt = api.InterpolatedSF([300, 450, 600, 800], [0.1, 1.0, 0.6, 0.2])

    def camera(samples, accumulate):
        world = scenes.cornell_box(api)
        cam, pipe = scenes.cornell_camera(api, world, samples=samples, sensitivity=3.0, **kw)
        power = PowerPipeline2D(filter=filt, display_progress=False, accumulate=accumulate)
        radiance = RadiancePipeline2D(display_progress=False, accumulate=accumulate)
        rgb = RGBPipeline2D(display_progress=False, accumulate=accumulate)
        cam.pipelines = [power, pipe, radiance, rgb]
        return cam, pipe, power, radiance, rgb

    def frames(pipe, power, radiance, rgb):
        return [np.array(getattr(f, name)) for f in (pipe.frame, power.frame, radiance.frame, rgb.xyz_frame)
                for name in ("mean", "variance", "samples")]
    cam, pipe, power, radiance, rgb = camera(2, passes > 1)
    reference.oracle_render(cam, pipe, 909, passes=passes)
    ref = frames(pipe, power, radiance, rgb)
    assert ref[3].max() > 0 and ref[6].max() > 0 and ref[5].max() == 2 * passes
    cam2, pipe2, power2, radiance2, rgb2 = camera(2 * passes, False)
    cam2.render_engine = CudaRenderEngine(seed=909, rng="mt", backend=hostsim_api.HostScene, passes=passes)
    cam2.observe()
    for ours, theirs in zip(frames(pipe2, power2, radiance2, rgb2), ref):
        np.testing.assert_array_equal(ours, theirs)
    # mono pipelines alone: no spectral frame is kept, the radiance pipeline rides with the power render
    cam3, pipe3, power3, radiance3, rgb3 = camera(2 * passes, False)
    cam3.pipelines = [radiance3, power3]
    cam3.render_engine = CudaRenderEngine(seed=909, rng="mt", backend=hostsim_api.HostScene, passes=passes)
    cam3.observe()
    mono = [np.array(getattr(f, name)) for f in (power3.frame, radiance3.frame) for name in ("mean", "variance", "samples")]
    for ours, theirs in zip(mono, ref[3:9]):
        np.testing.assert_array_equal(ours, theirs)


@pytest.mark.parametrize("passes", [1, 2])
def test_bayer_pipeline_matches_serial_reference(api, reference, passes):
    """BayerPipeline2D (pipeline/bayer.pyx): three filters, the mosaic (red, green / green, blue) picks one per pixel.  The
    device keeps all three filtered totals per work item and the update takes each pixel's own: the reference's frame, bit
    for bit, next to a spectral pipeline and with a task mask."""
    from raysect.optical.observer import BayerPipeline2D
    from source_b200.plugin import CudaRenderEngine
    kw = dict(pixels=(9, 7), bins=8, spectral_rays=2)
    filters = [api.InterpolatedSF([300, 550, 800], v) for v in ([0.0, 0.2, 1.0], [0.1, 1.0, 0.1], [1.0, 0.3, 0.0])]
    mask = np.ones((9, 7), dtype=bool)
    mask[3:5, 2:6] = False

    def camera(samples, accumulate):
        cam, pipe = scenes.cornell_camera(api, scenes.cornell_box(api), samples=samples, sensitivity=1.9, **kw)
        bayer = BayerPipeline2D(*filters, display_progress=False, accumulate=accumulate)
        cam.pipelines = [bayer, pipe]
        cam.frame_sampler = api.FullFrameSampler2D(mask)
        return cam, pipe, bayer
    cam, pipe, bayer = camera(2, passes > 1)
    m_ref, v_ref, n_ref = reference.oracle_render(cam, pipe, 2001, passes=passes)
    cam2, pipe2, bayer2 = camera(2 * passes, False)
    cam2.render_engine = CudaRenderEngine(seed=2001, rng="mt", backend=hostsim_api.HostScene, passes=passes)
    cam2.observe()
    for name in ("mean", "variance", "samples"):
        np.testing.assert_array_equal(np.array(getattr(bayer2.frame, name)), np.array(getattr(bayer.frame, name)))
    np.testing.assert_array_equal(np.array(pipe2.frame.mean), m_ref)
    bm = np.array(bayer.frame.mean)
    assert bm[mask].max() > 0 and not bm[~mask].any() and np.array(bayer.frame.samples)[mask].min() == 2 * passes


def test_mono_pipeline_accumulates_and_feeds_the_mono_adaptive_sampler(api, reference):
    """An accumulating PowerPipeline2D observed twice through CudaRenderEngine == two reference passes (its frame goes up to
    the device, is merged there and comes back), and the stock MonoAdaptiveSampler2D drives the engine from that frame."""
    from raysect.optical.observer import MonoAdaptiveSampler2D, PowerPipeline2D
    from source_b200.plugin import CudaRenderEngine
    kw = dict(pixels=(9, 7), bins=8, spectral_rays=2)
    filt = api.InterpolatedSF([300, 450, 600, 800], [0.1, 1.0, 0.6, 0.2])
    cam, pipe = scenes.cornell_camera(api, scenes.cornell_box(api), samples=2, **kw)
    power = PowerPipeline2D(filter=filt, display_progress=False, accumulate=True)
    cam.pipelines = [pipe, power]
    reference.oracle_render(cam, pipe, 1357, passes=2)
    ref = [np.array(getattr(power.frame, n)) for n in ("mean", "variance", "samples")]
    cam2, pipe2 = scenes.cornell_camera(api, scenes.cornell_box(api), samples=2, **kw)
    power2 = PowerPipeline2D(filter=filt, display_progress=False, accumulate=True)
    cam2.pipelines = [power2]
    cam2.render_engine = CudaRenderEngine(seed=1357, rng="mt", backend=hostsim_api.HostScene)
    cam2.observe()
    cam2.observe()
    for ours, theirs in zip((power2.frame.mean, power2.frame.variance, power2.frame.samples), ref):
        np.testing.assert_array_equal(np.array(ours), theirs)
    cam2.frame_sampler = MonoAdaptiveSampler2D(power2, ratio=4, fraction=0.3, min_samples=5, cutoff=0.0001)
    before = np.array(power2.frame.samples)
    cam2.observe()
    grown = np.array(power2.frame.samples) - before
    assert set(np.unique(grown)) <= {0, 2} and 0 < (grown > 0).sum()


def test_rgb_pipeline_accumulates_over_observes_and_feeds_the_rgb_adaptive_sampler(api, reference):
    """An accumulating RGBPipeline2D observed twice through CudaRenderEngine == two reference passes, and the stock
    RGBAdaptiveSampler2D (sampler2d.pyx) driving the engine from that pipeline's xyz_frame picks pixel lists the engine
    renders without touching the others."""
    from raysect.optical.observer import RGBAdaptiveSampler2D, RGBPipeline2D
    from source_b200.plugin import CudaRenderEngine
    kw = dict(pixels=(9, 7), bins=8, spectral_rays=2)
    _, xyz_ref = _rgb_reference(api, reference, 2, 515, None, **kw)
    world = scenes.cornell_box(api)
    cam, pipe = scenes.cornell_camera(api, world, samples=2, **kw)
    rgb = RGBPipeline2D(display_progress=False, accumulate=True)
    cam.pipelines = [rgb]
    cam.render_engine = CudaRenderEngine(seed=515, rng="mt", backend=hostsim_api.HostScene)
    cam.observe()
    cam.observe()
    f = rgb.xyz_frame
    for ours, ref in zip((f.mean, f.variance, f.samples), xyz_ref):
        np.testing.assert_array_equal(np.array(ours), ref)
    cam.frame_sampler = RGBAdaptiveSampler2D(rgb, ratio=4, fraction=0.3, min_samples=5, cutoff=0.0001)
    before = np.array(f.samples)
    cam.observe()
    after = np.array(rgb.xyz_frame.samples)
    grown = (after - before)[:, :, 0]
    assert set(np.unique(grown)) <= {0, 2} and 0 < (grown > 0).sum()
    assert (before[:, :, 0][grown == 0] == after[:, :, 0][grown == 0]).all()


def test_rgb_pipeline_needs_the_whole_slice_device_path(api):
    from raysect.optical.observer import RGBPipeline2D
    from source_b200.plugin import CudaRenderEngine
    world = scenes.cornell_box(api)
    cam, pipe = scenes.cornell_camera(api, world, pixels=(4, 4), bins=8)
    cam.pipelines = [RGBPipeline2D(display_progress=False)]
    cam.render_engine = CudaRenderEngine(seed=3, rng="mt", backend=hostsim_api.HostScene, bulk_update=False)
    with pytest.raises(NotImplementedError):
        cam.observe()


def test_render_engine_on_real_conductor_and_unity_emitter_objects(api, reference):
    """raysect.optical.material.Conductor / UnitySurfaceEmitter objects flattened from a live Raysect scenegraph"""
    from source_b200.plugin import CudaRenderEngine
    kw = dict(pixels=(12, 10), samples=3, bins=6)
    world = scenes.metal_scene(api)
    cam, pipe = scenes.cornell_camera(api, world, **kw)
    m_ref, v_ref, n_ref = reference.oracle_render(cam, pipe, 4711)
    world2 = scenes.metal_scene(api)
    cam2, pipe2 = scenes.cornell_camera(api, world2, **kw)
    cam2.render_engine = CudaRenderEngine(seed=4711, rng="mt", backend=hostsim_api.HostScene)
    cam2.observe()
    np.testing.assert_array_equal(np.array(pipe2.frame.mean), m_ref)
    np.testing.assert_array_equal(np.array(pipe2.frame.variance), v_ref)
    assert m_ref.sum() > 0


def test_render_engine_on_real_volume_emitter_objects(api, reference):
    """raysect UniformVolumeEmitter / UnityVolumeEmitter objects (NullSurface + homogeneous emission) from a live scenegraph"""
    from source_b200.plugin import CudaRenderEngine
    kw = dict(pixels=(10, 10), samples=3, bins=5)
    world = scenes.volume_scene(api)
    cam, pipe = scenes.cornell_camera(api, world, **kw)
    m_ref, v_ref, n_ref = reference.oracle_render(cam, pipe, 99)
    world2 = scenes.volume_scene(api)
    cam2, pipe2 = scenes.cornell_camera(api, world2, **kw)
    cam2.render_engine = CudaRenderEngine(seed=99, rng="mt", backend=hostsim_api.HostScene)
    cam2.observe()
    np.testing.assert_array_equal(np.array(pipe2.frame.mean), m_ref)
    np.testing.assert_array_equal(np.array(pipe2.frame.variance), v_ref)
    assert np.all(m_ref > 0)       # the fog box contains the camera: every pixel sees emission


def test_render_engine_on_real_rough_conductor_objects(api, reference):
    from source_b200.plugin import CudaRenderEngine
    kw = dict(pixels=(10, 10), samples=3, bins=5)
    world = scenes.rough_metal_scene(api)
    cam, pipe = scenes.cornell_camera(api, world, **kw)
    m_ref, v_ref, n_ref = reference.oracle_render(cam, pipe, 12)
    world2 = scenes.rough_metal_scene(api)
    cam2, pipe2 = scenes.cornell_camera(api, world2, **kw)
    cam2.render_engine = CudaRenderEngine(seed=12, rng="mt", backend=hostsim_api.HostScene)
    cam2.observe()
    np.testing.assert_array_equal(np.array(pipe2.frame.mean), m_ref)
    np.testing.assert_array_equal(np.array(pipe2.frame.variance), v_ref)
    assert m_ref.sum() > 0


def test_adaptive_sampling_loop_with_cuda_render_engine(api, reference):
    """The reference's progressive / adaptive loop (demos/cornell_box.py:145-174 with the spectral variant of the
    sampler): Raysect's own SpectralAdaptiveSampler2D reads the pipeline's frame (mean / variance / samples, written by
    CudaRenderEngine's bulk update) and hands back a PARTIAL task list for the next observe(); the engine renders just
    those pixels and merges them with combine_samples.  Three passes must leave exactly the frame the reference leaves."""
    from raysect.optical.observer import SpectralAdaptiveSampler2D
    from source_b200.plugin import CudaRenderEngine

    def run(engine_factory):
        world = scenes.cornell_box(api)
        cam, pipe = scenes.cornell_camera(api, world, pixels=(14, 14), samples=12, bins=4)
        cam.frame_sampler = SpectralAdaptiveSampler2D(pipe, fraction=0.3, ratio=4.0, min_samples=12, cutoff=0.0)
        pipe.accumulate = True
        engine_factory(cam, pipe)
        return np.array(pipe.frame.mean), np.array(pipe.frame.variance), np.array(pipe.frame.samples)

    def reference_loop(cam, pipe):
        for p in range(3):
            reference.oracle_render(cam, pipe, 900 + 1000 * p)      # per-pixel re-seeding engine, one observe()

    def cuda_loop(cam, pipe):
        for p in range(3):
            cam.render_engine = CudaRenderEngine(seed=900 + 1000 * p, rng="mt", backend=hostsim_api.HostScene)
            cam.observe()

    m_ref, v_ref, n_ref = run(reference_loop)
    m, v, n = run(cuda_loop)
    assert n_ref.min() == 12 and n_ref.max() == 36      # the sampler really did refine a subset of the pixels
    np.testing.assert_array_equal(n, n_ref)
    np.testing.assert_array_equal(m, m_ref)
    np.testing.assert_array_equal(v, v_ref)


@pytest.mark.parametrize("method", ["weighted", "mean", "percentile", "power_percentile"])
def test_mirror_adaptive_sampler_picks_the_reference_task_set(api, reference, method):
    """source_b200.SpectralAdaptiveSampler2D (stand-alone mirror) against Raysect's sampler on the same frame statistics"""
    import source_b200 as mirror
    from raysect.optical.observer import SpectralAdaptiveSampler2D
    world = scenes.cornell_box(api)
    cam, pipe = scenes.cornell_camera(api, world, pixels=(14, 12), samples=10, bins=5)
    pipe.accumulate = True
    mask = np.ones((14, 12), dtype=bool)
    mask[:2, :] = False
    kw = dict(fraction=0.35, ratio=3.0, min_samples=10, cutoff=0.0, reduction_method=method, percentile=80.0)
    ref_sampler = SpectralAdaptiveSampler2D(pipe, mask=mask.copy(), **kw)
    cam.frame_sampler = ref_sampler
    mpipe = mirror.SpectralPowerPipeline2D()
    msampler = mirror.SpectralAdaptiveSampler2D(mpipe, mask=mask.copy(), **kw)
    assert sorted(map(tuple, msampler.generate_tasks((14, 12)))) == sorted(ref_sampler.generate_tasks((14, 12)))   # no frame yet
    sizes = []
    for p in range(3):
        reference.oracle_render(cam, pipe, 50 + 500 * p)
        f = mirror.StatsArray3D(14, 12, 5)
        f.mean[...], f.variance[...], f.samples[...] = np.array(pipe.frame.mean), np.array(pipe.frame.variance), np.array(pipe.frame.samples)
        mpipe.frame = f
        got = sorted(map(tuple, msampler.generate_tasks((14, 12))))
        want = sorted(ref_sampler.generate_tasks((14, 12)))
        assert got == want
        sizes.append(len(want))
    assert 0 < sizes[-1] < mask.sum()


def test_standalone_mirror_adaptive_loop_matches_reference_loop(api, reference):
    """the same progressive loop without Raysect installed: mirror World / PinholeCamera / SpectralAdaptiveSampler2D over
    the host build of the device code, against Raysect running its own loop"""
    import parity
    import source_b200 as mirror
    from raysect.optical.observer import SpectralAdaptiveSampler2D
    kw = dict(pixels=(14, 14), samples=12, bins=4)
    skw = dict(fraction=0.3, ratio=4.0, min_samples=12, cutoff=0.0, reduction_method="weighted")
    world = scenes.cornell_box(api)
    cam, pipe = scenes.cornell_camera(api, world, **kw)
    cam.frame_sampler = SpectralAdaptiveSampler2D(pipe, **skw)
    pipe.accumulate = True
    for p in range(3):
        reference.oracle_render(cam, pipe, 900 + 1000 * p)
    mworld = scenes.cornell_box(mirror)
    mcam, mpipe = scenes.cornell_camera(mirror, mworld, **kw)
    mcam.frame_sampler = mirror.SpectralAdaptiveSampler2D(mpipe, **skw)
    mpipe.accumulate = True
    mworld._accel = parity._Accel(hostsim_api.HostScene(parity.flatten_world(mworld)))
    mworld._rebuild = False
    for p in range(3):
        mcam.seed = 900 + 1000 * p
        mcam.observe()
    mworld._accel.close()
    n_ref = np.array(pipe.frame.samples)
    assert n_ref.min() == 12 and n_ref.max() == 36
    np.testing.assert_array_equal(mpipe.frame.samples, n_ref)
    np.testing.assert_array_equal(mpipe.frame.mean, np.array(pipe.frame.mean))
    np.testing.assert_array_equal(mpipe.frame.variance, np.array(pipe.frame.variance))


def _write_mesh_files(tmp_path):
    """the bumpy icosphere as .obj (with and without normals, with v/vt/vn tokens and comments), ascii and binary .stl"""
    import struct
    verts, tris, normals = scenes.icosphere(2, radius=0.4, bumps=0.15)
    obj_n, obj_p = tmp_path / "with_normals.obj", tmp_path / "plain.obj"
    with open(obj_n, "w") as f:
        f.write("# test mesh\n")
        for v in verts:
            f.write("v %.9g %.9g %.9g\n" % tuple(v))
        f.write("vt 0.5 0.5\n")
        for n in normals:
            f.write("vn %.9g %.9g %.9g\n" % tuple(1.7 * n))          # not unit length: the importer normalises
        for t in tris:
            f.write("f %d/1/%d %d/1/%d %d/1/%d\n" % (t[0] + 1, t[3] + 1, t[1] + 1, t[4] + 1, t[2] + 1, t[5] + 1))
    with open(obj_p, "w") as f:
        for v in verts:
            f.write("v %.9g %.9g %.9g\n" % tuple(v))
        for k, t in enumerate(tris):
            f.write(("f %d %d %d\n" if k % 2 else "f %d/1 %d/1 %d/1\n") % (t[0] + 1, t[1] + 1, t[2] + 1))
    stl_a, stl_b = tmp_path / "ascii.stl", tmp_path / "binary.stl"
    with open(stl_a, "w") as f:
        f.write("solid test\n")
        for t in tris:
            f.write("facet normal 0 0 1\nouter loop\n")
            for k in t[:3]:
                f.write("vertex %.9g %.9g %.9g\n" % tuple(verts[k]))
            f.write("endloop\nendfacet\n")
        f.write("endsolid test\n")
    with open(stl_b, "wb") as f:
        f.write(b"binary stl".ljust(80, b" "))
        f.write(struct.pack("<i", len(tris)))
        for t in tris:
            f.write(struct.pack("<3f", 0, 0, 1))
            for k in t[:3]:
                f.write(struct.pack("<3f", *[float(c) for c in verts[k]]))
            f.write(struct.pack("<H", 0))
    return obj_n, obj_p, stl_a, stl_b


def test_mesh_importers_match_reference_importers(api, reference, tmp_path):
    """source_b200.import_obj / import_stl against raysect.primitive.import_obj / import_stl: identical MeshData arrays,
    identical kd-tree stream (built by this package's SAH builder), identical hits through the host build"""
    import io
    import source_b200 as mirror
    from raysect.primitive import import_obj, import_stl
    from source_b200.flatten import flatten_world, rsm_kdtree_stream
    files = _write_mesh_files(tmp_path)
    for path, ref_imp, our_imp, kw in [(files[0], import_obj, mirror.import_obj, dict(scaling=1.5)),
                                        (files[1], import_obj, mirror.import_obj, dict(scaling=0.7)),
                                        (files[2], import_stl, mirror.import_stl, dict(scaling=2.0, mode="ascii")),
                                        (files[3], import_stl, mirror.import_stl, dict(scaling=2.0)),
                                        (files[2], import_stl, mirror.import_stl, dict(scaling=1.0, mode="auto"))]:
        rworld, mworld = api.World(), mirror.World()
        rm = ref_imp(str(path), parent=rworld, material=api.AbsorbingSurface(), **kw)
        mm = our_imp(str(path), parent=mworld, material=mirror.AbsorbingSurface(), **kw)
        np.testing.assert_array_equal(mm.data.vertices, np.array(rm.data.vertices))
        np.testing.assert_array_equal(mm.data.triangles, np.array(rm.data.triangles))
        if rm.data.vertex_normals is not None:
            np.testing.assert_array_equal(mm.data.vertex_normals, np.array(rm.data.vertex_normals))
        else:
            assert mm.data.vertex_normals is None
        assert bool(mm.data.smoothing) == bool(rm.data.smoothing)
        buf = io.BytesIO()
        rm.data.save(buf)
        blob = buf.getvalue()
        assert bytes(mm.data.kdtree_stream) == blob[rsm_kdtree_stream(blob):]
        # .rsm writer: the whole file byte for byte, and the mirror reads back what Raysect wrote
        out = io.BytesIO()
        mm.save(out)
        assert out.getvalue() == blob
        back = mirror.Mesh.from_file(io.BytesIO(blob), parent=None)
        np.testing.assert_array_equal(back.data.triangles, mm.data.triangles)
        assert bytes(back.data.kdtree_stream) == bytes(mm.data.kdtree_stream)
        rng = np.random.default_rng(1)
        o = np.c_[rng.uniform(-0.5, 0.5, 300), rng.uniform(-0.5, 0.5, 300), np.full(300, -3.0)]
        d = np.c_[rng.uniform(-0.05, 0.05, 300), rng.uniform(-0.05, 0.05, 300), np.ones(300)]
        ref = reference.oracle_hit(rworld, o, d)
        be = hostsim_api.HostScene(flatten_world(mworld))
        r = be.hit_batch(o, d, geometry=True)
        be.close()
        np.testing.assert_array_equal(r.primitive, ref["primitive"])
        np.testing.assert_array_equal(r.distance, ref["distance"])
        np.testing.assert_array_equal(r.sub[ref["triangle"] >= 0], ref["triangle"][ref["triangle"] >= 0])
        assert (ref["primitive"] >= 0).sum() > 60


def test_ply_importer_matches_reference(api, reference, tmp_path):
    """source_b200.import_ply on the files the reference's own export_ply writes, binary little-endian and ascii.  The
    reference's reader refuses its writer's header (`vertex_index` written, ply.py:116, `vertex_indices` expected, :115 / :175):
    with that one word patched its binary reader gives the mirror's arrays and kd-tree stream exactly; the ascii result is held
    to the arrays written.  Plus a big-endian file with extra vertex properties and uint indices, and a quad that must be
    refused."""
    import io
    import struct
    import source_b200 as mirror
    from raysect.primitive import export_ply, import_ply
    from source_b200.flatten import rsm_kdtree_stream
    verts, tris, _ = scenes.icosphere(2, radius=0.4, bumps=0.1)
    tris = np.ascontiguousarray(np.asarray(tris)[:, :3])
    src = api.Mesh(verts, tris, smoothing=False, closed=True)
    for mode in ("binary", "ascii"):
        path = str(tmp_path / ("sphere_%s.ply" % mode))
        export_ply(src, path, mode=mode)
        mm = mirror.import_ply(path, scaling=1.5, parent=mirror.World(), material=mirror.AbsorbingSurface())
        assert mm.data.triangles.shape == (len(tris), 3) and not mm.data.smoothing
        np.testing.assert_array_equal(mm.data.triangles, np.asarray(tris, dtype=np.int32))
        if mode == "binary":
            patched = str(tmp_path / "sphere_binary_patched.ply")
            with open(path, "rb") as f, open(patched, "wb") as g:
                g.write(f.read().replace(b"int vertex_index\n", b"int vertex_indices\n", 1))
            rm = import_ply(patched, scaling=1.5, mode="binary", parent=api.World(), material=api.AbsorbingSurface())
            mm = mirror.import_ply(patched, scaling=1.5, parent=mirror.World(), material=mirror.AbsorbingSurface())
            np.testing.assert_array_equal(mm.data.vertices, np.array(rm.data.vertices))
            buf = io.BytesIO()
            rm.data.save(buf)
            blob = buf.getvalue()
            assert bytes(mm.data.kdtree_stream) == blob[rsm_kdtree_stream(blob):]
        else:
            np.testing.assert_allclose(mm.data.vertices, (np.asarray(verts, dtype=np.float32).astype(np.float64) * 1.5).astype(np.float32),
                                       rtol=2e-6)
    big = tmp_path / "big_endian.ply"
    with open(big, "wb") as f:
        f.write(b"ply\nformat binary_big_endian 1.0\ncomment extra properties\nelement vertex 3\nproperty double x\nproperty double y\n"
                b"property double z\nproperty uchar red\nelement face 1\nproperty list uchar uint vertex_indices\nend_header\n")
        for v in ((0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (0.0, 1.0, 0.5)):
            f.write(struct.pack(">dddB", *v, 200))
        f.write(struct.pack(">BIII", 3, 0, 1, 2))
    mm = mirror.import_ply(str(big), scaling=2.0)
    np.testing.assert_array_equal(mm.data.vertices, np.array([[0, 0, 0], [2, 0, 0], [0, 2, 1]], dtype=np.float32))
    np.testing.assert_array_equal(mm.data.triangles, [[0, 1, 2]])
    quad = tmp_path / "quad.ply"
    quad.write_text("ply\nformat ascii 1.0\nelement vertex 4\nproperty float x\nproperty float y\nproperty float z\nelement face 1\n"
                    "property list uchar int vertex_indices\nend_header\n0 0 0\n1 0 0\n1 1 0\n0 1 0\n4 0 1 2 3\n")
    with pytest.raises(ValueError):
        mirror.import_ply(str(quad))
    with pytest.raises(ValueError):
        mirror.import_ply(str(quad), mode="nonsense")


def test_vtk_importer_matches_reference(api, reference, tmp_path):
    """source_b200.import_vtk against raysect.primitive.import_vtk on a legacy ASCII unstructured grid of triangles: identical
    arrays, kd-tree stream and mesh name; binary mode and non-triangular cells are refused like the reference refuses them."""
    import io
    import source_b200 as mirror
    from raysect.primitive import import_vtk
    from source_b200.flatten import rsm_kdtree_stream
    verts, tris, _ = scenes.icosphere(2, radius=0.4, bumps=0.1)
    tris = np.asarray(tris)[:, :3]
    path = tmp_path / "sphere.vtk"
    with open(path, "w") as f:
        f.write("# vtk DataFile Version 2.0\nbumpy sphere\nASCII\nDATASET UNSTRUCTURED_GRID\nPOINTS %d float\n" % len(verts))
        for v in verts:
            f.write("%r %r %r\n" % (float(v[0]), float(v[1]), float(v[2])))
        f.write("CELLS %d %d\n" % (len(tris), 4 * len(tris)))
        for tri in tris:
            f.write("3 %d %d %d\n" % tuple(tri))
        f.write("CELL_TYPES %d\n" % len(tris) + "5\n" * len(tris))
    rm = import_vtk(str(path), scaling=0.8, parent=api.World(), material=api.AbsorbingSurface())
    mm = mirror.import_vtk(str(path), scaling=0.8, parent=mirror.World(), material=mirror.AbsorbingSurface())
    np.testing.assert_array_equal(mm.data.vertices, np.array(rm.data.vertices))
    np.testing.assert_array_equal(mm.data.triangles, np.array(rm.data.triangles))
    assert mm.name == rm.name == "bumpy sphere" and not mm.data.smoothing
    buf = io.BytesIO()
    rm.data.save(buf)
    blob = buf.getvalue()
    assert bytes(mm.data.kdtree_stream) == blob[rsm_kdtree_stream(blob):]
    with pytest.raises(NotImplementedError):
        mirror.import_vtk(str(path), mode="binary")
    bad = tmp_path / "quad.vtk"
    bad.write_text("# vtk DataFile Version 2.0\nq\nASCII\nDATASET UNSTRUCTURED_GRID\nPOINTS 3 float\n0 0 0\n1 0 0\n0 1 0\nCELLS 1 4\n"
                   "3 0 1 2\nCELL_TYPES 1\n9\n")
    with pytest.raises(ValueError):
        mirror.import_vtk(str(bad))


def test_render_engine_with_real_orthographic_camera(api, reference):
    from source_b200.plugin import CudaRenderEngine
    world = scenes.cornell_box(api)
    cam, pipe = scenes.orthographic_camera(api, world, pixels=(10, 8), samples=2, bins=4)
    m_ref, v_ref, n_ref = reference.oracle_render(cam, pipe, 15)
    world2 = scenes.cornell_box(api)
    cam2, pipe2 = scenes.orthographic_camera(api, world2, pixels=(10, 8), samples=2, bins=4)
    cam2.render_engine = CudaRenderEngine(seed=15, rng="mt", backend=hostsim_api.HostScene)
    cam2.observe()
    np.testing.assert_array_equal(np.array(pipe2.frame.mean), m_ref)
    np.testing.assert_array_equal(np.array(pipe2.frame.variance), v_ref)
    assert m_ref.sum() > 0


def _ccd_camera(api, world, pipelines, samples, bins=8, spectral_rays=2, pixels=(9, 7)):
    from raysect.optical.observer import CCDArray
    cam = CCDArray(pixels, width=0.4, parent=world, transform=api.translate(0.1, -0.05, -0.9) * api.rotate(8, -5, 3), pipelines=pipelines)
    cam.spectral_rays = 1
    cam.spectral_bins = bins
    cam.spectral_rays = spectral_rays
    cam.pixel_samples = samples
    cam.ray_extinction_min_depth = 2
    cam.ray_extinction_prob = 0.1
    cam.quiet = True
    return cam


@pytest.mark.parametrize("passes", [1, 2])
def test_ccd_array_matches_serial_reference(api, reference, passes):
    """CCDArray (imaging/ccd.pyx): a bare sensor inside the Cornell box -- every sample leaves a random point of its pixel
    in a cosine-weighted direction, the pixel task drawing all its points before all its directions -- through
    CudaRenderEngine == the reference's SerialEngine, bit for bit, for the spectral and the (default) RGB pipeline."""
    from raysect.optical.observer import RGBPipeline2D
    from source_b200.plugin import CudaRenderEngine
    world = scenes.cornell_box(api)
    pipe, rgb = api.SpectralPowerPipeline2D(), RGBPipeline2D(display_progress=False, accumulate=passes > 1)
    cam = _ccd_camera(api, world, [pipe, rgb], samples=3)
    m_ref, v_ref, n_ref = reference.oracle_render(cam, pipe, 616, passes=passes)
    assert m_ref.max() > 0 and n_ref.min() == 3 * passes
    world2 = scenes.cornell_box(api)
    pipe2, rgb2 = api.SpectralPowerPipeline2D(), RGBPipeline2D(display_progress=False)
    cam2 = _ccd_camera(api, world2, [pipe2, rgb2], samples=3 * passes)
    cam2.render_engine = CudaRenderEngine(seed=616, rng="mt", backend=hostsim_api.HostScene, passes=passes)
    cam2.observe()
    np.testing.assert_array_equal(np.array(pipe2.frame.samples), n_ref)
    np.testing.assert_array_equal(np.array(pipe2.frame.mean), m_ref)
    np.testing.assert_array_equal(np.array(pipe2.frame.variance), v_ref)
    for name in ("mean", "variance", "samples"):
        np.testing.assert_array_equal(np.array(getattr(rgb2.xyz_frame, name)), np.array(getattr(rgb.xyz_frame, name)))


def test_checkerboard_emitter_matches_serial_reference(api, reference):
    """Checkerboard (emitter/checkerboard.pyx): the demos' favourite light -- squares of two emission spectra picked by
    the parity of the local hit point's cell.  A rotated, translated checkerboard slab lighting a Lambert sphere and seen
    directly, squares 0.23 m wide so that many cells (and negative coordinates) are hit: bit-exact frame."""
    from raysect.optical.material import Checkerboard
    from source_b200.plugin import CudaRenderEngine

    def scene():
        world = api.World()
        api.Box(api.Point3D(-1.5, -0.05, -1.5), api.Point3D(1.5, 0.0, 1.5), parent=world,
                transform=api.translate(0.1, -0.6, 0.3) * api.rotate(20, 5, -8),
                material=Checkerboard(0.23, api.ConstantSF(1.0), api.InterpolatedSF([300, 500, 800], [0.2, 1.5, 0.4]), 0.3, 2.0))
        api.Sphere(0.35, parent=world, transform=api.translate(-0.2, 0.0, 0.2), material=api.Lambert(api.ConstantSF(0.8)))
        return world
    kw = dict(pixels=(14, 11), samples=3, bins=9, spectral_rays=1)
    w1 = scene()
    cam, pipe = scenes.cornell_camera(api, w1, **kw)
    cam.transform = api.translate(0, 0.3, -2.6) * api.rotate(0, -8, 0)
    m_ref, v_ref, n_ref = reference.oracle_render(cam, pipe, 1212)
    w2 = scene()
    cam2, pipe2 = scenes.cornell_camera(api, w2, **kw)
    cam2.transform = api.translate(0, 0.3, -2.6) * api.rotate(0, -8, 0)
    cam2.render_engine = CudaRenderEngine(seed=1212, rng="mt", backend=hostsim_api.HostScene)
    cam2.observe()
    np.testing.assert_array_equal(np.array(pipe2.frame.mean), m_ref)
    np.testing.assert_array_equal(np.array(pipe2.frame.variance), v_ref)
    # both kinds of square are seen: the frame holds pixels lit at both scales
    lit = m_ref.sum(axis=2)
    assert (lit > 0).sum() > 40 and len(np.unique(np.round(m_ref[:, :, 0][m_ref[:, :, 0] > 0], 12))) > 3


def test_vector_camera_matches_serial_reference(api, reference):
    """VectorCamera (imaging/vector.pyx): per-pixel origins and directions; pixels off the edge are sub-sampled by slerping
    between the diagonal neighbours' directions with two draws per sample, edge pixels trace their own direction and draw
    nothing -- so the streams of edge and interior pixels start at different cursors.  Bit-exact frame, RGB and spectral."""
    from raysect.core import Point3D, Vector3D
    from raysect.optical.observer import RGBPipeline2D, VectorCamera
    from source_b200.plugin import CudaRenderEngine
    nx, ny = 8, 7
    origins = np.empty((nx, ny), dtype=object)
    directions = np.empty((nx, ny), dtype=object)
    for x in range(nx):
        for y in range(ny):
            origins[x, y] = Point3D(0.02 * (x - nx / 2), 0.02 * (y - ny / 2), 0.0)
            directions[x, y] = Vector3D(-0.9 * (x + 0.5 - nx / 2) / nx, -0.9 * (y + 0.5 - ny / 2) / ny, 1.0 + 0.01 * x * y)   # not unit length

    def camera(world):
        pipe, rgb = api.SpectralPowerPipeline2D(), RGBPipeline2D(display_progress=False)
        cam = VectorCamera(origins, directions, frame_sampler=api.FullFrameSampler2D(), pipelines=[pipe, rgb], sensitivity=1.4,
                           parent=world, transform=api.translate(0.05, 0.0, -3.1) * api.rotate(3, -2, 1))
        cam.spectral_rays = 1
        cam.spectral_bins = 8
        cam.spectral_rays = 2
        cam.pixel_samples = 3
        cam.quiet = True
        return cam, pipe, rgb
    cam, pipe, rgb = camera(scenes.cornell_box(api))
    m_ref, v_ref, n_ref = reference.oracle_render(cam, pipe, 3030)
    cam2, pipe2, rgb2 = camera(scenes.cornell_box(api))
    cam2.render_engine = CudaRenderEngine(seed=3030, rng="mt", backend=hostsim_api.HostScene)
    cam2.observe()
    np.testing.assert_array_equal(np.array(pipe2.frame.mean), m_ref)
    np.testing.assert_array_equal(np.array(pipe2.frame.variance), v_ref)
    np.testing.assert_array_equal(np.array(rgb2.xyz_frame.mean), np.array(rgb.xyz_frame.mean))
    assert m_ref[0].max() > 0 and m_ref[3, 3].max() > 0      # edge and interior pixels both see light


def test_torus_matches_reference(api, reference):
    """Torus (primitive/torus.pyx; quartic roots by Van der Waerden's method + one Newton step, core/math/cython/utility.pyx:
    423-733): World.hit ids / distances / local geometry / exiting and World.contains against the reference's own KDTree on
    rays through, past, tangent to and from inside two tori; then a rendered frame with a torus lamp and a glass torus."""
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
    o[:600] = rng.uniform(-0.2, 0.2, (600, 3)) + np.array([0.1, -0.2, 0.3])          # from inside the first torus' hole / tube region
    d[600:900] = (np.array([0.1, -0.2, 0.3]) - o[600:900]) + rng.normal(scale=0.3, size=(300, 3))
    md = np.where(rng.uniform(size=n) < 0.3, rng.uniform(0.2, 3.0, n), np.inf)
    world = scene()
    ref = reference.oracle_hit(world, o, d, md)
    pts = rng.uniform(-1.6, 1.6, (2000, 3))
    ref_cnt, ref_prims = reference.oracle_contains(world, pts)
    acc = CudaAccelerator(backend=hostsim_api.HostScene)
    world.accelerator = acc
    world.build_accelerator(force=True)
    r = acc.hit_batch(o, d, md, geometry=True)
    assert (ref["primitive"] == 0).sum() > 150 and (ref["primitive"] == 1).sum() > 60 and ref["exiting"][ref["primitive"] >= 0].any()
    parity.check_hits(r, ref)
    cnt, prims = acc.contains_batch(pts, 4)
    np.testing.assert_array_equal(cnt, ref_cnt)
    for k in range(len(pts)):
        assert sorted(prims[k, :cnt[k]].tolist()) == sorted(ref_prims[k, :ref_cnt[k]].tolist())
    assert ref_cnt.max() >= 1 and (ref_cnt > 0).sum() > 50
    # a frame
    kw = dict(pixels=(12, 10), samples=3, bins=8, spectral_rays=1)
    cam, pipe = scenes.cornell_camera(api, scene(), **kw)
    cam.transform = api.translate(0, 0, -3.5)
    m_ref, v_ref, n_ref = reference.oracle_render(cam, pipe, 404)
    cam2, pipe2 = scenes.cornell_camera(api, scene(), **kw)
    cam2.transform = api.translate(0, 0, -3.5)
    cam2.render_engine = CudaRenderEngine(seed=404, rng="mt", backend=hostsim_api.HostScene)
    cam2.observe()
    np.testing.assert_array_equal(np.array(pipe2.frame.mean), m_ref)
    np.testing.assert_array_equal(np.array(pipe2.frame.variance), v_ref)
    assert (m_ref.sum(axis=2) > 0).sum() > 20


def test_lens_library_matches_reference(api, reference):
    """The lens library (raysect/primitive/lens/spherical.pyx): EncapsulatedPrimitive subclasses hiding
    Intersect(Cylinder, Intersect(Sphere, Sphere))-style CSG trees.  To the device a lens IS its hidden primitive with that
    primitive's own matrices: hits (ids of the WRAPPERS, distances, local geometry), contains and a frame focused through a
    bi-convex and a meniscus lens, bit for bit."""
    from raysect.primitive.lens.spherical import BiConcave, BiConvex, Meniscus, PlanoConvex
    from source_b200.plugin import CudaAccelerator, CudaRenderEngine

    def scene():
        world = api.World()
        glass = api.schott("N-BK7")
        BiConvex(0.8, 0.25, 1.2, 1.5, parent=world, transform=api.translate(-0.5, 0.0, 0.0) * api.rotate(10, 5, 0), material=glass)
        Meniscus(0.7, 0.12, 0.9, 1.4, parent=world, transform=api.translate(0.5, 0.1, 0.2) * api.rotate(-8, 12, 3), material=glass)
        PlanoConvex(0.5, 0.15, 0.6, parent=world, transform=api.translate(0.0, 0.8, -0.3) * api.rotate(0, 80, 0), material=glass)
        BiConcave(0.5, 0.05, 0.9, 0.7, parent=world, transform=api.translate(0.0, -0.8, 0.4), material=glass)
        api.Box(api.Point3D(-2, -2, 2.0), api.Point3D(2, 2, 2.1), parent=world,
                material=api.UniformSurfaceEmitter(api.InterpolatedSF([300, 550, 800], [0.3, 1.0, 0.5])))
        return world
    rng = np.random.default_rng(8)
    n = 3000
    o = np.c_[rng.uniform(-1.2, 1.2, (n, 2)), rng.uniform(-2.0, -1.0, n)]
    d = np.c_[rng.normal(scale=0.25, size=(n, 2)), np.ones(n)]
    o[:500] = rng.uniform(-0.2, 0.2, (500, 3)) + np.array([-0.5, 0.0, 0.1])      # from inside the bi-convex lens
    d[:500] = rng.normal(size=(500, 3))
    world = scene()
    ref = reference.oracle_hit(world, o, d)
    pts = np.c_[rng.uniform(-1.0, 1.0, (1500, 2)), rng.uniform(-0.2, 0.6, 1500)]
    ref_cnt, ref_prims = reference.oracle_contains(world, pts)
    acc = CudaAccelerator(backend=hostsim_api.HostScene)
    world.accelerator = acc
    world.build_accelerator(force=True)
    r = acc.hit_batch(o, d, geometry=True)
    assert all((ref["primitive"] == k).sum() > 20 for k in range(4)) and ref["exiting"][ref["primitive"] == 0].any()
    parity.check_hits(r, ref)
    cnt, prims = acc.contains_batch(pts, 4)
    np.testing.assert_array_equal(cnt, ref_cnt)
    np.testing.assert_array_equal(prims[:, 0][cnt > 0], ref_prims[:, 0][ref_cnt > 0])
    assert (ref_cnt > 0).sum() > 30
    # the scalar seam: World.hit hands back an Intersection labelled with the wrapper and carrying the hidden primitive's matrices
    from raysect.core import Point3D, Vector3D
    k = int(np.argmax(ref["primitive"] == 1))
    ray = api.Ray(Point3D(*o[k]), Vector3D(*d[k]))
    it = world.hit(ray)
    world2 = scene()
    it_ref = world2.hit(api.Ray(Point3D(*o[k]), Vector3D(*d[k])))
    assert it.primitive is world.primitives[1] and it.ray_distance == it_ref.ray_distance
    for a, b in ((it.world_to_primitive, it_ref.world_to_primitive), (it.primitive_to_world, it_ref.primitive_to_world)):
        assert [[a[i, j] for j in range(4)] for i in range(4)] == [[b[i, j] for j in range(4)] for i in range(4)]
    kw = dict(pixels=(12, 10), samples=3, bins=8, spectral_rays=2)
    cam, pipe = scenes.cornell_camera(api, scene(), **kw)
    cam.transform = api.translate(0, 0, -2.5)
    m_ref, v_ref, n_ref = reference.oracle_render(cam, pipe, 606)
    cam2, pipe2 = scenes.cornell_camera(api, scene(), **kw)
    cam2.transform = api.translate(0, 0, -2.5)
    cam2.render_engine = CudaRenderEngine(seed=606, rng="mt", backend=hostsim_api.HostScene)
    cam2.observe()
    np.testing.assert_array_equal(np.array(pipe2.frame.mean), m_ref)
    np.testing.assert_array_equal(np.array(pipe2.frame.variance), v_ref)
    assert (m_ref.sum(axis=2) > 0).sum() > 60


@pytest.mark.parametrize("kind", ["pixel", "sightline"])
def test_pixel_observer_matches_serial_reference(api, reference, kind):
    """Pixel (nonimaging/pixel.pyx), a 0-D observer, with a spectral and two mono 0-D pipelines: four tasks of 5 samples, two
    spectral slices.  The reference is driven by an engine that re-seeds before every task with the seed the device's stream
    for that (slice, task) uses; every pipeline's accumulated statistics must come out bit for bit."""
    from raysect.core.math.random import seed as reseed
    from raysect.core.workflow import RenderEngine
    from raysect.optical.observer import Pixel, PowerPipeline0D, RadiancePipeline0D, SightLine, SpectralPowerPipeline0D
    from source_b200.plugin import CudaRenderEngine
    filt = api.InterpolatedSF([300, 450, 600, 800], [0.1, 1.0, 0.6, 0.2])

    def observer(world):
        pipes = [SpectralPowerPipeline0D(display_progress=False), PowerPipeline0D(filter=filt), RadiancePipeline0D()]
        kw = dict(parent=world, transform=api.translate(0.1, -0.1, -0.9) * api.rotate(6, -4, 2), pixel_samples=20, samples_per_task=5,
                  spectral_bins=8, spectral_rays=2, quiet=True)
        # (SightLine: a single line of sight, nonimaging/sightline.pyx -- the same 0-D machinery, no draws before the trace)
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
    px2.render_engine = CudaRenderEngine(seed=5150, rng="mt", backend=hostsim_api.HostScene)
    px2.observe()
    for name in ("mean", "variance", "samples"):
        np.testing.assert_array_equal(np.array(getattr(pipes2[0].samples, name)), np.array(getattr(pipes[0].samples, name)))
    for a, b in zip(pipes2[1:], pipes[1:]):
        assert (a.value.mean, a.value.variance, a.value.samples) == (b.value.mean, b.value.variance, b.value.samples)
    assert pipes[1].value.mean > 0 and pipes[2].value.mean > 0 and pipes[1].value.samples == 20
    px2.pixel_samples = 18
    with pytest.raises(NotImplementedError):
        px2.observe()


def test_mirror_ccd_array_and_vector_camera_match_reference(api, reference):
    """The stand-alone mirror's CCDArray and VectorCamera (no Raysect needed at run time) against the reference's, through the
    host build: identical spectral frames."""
    import source_b200 as mirror
    from raysect.core import Point3D, Vector3D
    from raysect.optical.observer import CCDArray, VectorCamera

    def settle(cam):
        cam.spectral_rays = 1
        cam.spectral_bins = 8
        cam.spectral_rays = 2
        cam.pixel_samples = 3
        cam.ray_extinction_min_depth = 2
        cam.ray_extinction_prob = 0.1
        return cam

    def observe_on_host_build(world, cam, seed):
        from source_b200 import _cabi as cabi
        from source_b200.flatten import flatten_world
        cam.seed, cam.rng_mode = seed, cabi.RNG_MT19937_64
        world._accel = parity._Accel(hostsim_api.HostScene(flatten_world(world)))
        world._rebuild = False
        cam.observe()
        world._accel.close()
        world._accel, world._rebuild = None, True
    # CCD
    pipe = api.SpectralPowerPipeline2D()
    cam = settle(CCDArray((9, 7), width=0.4, parent=scenes.cornell_box(api), transform=api.translate(0.1, -0.05, -0.9) * api.rotate(8, -5, 3),
                          pipelines=[pipe]))
    cam.quiet = True
    m_ref, v_ref, _ = reference.oracle_render(cam, pipe, 4711)
    mworld = scenes.cornell_box(mirror)
    mpipe = mirror.SpectralPowerPipeline2D()
    mcam = settle(mirror.CCDArray((9, 7), width=0.4, parent=mworld, transform=mirror.translate(0.1, -0.05, -0.9) * mirror.rotate(8, -5, 3),
                                  pipelines=[mpipe]))
    observe_on_host_build(mworld, mcam, 4711)
    np.testing.assert_array_equal(np.array(mpipe.frame.mean), m_ref)
    np.testing.assert_array_equal(np.array(mpipe.frame.variance), v_ref)
    # vector camera
    nx, ny = 8, 7
    o_arr, d_arr = np.zeros((nx, ny, 3)), np.zeros((nx, ny, 3))
    origins, directions = np.empty((nx, ny), dtype=object), np.empty((nx, ny), dtype=object)
    for x in range(nx):
        for y in range(ny):
            o_arr[x, y] = (0.02 * (x - nx / 2), 0.02 * (y - ny / 2), 0.0)
            d_arr[x, y] = (-0.9 * (x + 0.5 - nx / 2) / nx, -0.9 * (y + 0.5 - ny / 2) / ny, 1.0 + 0.01 * x * y)
            origins[x, y], directions[x, y] = Point3D(*o_arr[x, y]), Vector3D(*d_arr[x, y])
    pipe = api.SpectralPowerPipeline2D()
    cam = settle(VectorCamera(origins, directions, frame_sampler=api.FullFrameSampler2D(), pipelines=[pipe], sensitivity=1.4,
                              parent=scenes.cornell_box(api), transform=api.translate(0.05, 0.0, -3.1) * api.rotate(3, -2, 1)))
    cam.quiet = True
    m_ref, v_ref, _ = reference.oracle_render(cam, pipe, 3030)
    mworld = scenes.cornell_box(mirror)
    mpipe = mirror.SpectralPowerPipeline2D()
    mcam = settle(mirror.VectorCamera(o_arr, d_arr, pipelines=[mpipe], sensitivity=1.4, parent=mworld,
                                      transform=mirror.translate(0.05, 0.0, -3.1) * mirror.rotate(3, -2, 1)))
    observe_on_host_build(mworld, mcam, 3030)
    np.testing.assert_array_equal(np.array(mpipe.frame.mean), m_ref)
    np.testing.assert_array_equal(np.array(mpipe.frame.variance), v_ref)


def test_unsupported_objects_fail_loudly(api):
    from raysect.optical.observer import Pipeline2D
    from source_b200.plugin import CudaRenderEngine
    world = scenes.cornell_box(api)

    class HomeMadePipeline(Pipeline2D):
        pass
    cam = api.PinholeCamera((4, 4), parent=world, pipelines=[HomeMadePipeline()], frame_sampler=api.FullFrameSampler2D())
    cam.quiet = True
    cam.render_engine = CudaRenderEngine(backend=hostsim_api.HostScene)
    with pytest.raises(NotImplementedError):
        cam.observe()
    from raysect.primitive import Torus
    from raysect.primitive.utility import EncapsulatedPrimitive
    from source_b200.plugin import CudaAccelerator

    class Wrapped(EncapsulatedPrimitive):          # an EncapsulatedPrimitive that answers hit() itself: Python code the device cannot run
        def hit(self, ray):
            return None
    wrapped = Wrapped(api.Sphere(0.2), parent=world, material=api.AbsorbingSurface())
    world.accelerator = CudaAccelerator(backend=hostsim_api.HostScene)
    with pytest.raises(NotImplementedError):
        world.build_accelerator(force=True)
    wrapped.parent = None
    api.Union(Torus(1.0, 0.2), api.Sphere(0.5), parent=world, material=api.AbsorbingSurface())      # a torus as a CSG operand
    with pytest.raises(NotImplementedError):
        world.build_accelerator(force=True)


def test_mirror_object_model_against_live_reference_on_random_inputs(api, reference):
    """The stand-alone mirror (math3d, scenegraph bounding volumes, spectral functions) against Raysect's own classes on
    seeded random inputs -- everything the flattener feeds to the device must be bit-identical on both object models."""
    import source_b200 as mirror
    rng = np.random.default_rng(2025)

    def mat(m):
        return np.array([[m[i, j] for j in range(4)] for i in range(4)])
    for _ in range(40):
        yaw, pitch, roll = rng.uniform(-180, 180, 3)
        t = rng.uniform(-3, 3, 3)
        sc = rng.uniform(0.3, 2.0, 3)
        sh = rng.uniform(-0.4, 0.4)

        def build(a):
            m = a.AffineMatrix3D([[sc[0], sh, 0, 0], [0, sc[1], 0.5 * sh, 0], [0, 0, sc[2], 0], [0, 0, 0, 1]])
            return a.translate(*t) * a.rotate(yaw, pitch, roll) * m * a.rotate_x(roll) * a.rotate_y(yaw) * a.rotate_z(pitch)
        mr, mm = build(api), build(mirror)
        np.testing.assert_array_equal(mat(mm), mat(mr))
        np.testing.assert_array_equal(mat(mm.inverse()), mat(mr.inverse()))
        world_r, world_m = api.World(), mirror.World()
        node_r = api.Node(world_r, api.translate(*rng.uniform(-1, 1, 3)) * api.rotate(*rng.uniform(-90, 90, 3)))
        node_m = mirror.Node(world_m, mirror.AffineMatrix3D(mat(node_r.transform).tolist()))
        r, h = rng.uniform(0.1, 1.5, 2)
        lo, hi = rng.uniform(-1, 0, 3), rng.uniform(0, 1, 3)
        shapes = []
        for a, node, m in ((api, node_r, mr), (mirror, node_m, mm)):
            shapes.append([a.Sphere(r, node, m), a.Box(a.Point3D(*lo), a.Point3D(*hi), node, m), a.Cylinder(r, h, node, m),
                           a.Cone(r, h, node, m),
                           a.Intersect(a.Sphere(r, transform=a.translate(0.1, 0, 0)), a.Box(a.Point3D(*lo), a.Point3D(*hi)), node, m)])
        for pr, pm in zip(*shapes):
            np.testing.assert_array_equal(mat(pm.to_root()), mat(pr.to_root()))
            np.testing.assert_array_equal(mat(pm.to_local()), mat(pr.to_local()))
            br, bm = pr.bounding_box(), pm.bounding_box()
            assert (bm.lower.x, bm.lower.y, bm.lower.z, bm.upper.x, bm.upper.y, bm.upper.z) == \
                   (br.lower.x, br.lower.y, br.lower.z, br.upper.x, br.upper.y, br.upper.z)
            sr, sm = pr.bounding_sphere(), pm.bounding_sphere()
            assert (sm.centre.x, sm.centre.y, sm.centre.z, sm.radius) == (sr.centre.x, sr.centre.y, sr.centre.z, sr.radius)
    # spectral functions: bin averages over random ranges, inside / across / outside the tabulated range
    w = np.sort(rng.uniform(300, 800, 25))
    v = rng.uniform(0, 2, 25)
    fr, fm = api.InterpolatedSF(w, v), mirror.InterpolatedSF(w, v)
    gr, gm = api.schott("N-BK7"), mirror.schott("N-BK7")
    for _ in range(60):
        a0 = rng.uniform(250, 820)
        a1 = a0 + rng.uniform(0.01, 400)
        bins = int(rng.integers(1, 40))
        np.testing.assert_array_equal(np.asarray(fm.sample(a0, a1, bins)), np.asarray(fr.sample(a0, a1, bins)))
        assert fm.average(a0, a1) == fr.average(a0, a1)
        np.testing.assert_array_equal(np.asarray(gm.transmission.sample(a0, a1, bins)), np.asarray(gr.transmission.sample(a0, a1, bins)))
        assert gm.index.average(a0, a1) == gr.index.average(a0, a1)


def test_accelerator_on_an_empty_world(api):
    """World() without primitives: Raysect's own accelerator answers None / []; so does the plugin"""
    from raysect.core import Point3D, Vector3D
    from source_b200.plugin import CudaAccelerator
    world = api.World()
    world.accelerator = CudaAccelerator(backend=hostsim_api.HostScene)
    assert world.hit(api.Ray(Point3D(0, 0, -3), Vector3D(0, 0, 1))) is None
    assert world.contains(Point3D(0, 0, 0)) == []


def test_accelerator_with_real_parabola_objects(api, reference):
    """raysect.primitive.Parabola objects (world-level and CSG operands) flattened from a live scenegraph"""
    from raysect.core import Point3D, Vector3D
    from source_b200.plugin import CudaAccelerator
    world = scenes.parabola_scene(api)
    o, d = scenes.parabola_rays(400)
    ref = reference.oracle_hit(world, o, d)
    world.accelerator = CudaAccelerator(backend=hostsim_api.HostScene)
    index = {id(p): i for i, p in enumerate(world.primitives)}
    hits = 0
    for i in range(len(o)):
        it = world.hit(api.Ray(Point3D(*o[i]), Vector3D(*d[i])))
        if ref["primitive"][i] < 0:
            assert it is None
            continue
        hits += 1
        assert index[id(it.primitive)] == ref["primitive"][i]
        assert it.ray_distance == ref["distance"][i]
        assert bool(it.exiting) == bool(ref["exiting"][i])
        g = ref["geometry"][i]
        assert (it.hit_point.x, it.hit_point.y, it.hit_point.z) == tuple(g[0:3])
        assert (it.normal.x, it.normal.y, it.normal.z) == tuple(g[9:12])
    assert hits > 150


@pytest.mark.parametrize("bulk", [True, False])
def test_power_and_radiance_pipelines_side_by_side(api, reference, bulk):
    """one camera feeding a SpectralPowerPipeline2D AND a SpectralRadiancePipeline2D (sensitivity 2.5): the power
    frame carries the sensitivity, the radiance frame does not; both bit-exact against the reference's processors"""
    from raysect.optical.observer import SpectralRadiancePipeline2D
    from source_b200.plugin import CudaRenderEngine

    def make():
        world = scenes.cornell_box(api)
        cam, power = scenes.cornell_camera(api, world, pixels=(10, 8), samples=3, bins=6, spectral_rays=2, sensitivity=2.5)
        radiance = SpectralRadiancePipeline2D()
        cam.pipelines = [radiance, power]
        return cam, power, radiance
    cam, power, radiance = make()
    reference.oracle_render(cam, power, 77)
    cam2, power2, radiance2 = make()
    cam2.render_engine = CudaRenderEngine(seed=77, rng="mt", bulk_update=bulk, backend=hostsim_api.HostScene)
    cam2.observe()
    for a, b in ((power2, power), (radiance2, radiance)):
        np.testing.assert_array_equal(np.array(a.frame.samples), np.array(b.frame.samples))
        np.testing.assert_array_equal(np.array(a.frame.mean), np.array(b.frame.mean))
        np.testing.assert_array_equal(np.array(a.frame.variance), np.array(b.frame.variance))
    assert np.array(power.frame.mean).sum() > 2.0 * np.array(radiance.frame.mean).sum() > 0
    # the stand-alone mirror: same two pipelines
    import parity
    import source_b200 as mirror
    mworld = scenes.cornell_box(mirror)
    mcam, mpower = scenes.cornell_camera(mirror, mworld, pixels=(10, 8), samples=3, bins=6, spectral_rays=2, sensitivity=2.5)
    mrad = mirror.SpectralRadiancePipeline2D()
    mcam.pipelines = [mrad, mpower]
    mcam.seed = 77
    mworld._accel = parity._Accel(hostsim_api.HostScene(parity.flatten_world(mworld)))
    mworld._rebuild = False
    mcam.observe()
    mworld._accel.close()
    np.testing.assert_array_equal(mpower.frame.mean, np.array(power.frame.mean))
    np.testing.assert_array_equal(mrad.frame.mean, np.array(radiance.frame.mean))
    np.testing.assert_array_equal(mrad.frame.variance, np.array(radiance.frame.variance))


@pytest.mark.parametrize("bulk", [True, False])
def test_render_engine_over_several_devices_equals_one_device(api, reference, bulk):
    """CudaRenderEngine(devices=[...]): tiles dealt to three "devices" (three host-build scenes, one thread each); with
    bulk_update the first member gathers the others' rows (DeviceGroup / rsb_comm_gather_slices) and one update_frame
    merges the slice, without it the frames are summed on the host and fed to update() pixel by pixel -- either way
    bit-identical to the reference / to a single device, with passes and a partial task mask"""
    from source_b200.plugin import CudaRenderEngine
    kw = dict(pixels=(40, 36), samples=4, bins=4, spectral_rays=2)
    mask = np.ones((40, 36), dtype=bool)
    mask[5:9, :] = False
    world = scenes.cornell_box(api)
    cam, pipe = scenes.cornell_camera(api, world, **kw)
    cam.pixel_samples = 2
    cam.frame_sampler = api.FullFrameSampler2D(mask)
    m_ref, v_ref, n_ref = reference.oracle_render(cam, pipe, 31, passes=2)
    world2 = scenes.cornell_box(api)
    cam2, pipe2 = scenes.cornell_camera(api, world2, **kw)
    cam2.frame_sampler = api.FullFrameSampler2D(mask)
    engine = CudaRenderEngine(seed=31, rng="mt", backend=hostsim_api.HostScene, passes=2, devices=[0, 1, 2], bulk_update=bulk)
    assert engine.worker_count() == 3
    cam2.render_engine = engine
    cam2.observe()
    from source_b200.engine import DeviceGroup
    assert isinstance(engine._accel, DeviceGroup) == bulk
    np.testing.assert_array_equal(np.array(pipe2.frame.samples), n_ref)
    np.testing.assert_array_equal(np.array(pipe2.frame.mean), m_ref)
    np.testing.assert_array_equal(np.array(pipe2.frame.variance), v_ref)
    assert engine.ray_count > 0 and m_ref[mask].sum() > 0 and not np.array(pipe2.frame.mean)[~mask].any()


def test_device_group_feeds_rgb_and_spectral_pipelines_of_a_whole_frame(api, reference):
    """Two "devices", the whole frame in one task (WholeFrameSampler2D -> diagonal tile dealing), an RGB and a spectral
    pipeline side by side: the XYZ statistics are merged member by member, the spectral rows gathered on the first member
    -- the frames of the single-device reference run, bit for bit."""
    from raysect.optical.observer import RGBPipeline2D
    from source_b200.plugin import CudaRenderEngine, WholeFrameSampler2D
    kw = dict(pixels=(40, 36), bins=6, spectral_rays=2)
    (m_ref, v_ref, n_ref), xyz_ref = _rgb_reference(api, reference, 1, 77, 1.5, **kw)
    world = scenes.cornell_box(api)
    cam, pipe = scenes.cornell_camera(api, world, samples=2, sensitivity=1.5, **kw)
    rgb = RGBPipeline2D(display_progress=False)
    cam.pipelines = [pipe, rgb]
    cam.frame_sampler = WholeFrameSampler2D()
    cam.render_engine = CudaRenderEngine(seed=77, rng="mt", backend=hostsim_api.HostScene, devices=[0, 1])
    cam.observe()
    f = rgb.xyz_frame
    for ours, ref in zip((f.mean, f.variance, f.samples), xyz_ref):
        np.testing.assert_array_equal(np.array(ours), ref)
    np.testing.assert_array_equal(np.array(pipe.frame.mean), m_ref)
    np.testing.assert_array_equal(np.array(pipe.frame.variance), v_ref)
    np.testing.assert_array_equal(np.array(pipe.frame.samples), n_ref)


def test_subclasses_that_override_evaluated_methods_are_rejected(api):
    """no silent fallback: a user subclass is its base class to the device only if it inherits hit / evaluate_surface /
    ... unchanged"""
    from raysect.optical.material import Lambert
    from raysect.primitive import Sphere
    from source_b200.flatten import flatten_world

    class Tagged(Lambert):
        pass

    class Glowing(Lambert):
        def evaluate_surface(self, *args, **kwargs):
            return None

    class Fuzzy(Sphere):
        def hit(self, ray):
            return None

    world = api.World()
    api.Sphere(0.5, world, api.translate(0, 0, 0), Tagged(api.ConstantSF(0.5)))
    flatten_world(world)
    world = api.World()
    api.Sphere(0.5, world, api.translate(0, 0, 0), Glowing(api.ConstantSF(0.5)))
    with pytest.raises(NotImplementedError, match="overrides evaluate_surface"):
        flatten_world(world)
    world = api.World()
    Fuzzy(0.5, world, api.translate(0, 0, 0), api.AbsorbingSurface())
    with pytest.raises(NotImplementedError, match="overrides hit"):
        flatten_world(world)
