"""GPU suite: the CUDA kernels, called through the C ABI, against the golden vectors of the compiled
reference.  ids, kd trees, t and local-space geometry are bit-exact; rendered frames are held to the
north-star tolerance (1e-6 relative) with the path-divergence rate reported."""
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


def test_sphere_field(make_backend):
    parity.spheres(make_backend)


@pytest.mark.parametrize("smoothing", [True, False])
def test_mesh(make_backend, smoothing):
    parity.mesh(make_backend, smoothing)


def test_cornell_frames(make_backend):
    fr = parity.cornell(make_backend, exact=False, rtol=1e-6, max_divergent_fraction=0.02)
    print("divergent pixel fractions:", fr)


def test_prism_csg_dispersion(make_backend):
    fr = parity.prism(make_backend, exact=False, rtol=1e-6, max_divergent_fraction=0.02)
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
    fr = parity.compare_frame(f_gpu, g, exact=False, rtol=1e-6, max_divergent_fraction=0.02)
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
