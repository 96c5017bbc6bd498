"""CPU suite, part 3: pins oracle/rs_oracle.c (the independent plain-C restatement) against the golden vectors
of the compiled reference, bit for bit, and cross-checks it against the host build of the product's device code
on inputs the goldens do not cover."""
import numpy as np
import pytest

import parity
from oracle import c_oracle


@pytest.fixture(scope="module")
def make_backend(lib):
    c_oracle.build()
    return c_oracle.OracleScene


def test_rng_known_answers():
    g = parity.golden("rng_kat")
    np.testing.assert_array_equal(c_oracle.uniform(int(g["seed"]), len(g["uniform"])), g["uniform"])
    np.testing.assert_array_equal(c_oracle.uniform(77, 700), g["seed77"])


def test_zoo_hits_contains(make_backend):
    parity.zoo(make_backend)


def test_edge_cases(make_backend):
    parity.edge(make_backend)


def test_non_rigid_transforms(make_backend):
    parity.scaled(make_backend, exact=True)


def test_configuration_extremes(make_backend):
    parity.extremes(make_backend, exact=True)


def test_parabola_primitive(make_backend):
    parity.parabola(make_backend, exact=True)


def test_sphere_field(make_backend):
    parity.spheres(make_backend)


@pytest.mark.parametrize("smoothing", [True, False])
def test_mesh(make_backend, smoothing):
    parity.mesh(make_backend, smoothing)


def test_cornell_frames_bit_exact(make_backend):
    parity.cornell(make_backend, exact=True)


def test_accumulated_passes_bit_exact(make_backend):
    parity.cornell_passes(make_backend, exact=True)


def test_conductor_and_unity_emitter_bit_exact(make_backend):
    parity.metal(make_backend, exact=True)


def test_volume_emitters_bit_exact(make_backend):
    parity.volumes(make_backend, exact=True)


def test_orthographic_camera_bit_exact(make_backend):
    parity.orthographic(make_backend, exact=True)


def test_rough_conductor_bit_exact(make_backend):
    parity.rough_metal(make_backend, exact=True)


def test_prism_csg_dispersion_bit_exact(make_backend):
    parity.prism(make_backend, exact=True)


def test_transforms_with_m33_off_one_bit_exact(make_backend):
    parity.w_matrices(make_backend, exact=True)


def test_prism_512_spectral_slices_bit_exact(make_backend):
    parity.prism_512(make_backend, exact=True)


def test_c_oracle_vs_host_build_on_fresh_inputs(make_backend):
    """two independent formulations (recursive iterators vs event lists / log replay) on inputs without goldens"""
    import hostsim_api
    import scenes
    import source_b200 as api
    from source_b200.flatten import flatten_world
    flat = flatten_world(scenes.primitive_zoo(api))
    o, d = scenes.zoo_rays(20000, seed=1234)
    a, b = make_backend(flat).hit_batch(o, d), hostsim_api.HostScene(flat).hit_batch(o, d)
    for f in ("primitive", "distance", "exiting", "geometry"):
        np.testing.assert_array_equal(getattr(a, f), getattr(b, f))
    world = scenes.prism_scene(api)
    kw = dict(pixels=(20, 20), samples=5, bins=8, spectral_rays=2, path_weight=0.6)
    _, fa = parity.observe(make_backend, world, 2024, **kw)
    _, fb = parity.observe(hostsim_api.HostScene, world, 2024, **kw)
    np.testing.assert_array_equal(fa.mean, fb.mean)
    np.testing.assert_array_equal(fa.variance, fb.variance)
