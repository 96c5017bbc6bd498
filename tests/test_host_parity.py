"""CPU suite, part 1: the arithmetic of the device headers (compiled for the host, tests/hostsim) against
the golden vectors of the compiled reference -- bit for bit, including whole rendered frames."""
import numpy as np
import pytest

import hostsim_api
import parity


@pytest.fixture(scope="module")
def make_backend(lib):
    hostsim_api.build()
    return hostsim_api.HostScene


def test_rng_known_answers():
    g = parity.golden("rng_kat")
    # the reference's own vector, raysect/core/math/tests/test_random.py:37-253
    np.testing.assert_array_equal(hostsim_api.rng_uniform(int(g["seed"]), len(g["uniform"])), g["uniform"])
    np.testing.assert_array_equal(hostsim_api.rng_uniform(77, 700), g["seed77"])   # crosses two 312-word refills
    assert g["uniform"][0] == 0.8114659955555504


def test_seed_through_the_seed_independent_table_is_init_by_array64(lib):
    """seed(d) keys init_by_array64 with (0, ..., 0, d): 623 of its 935 steps do not depend on d.  The 312-step form the
    device seeds its streams with must leave exactly the state of the restated reference algorithm."""
    hostsim_api.build()
    for d in (1, 2, 77, 1234567890, 2**63 + 12345, 2**64 - 1, 999 + 5 * 1024 * 1024):
        np.testing.assert_array_equal(hostsim_api.mt_state(d, 1), hostsim_api.mt_state(d, 0))


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


def test_prism_512_spectral_slices_bit_exact(make_backend):
    parity.prism_512(make_backend, exact=True)


def test_transforms_with_m33_off_one_bit_exact(make_backend):
    parity.w_matrices(make_backend, exact=True)


def test_philox_mode_statistics(make_backend):
    """The counter-based stream must estimate the same radiance: compare frame-integrated power of a
    Philox render with the MT19937 golden within 5 standard errors of the golden's own variance."""
    import scenes
    import source_b200 as api
    from source_b200 import _cabi as cabi
    g = parity.golden("cornell_32x32_s4_b15")
    world = scenes.cornell_box(api)
    cam, frame = parity.observe(make_backend, world, 5, rng_mode=cabi.RNG_PHILOX, pixels=(32, 32), samples=4, bins=15)
    total, total_ref = frame.mean.sum(), g["mean"].sum()
    sigma = np.sqrt((g["variance"] / 4).sum())
    assert abs(total - total_ref) < 5 * np.sqrt(2) * sigma
    assert not np.array_equal(frame.mean, g["mean"])


class _Mode:
    """HostScene factory that evaluates World.hit through one of the new walk forms (rsb_trav.h)"""

    def __init__(self, mode):
        self.mode = mode

    def __call__(self, flat):
        s = hostsim_api.HostScene(flat)
        s.hit_mode = self.mode
        return s


@pytest.mark.parametrize("smoothing", [True, False])
def test_split_pipeline_mesh_hits(lib, smoothing):
    """world walk suspended in front of Mesh.hit + stand-alone mesh query over single node visits + resume with the
    per-ray memo == the reference (ids, t, triangle, u v w, instanced meshes)"""
    hostsim_api.build()
    parity.mesh(_Mode(1), smoothing)


def test_split_pipeline_on_every_primitive_type(lib):
    hostsim_api.build()
    parity.zoo(_Mode(1))
    parity.edge(_Mode(1))
    parity.scaled(_Mode(1), exact=True)


def test_single_visit_walk_on_mesh_free_scenes(lib):
    hostsim_api.build()
    parity.zoo(_Mode(2))
    parity.edge(_Mode(2))
    parity.spheres(_Mode(2))
    parity.parabola(_Mode(2), exact=True)


def test_config5_sphere_field_at_size(lib):
    """10,000 spheres from the reference generator after seed(7), the device sweep's own rays (both orders)"""
    hostsim_api.build()
    stream = hostsim_api.rng_uniform(7, 40000)
    for mode in (0, 2):
        be, _ = parity.sweep10k(_Mode(mode), stream)
        be.close()


def test_reference_bunny_fixture(lib):
    import os
    import scenes
    if not os.path.exists(scenes.BUNNY_RSM):
        pytest.skip("demos/resources/stanford_bunny.rsm did not travel with this snapshot (oracle/_ref/resources)")
    hostsim_api.build()
    for mode in (0, 1):
        parity.bunny_rsm(_Mode(mode))


def test_config4_million_triangle_bunny(lib):
    import os
    import scenes
    if not os.path.exists(scenes.BUNNY_OBJ) and not os.path.exists(os.path.join(scenes.MESH_CACHE, "bunny_1000000.rsm")):
        pytest.skip("demos/resources/stanford_bunny.obj did not travel with this snapshot (oracle/_ref/resources)")
    hostsim_api.build()
    for mode in (0, 1):
        be, _ = parity.cornell_bunny_1m(_Mode(mode))
        be.close()


def test_progressive_loop_draws_fresh_samples_every_pass(make_backend):
    """observe() called three times on ONE camera into an accumulating pipeline (demos/cornell_box.py:160-174) == the
    reference's own loop (golden cornell_16x12_s2_p3_b16_r4: three observe() calls, re-seeded per pass / slice / pixel):
    the camera's seed moves past the streams each call used, so no pass repeats the samples of another."""
    import scenes
    import source_b200 as api
    world = scenes.cornell_box(api)
    cam, pipe = scenes.cornell_camera(api, world, pixels=(16, 12), samples=2, bins=16, spectral_rays=4, path_weight=0.5)
    cam.seed = 999
    pipe.accumulate = True
    world._accel = parity._Accel(make_backend(parity.flatten_world(world)))
    world._rebuild = False
    means = []
    for _ in range(3):
        cam.observe()
        means.append(pipe.frame.mean.copy())
    world._accel.close()
    assert not np.array_equal(means[0], means[1]) and not np.array_equal(means[1], means[2])
    parity.compare_frame(pipe.frame, parity.golden("cornell_16x12_s2_p3_b16_r4"), exact=True)
