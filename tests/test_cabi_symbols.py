"""The C-ABI shared library loads on a machine without a GPU and exports exactly the entry points that
include/raysect_b200.h declares; host-only entry points work, device entry points fail loudly."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from source_b200 import _cabi as cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "raysect_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rsb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libraysect_b200.so does not export %s" % n
    assert sorted(cabi.SIGNATURES) == names, "ctypes SIGNATURES and the header disagree"
    assert lib.rsb_version() >= 100


def test_no_oracle_or_cpu_fallback_in_product():
    """product sources never reference oracle/ or the test-only host build"""
    for root, _, files in os.walk(os.path.join(ROOT, "source_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "hostsim" not in text, f


def test_device_entry_points_fail_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_uint64()
    rc = lib.rsb_context_create(0, C.byref(h))
    assert rc == cabi.ERR_CUDA
    assert b"no CPU fallback" in lib.rsb_last_error()
    with pytest.raises(cabi.RsbError):
        from source_b200.engine import Device
        Device(0)


def test_round2_entry_points_reject_bad_arguments_without_touching_a_device(lib):
    """the argument checks of the entry points added in round 2 run before any CUDA call: null handles, empty
    communicators and channel counts out of range come back as RSB_ERR_ARG with a message, on a machine without a GPU"""
    z = np.zeros(4)
    zi = np.zeros(4, dtype=np.int32)
    pd, pi = cabi.ptr(z, C.c_double), cabi.ptr(zi, C.c_int32)
    comm = C.c_uint64()
    calls = [
        lambda: lib.rsb_set_query_reorder(0, 1),
        lambda: lib.rsb_comm_create(0, None, C.byref(comm)),
        lambda: lib.rsb_comm_create(1, (C.c_uint64 * 1)(0), C.byref(comm)),
        lambda: lib.rsb_comm_gather_slices(0, 0),
        lambda: lib.rsb_slice_update_frame(0, 4, 0, 1, pd, pd, pi),
        lambda: lib.rsb_slice_update_xyz_frame(0, 1, pd, pd, pi),
        lambda: lib.rsb_slice_update_proj_frame(0, 0, 1, 1, pd, pd, pi),
        lambda: lib.rsb_slice_read(0, pd, pd),
        lambda: lib.rsb_host_pin(0, None, 16),
        lambda: lib.rsb_render_slices_xyz(0, 0, None, None, None, None, 1, 1, 0, 0, None, None, None, 1, None),
        lambda: lib.rsb_render_slices_proj(0, 0, None, None, None, None, 1, 1, 0, 0, None, 9, pi, pd, pd, 1, None),
    ]
    for k, call in enumerate(calls):
        assert call() == cabi.ERR_ARG, k
        assert lib.rsb_last_error()
    assert lib.rsb_comm_destroy(0) == 0


def test_kdtree_builder_stream_roundtrip_and_errors(lib):
    from source_b200.flatten import kdtree_build
    rng = np.random.default_rng(0)
    lo = rng.uniform(-1, 1, (200, 3))
    boxes = np.c_[lo, lo + rng.uniform(0.01, 0.2, (200, 3))]
    stream = kdtree_build(boxes, 0, 1, 80.0, 0.2)
    import struct
    max_depth, min_items, hit_cost, bonus = struct.unpack_from("<iidd", stream, 0)
    assert max_depth == int(np.ceil(8 + 1.3 * np.log(200))) and min_items == 1 and hit_cost == 80.0 and bonus == 0.2
    bounds = struct.unpack_from("<6d", stream, 24)
    np.testing.assert_array_equal(bounds[:3], boxes[:, :3].min(axis=0))
    np.testing.assert_array_equal(bounds[3:], boxes[:, 3:].max(axis=0))
    # every item appears in at least one leaf
    n_nodes = struct.unpack_from("<i", stream, 72)[0]
    off, seen = 76, set()
    for _ in range(n_nodes):
        t = struct.unpack_from("<i", stream, off)[0]
        if t == -1:
            cnt = struct.unpack_from("<i", stream, off + 4)[0]
            seen.update(struct.unpack_from("<%di" % cnt, stream, off + 8))
            off += 8 + 4 * cnt
        else:
            off += 16
    assert off == len(stream) and seen == set(range(200))
    with pytest.raises(cabi.RsbError):
        kdtree_build(boxes, 0, 1, 80.0, 1.5)      # empty_bonus outside [0, 1]: the reference raises ValueError


def test_object_model_matches_reference_golden():
    """transforms, bounding boxes/spheres and spectral resampling of the mirror classes vs the reference"""
    import parity
    import scenes
    import source_b200 as api
    g = parity.golden("object_model")
    world = scenes.primitive_zoo(api)
    rows = []
    for p in world.primitives:
        b, s = p.bounding_box(), p.bounding_sphere()
        rows.append([p.to_local()[i, j] for i in range(3) for j in range(4)] + [p.to_root()[i, j] for i in range(3) for j in range(4)]
                    + [b.lower.x, b.lower.y, b.lower.z, b.upper.x, b.upper.y, b.upper.z, s.centre.x, s.centre.y, s.centre.z, s.radius])
    np.testing.assert_array_equal(np.array(rows), g["zoo_rows"])
    glass, sf11 = api.schott("N-BK7"), api.schott("SF11")
    white = api.InterpolatedSF(scenes.CB_WAVELENGTHS, scenes.CB_WHITE)
    light = api.InterpolatedSF(*scenes.CB_LIGHT)
    for tag, (lo, hi, bins) in dict(a=(375.0, 740.0, 15), b=(400.0, 700.0, 64), c=(520.5, 530.25, 3), d=(300.0, 420.0, 7)).items():
        np.testing.assert_array_equal(white.sample(lo, hi, bins), g["white_" + tag])
        np.testing.assert_array_equal(light.sample(lo, hi, bins), g["light_" + tag])
        np.testing.assert_array_equal(glass.transmission.sample(lo, hi, bins), g["bk7_t_" + tag])
        assert glass.index.average(lo, hi) == float(g["bk7_n_" + tag])
        np.testing.assert_array_equal(sf11.transmission.sample(lo, hi, bins), g["sf11_t_" + tag])
        assert sf11.index.average(lo, hi) == float(g["sf11_n_" + tag])


def test_combine_samples_matches_host_build_of_device_code():
    """numpy StatsArray3D.combine (observer.py) vs stats_combine (rsb_path.h) incl. the n in {0, 1} branches"""
    import hostsim_api
    from source_b200.observer import combine_samples
    rng = np.random.default_rng(3)
    n_pix, fb, sb, off = 60, 7, 3, 2
    fm = rng.normal(size=(n_pix, fb)); fv = rng.uniform(0, 2, (n_pix, fb)); fs = rng.integers(0, 4, (n_pix, fb)).astype(np.int32)
    fm[fs == 0] = 0; fv[fs <= 1] = 0
    for samples in (1, 2, 9):
        m = rng.normal(size=(n_pix, sb)); v = rng.uniform(0, 2, (n_pix, sb))
        if samples == 1:
            v[...] = 0
        gm, gv, gs = fm.copy(), fv.copy(), fs.copy()
        hostsim_api.lib().hs_frame_combine(n_pix, fb, off, sb, m.ctypes.data, v.ctypes.data, samples, gm.ctypes.data, gv.ctypes.data, gs.ctypes.data)
        mt, vt, nt = combine_samples(fm[:, off:off + sb], fv[:, off:off + sb], fs[:, off:off + sb], m, v, samples)
        np.testing.assert_array_equal(nt, gs[:, off:off + sb])
        np.testing.assert_array_equal(mt, gm[:, off:off + sb])
        np.testing.assert_array_equal(vt, gv[:, off:off + sb])


def test_reciprocal_division_is_exact_ieee_division():
    """div_count (rsb_path.h): x / n via reciprocal + FMA residual correction == IEEE x / n, incl. edge values"""
    import ctypes as C
    import hostsim_api
    rng = np.random.default_rng(11)
    n = 2_000_000
    x = rng.standard_normal(n) * 10.0 ** rng.uniform(-30, 30, n)
    x[:8] = [0.0, -0.0, 1e-300, -1e-300, 1e300, np.inf, -np.inf, np.nan]
    d = rng.integers(1, 5000, n).astype(np.int32)
    d[n // 2:] = rng.integers(1, 2_000_000_000, n - n // 2)
    out = np.zeros(n)
    lib = hostsim_api.lib()
    lib.hs_div_count.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hs_div_count(n, x.ctypes.data, d.ctypes.data, out.ctypes.data)
    with np.errstate(all="ignore"):
        ref = x / d
    assert np.array_equal(out.view(np.uint64), ref.view(np.uint64)) or np.array_equal(
        out[~np.isnan(ref)].view(np.uint64), ref[~np.isnan(ref)].view(np.uint64))


def test_exact_reciprocal_division_general_divisors():
    """div_exact/exact_recip (rsb_math.h), used for the kd plane distance: identical bits to x / d for arbitrary
    doubles, including divisors with an all-ones significand, denormals, infinities and zeros (fallback path)"""
    import ctypes as C
    import hostsim_api
    rng = np.random.default_rng(5)
    n = 3_000_000
    x = rng.standard_normal(n) * 10.0 ** rng.uniform(-40, 40, n)
    d = rng.standard_normal(n) * 10.0 ** rng.uniform(-40, 40, n)
    allones = np.array([0x3FFFFFFFFFFFFFFF, 0x400FFFFFFFFFFFFF, 0xBFEFFFFFFFFFFFFF], dtype=np.uint64).view(np.float64)
    d[:3] = allones
    d[3:9] = [5e-324, 1e-310, np.inf, -np.inf, 1e308, -1e-308]
    x[9:15] = [0.0, -0.0, np.inf, 1e-320, 1e308, -1e300]
    out = np.zeros(n)
    lib = hostsim_api.lib()
    lib.hs_div_exact.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hs_div_exact(n, x.ctypes.data, d.ctypes.data, out.ctypes.data)
    with np.errstate(all="ignore"):
        ref = x / d
    ok = ~np.isnan(ref)
    assert np.array_equal(out[ok].view(np.uint64), ref[ok].view(np.uint64))
    assert np.all(np.isnan(out[~ok]))


def test_kd_plane_distance_through_reciprocal_is_the_ieee_quotient():
    """div_recip1 (rsb_math.h): (split - origin) / direction through the per-ray reciprocal with ONE fused residual
    correction must give the bits of the IEEE division for every input the traversal can meet -- adversarial
    significands (few bits set / all ones / near powers of two), zero and tiny numerators, extreme exponents (which
    must take the fallback).  Sign of a zero quotient excepted (the traversal never looks at it)."""
    import ctypes as C
    import hostsim_api
    rng = np.random.default_rng(11)
    n = 4_000_000

    def adversarial(k):
        m = rng.integers(0, 1 << 52, k, dtype=np.uint64)
        few = (np.uint64(1) << rng.integers(0, 52, k).astype(np.uint64)) | (np.uint64(1) << rng.integers(0, 52, k).astype(np.uint64))
        sel = rng.integers(0, 4, k)
        m = np.where(sel == 1, few, m)
        m = np.where(sel == 2, np.uint64((1 << 52) - 1) - (few & np.uint64(0xFFF)), m)
        m = np.where(sel == 3, few & np.uint64(0xFFF), m)
        e = (1023 + rng.integers(-30, 31, k)).astype(np.uint64)
        s = rng.integers(0, 2, k).astype(np.uint64)
        return ((s << np.uint64(63)) | (e << np.uint64(52)) | m).view(np.float64)
    x, d = adversarial(n), adversarial(n)
    x[:8] = [0.0, -0.0, 1e-320, 1e-200, 1e200, np.inf, 3e-141, 3e140]
    d[8:16] = [1e-200, 1e200, 5e-324, -1e-150, 1e150, 1.0, -1.0, 0.5]
    d[16:19] = np.array([0x3FFFFFFFFFFFFFFF, 0x400FFFFFFFFFFFFF, 0xBFEFFFFFFFFFFFFF], dtype=np.uint64).view(np.float64)
    out = np.zeros(n)
    lib = hostsim_api.lib()
    lib.hs_div_recip1.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hs_div_recip1(n, x.ctypes.data, d.ctypes.data, out.ctypes.data)
    with np.errstate(all="ignore"):
        ref = x / d
    ok = ~np.isnan(ref) & (ref != 0)
    assert np.array_equal(out[ok].view(np.uint64), ref[ok].view(np.uint64))
    zero = ref == 0
    assert np.all(out[zero] == 0)


def test_refine_mesh_is_exact_closed_and_deterministic():
    """tests/scenes.refine_mesh (SURVEY config 4 recipe: bunny -> exactly 1,000,000 triangles): exact count, every edge
    shared by two triangles with opposite directions (closed, consistently oriented), same output twice"""
    from collections import Counter
    import scenes
    v, t, _ = scenes.icosphere(1)
    for target in (320, 322, 1000, 1282):
        V, T = scenes.refine_mesh(v, t[:, :3], target)
        assert len(T) == target and V.dtype == np.float32
        und, dire = Counter(), Counter()
        for a, b, c in T:
            for e in ((a, b), (b, c), (c, a)):
                und[tuple(sorted(e))] += 1
                dire[e] += 1
        assert set(und.values()) == {2} and max(dire.values()) == 1
        V2, T2 = scenes.refine_mesh(v, t[:, :3], target)
        assert np.array_equal(V, V2) and np.array_equal(T, T2)
    with pytest.raises(ValueError):
        scenes.refine_mesh(v, t[:, :3], 321)


def test_threaded_kdtree_build_writes_the_serial_stream(monkeypatch):
    """rsb_kdtree_build forks its top levels onto threads (RSB_KD_THREADS) and splices the subtrees: the stream must be
    the serial build's byte for byte (81,920 triangle boxes, mesh tree parameters)"""
    import scenes
    from source_b200.flatten import kdtree_build
    v, t, _ = scenes.icosphere(6, radius=0.45, bumps=0.15)
    v = np.ascontiguousarray(v, dtype=np.float32)
    t = np.ascontiguousarray(t, dtype=np.int32)
    boxes = np.zeros((len(t), 6))
    lib = cabi.load()
    cabi.check(lib.rsb_mesh_triangle_boxes(cabi.ptr(v, C.c_float), len(v), cabi.ptr(t, C.c_int32), len(t), t.shape[1],
                                           cabi.ptr(boxes, C.c_double)))
    streams = []
    for threads in ("1", "3", "8"):
        monkeypatch.setenv("RSB_KD_THREADS", threads)
        streams.append(kdtree_build(boxes, 0, 1, 5.0, 0.25))
    assert streams[0] == streams[1] == streams[2] and len(streams[0]) > 1_000_000
