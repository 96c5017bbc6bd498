"""ctypes binding of ``libraysect_b200.so`` (C ABI: ``include/raysect_b200.h``).

The library is loaded from the package directory (built in-tree by ``source_b200/csrc/build.sh``
or ``__graft_entry__.build()``).  There is no CPU fallback: if the shared object is missing the
import of any compute entry point raises, and every compute call raises ``RsbError`` when no
sm_100 device is available.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RSB_LIBRARY") or os.path.join(_HERE, "libraysect_b200.so")

OK, ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_OVERFLOW = 0, 1, 2, 3, 4

PRIM_PARABOLA = -1
PRIM_SPHERE, PRIM_BOX, PRIM_CYLINDER, PRIM_CONE, PRIM_MESH, PRIM_UNION, PRIM_INTERSECT, PRIM_SUBTRACT = range(8)
PRIM_PARABOLA = -1   # analytic primitives are the types <= PRIM_CONE (include/raysect_b200.h)
PRIM_TORUS = -2
MAT_ABSORBER, MAT_EMITTER, MAT_LAMBERT, MAT_DIELECTRIC, MAT_CONDUCTOR, MAT_VOLUME_EMITTER, MAT_ROUGH_CONDUCTOR = range(7)
CAMERA_PINHOLE, CAMERA_ORTHOGRAPHIC, CAMERA_CCD, CAMERA_VECTOR, CAMERA_PIXEL = 0, 1, 2, 3, 4
PROJ_XYZ, PROJ_POWER, PROJ_RADIANCE, PROJ_MAX = 0, 1, 2, 8
RNG_MT19937_64, RNG_PHILOX = 0, 1

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)
c_uint64_p = C.POINTER(C.c_uint64)


class RsbError(RuntimeError):
    """Raised for every non-zero status of the C ABI (message from rsb_last_error())."""

    def __init__(self, code, message):
        super().__init__("libraysect_b200 error %d: %s" % (code, message))
        self.code = code


class RsbMeshDesc(C.Structure):
    _fields_ = [
        ("vertices", c_float_p),
        ("triangles", c_int32_p),
        ("vertex_normals", c_float_p),
        ("face_normals", c_float_p),
        ("kdtree", c_uint8_p),
        ("kdtree_bytes", C.c_int64),
        ("n_vertices", C.c_int32),
        ("n_triangles", C.c_int32),
        ("tri_stride", C.c_int32),
        ("n_vertex_normals", C.c_int32),
        ("smoothing", C.c_int32),
        ("closed", C.c_int32),
    ]


class RsbSceneDesc(C.Structure):
    _fields_ = [
        ("n_primitives", C.c_int32),
        ("n_world", C.c_int32),
        ("prim_type", c_int32_p),
        ("prim_material", c_int32_p),
        ("prim_child_a", c_int32_p),
        ("prim_child_b", c_int32_p),
        ("prim_mesh", c_int32_p),
        ("prim_parent", c_int32_p),
        ("prim_params", c_double_p),
        ("prim_to_local", c_double_p),
        ("prim_to_root", c_double_p),
        ("prim_root_inv", c_double_p),
        ("prim_bbox", c_double_p),
        ("world_kdtree", c_uint8_p),
        ("world_kdtree_bytes", C.c_int64),
        ("n_meshes", C.c_int32),
        ("n_materials", C.c_int32),
        ("meshes", C.POINTER(RsbMeshDesc)),
        ("mat_type", c_int32_p),
        ("mat_transmission_only", c_int32_p),
        ("n_important", C.c_int32),
        ("pad", C.c_int32),
        ("imp_sphere", c_double_p),
        ("imp_weight", c_double_p),
    ]


class RsbCamera(C.Structure):
    _fields_ = [
        ("nx", C.c_int32),
        ("ny", C.c_int32),
        ("pixel_samples", C.c_int32),
        ("kind", C.c_int32),
        ("image_delta", C.c_double),
        ("image_start_x", C.c_double),
        ("image_start_y", C.c_double),
        ("sensitivity", C.c_double),
        ("to_root", C.c_double * 12),
        ("to_root_w", C.c_double),
        ("pixel_origins", c_double_p),
        ("pixel_directions", c_double_p),
    ]


class RsbRayConfig(C.Structure):
    _fields_ = [
        ("bins", C.c_int32),
        ("extinction_min_depth", C.c_int32),
        ("max_depth", C.c_int32),
        ("importance_sampling", C.c_int32),
        ("min_wavelength", C.c_double),
        ("max_wavelength", C.c_double),
        ("extinction_prob", C.c_double),
        ("important_path_weight", C.c_double),
        ("max_distance", C.c_double),
    ]


class RsbSpectral(C.Structure):
    _fields_ = [
        ("bins", C.c_int32),
        ("n_materials", C.c_int32),
        ("tables", c_double_p),
        ("scale", c_double_p),
        ("index_in", c_double_p),
        ("index_out", c_double_p),
        ("n_tables", C.c_int32),
        ("pad", C.c_int32),
        ("table2", c_int32_p),
    ]


class RsbRngDesc(C.Structure):
    _fields_ = [("mode", C.c_int32), ("pad", C.c_int32), ("seed", C.c_uint64)]


class RsbCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("rays", "branches", "leaves", "items", "prim_tests", "tri_tests", "paths", "contains",
                 "table_reads", "contains_nodes", "contains_items", "contains_prim_tests")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class RsbRenderStats(C.Structure):
    _fields_ = [("slots", C.c_int64), ("waves", C.c_int64), ("launches", C.c_int64), ("trace_launches", C.c_int64),
                ("trace_ms", C.c_double), ("shade_ms", C.c_double), ("finalize_ms", C.c_double), ("regen_ms", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


RENDER_COUNT, RENDER_TIME_TRACE = 1, 2

# every symbol include/raysect_b200.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
_U64 = C.c_uint64
SIGNATURES = {
    "rsb_last_error": (C.c_char_p, []),
    "rsb_version": (C.c_int, []),
    "rsb_free": (None, [_VP]),
    "rsb_kdtree_build": (C.c_int, [c_double_p, C.c_int64, C.c_int32, C.c_int32, C.c_double, C.c_double,
                                   C.POINTER(_VP), C.POINTER(C.c_int64)]),
    "rsb_mesh_face_normals": (C.c_int, [c_float_p, C.c_int32, c_int32_p, C.c_int32, C.c_int32, c_float_p]),
    "rsb_mesh_triangle_boxes": (C.c_int, [c_float_p, C.c_int32, c_int32_p, C.c_int32, C.c_int32, c_double_p]),
    "rsb_context_create": (C.c_int, [C.c_int, c_uint64_p]),
    "rsb_context_destroy": (C.c_int, [_U64]),
    "rsb_device_info": (C.c_int, [_U64, c_int32_p, c_int32_p, c_int32_p, c_uint64_p]),
    "rsb_scene_create": (C.c_int, [_U64, C.POINTER(RsbSceneDesc), c_uint64_p]),
    "rsb_scene_destroy": (C.c_int, [_U64, _U64]),
    "rsb_hit_batch": (C.c_int, [_U64, _U64, C.c_int64, c_double_p, c_double_p, c_double_p, c_int32_p, c_double_p,
                                c_int32_p, c_uint8_p, c_int32_p, c_double_p, c_float_p]),
    "rsb_hit_batch_dev": (C.c_int, [_U64, _U64, _VP, C.c_int64, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP,
                                    C.c_int32]),
    "rsb_hit_sweep_dev": (C.c_int, [_U64, _U64, _VP, C.c_int64, C.c_int64, _U64, c_double_p, c_double_p, C.c_double,
                                    C.c_int32, _VP, _VP, _VP, C.c_int32]),
    "rsb_contains_batch": (C.c_int, [_U64, _U64, C.c_int64, c_double_p, C.c_int32, c_int32_p, c_int32_p]),
    "rsb_rng_uniform": (C.c_int, [_U64, _U64, C.c_int64, c_double_p]),
    "rsb_render": (C.c_int, [_U64, _U64, C.POINTER(RsbCamera), C.POINTER(RsbRayConfig), C.POINTER(RsbSpectral),
                             C.POINTER(RsbRngDesc), C.c_int64, c_int32_p, c_double_p, c_double_p, c_uint64_p]),
    "rsb_render_slice": (C.c_int, [_U64, _U64, C.POINTER(RsbCamera), C.POINTER(RsbRayConfig), C.POINTER(RsbSpectral),
                                   C.POINTER(RsbRngDesc), C.c_int32, _U64, C.c_int64, c_int32_p, c_uint64_p]),
    "rsb_render_slices": (C.c_int, [_U64, _U64, C.POINTER(RsbCamera), C.POINTER(RsbRayConfig), C.POINTER(RsbSpectral),
                                    C.POINTER(RsbRngDesc), C.c_int32, C.c_int32, _U64, C.c_int64, c_int32_p, c_uint64_p]),
    "rsb_render_slices_dev": (C.c_int, [_U64, _U64, _VP, C.POINTER(RsbCamera), C.POINTER(RsbRayConfig),
                                        C.POINTER(RsbSpectral), C.POINTER(RsbRngDesc), C.c_int32, C.c_int32, _U64, C.c_int64, _VP,
                                        _VP, _VP, _VP, C.c_int32]),
    "rsb_render_slices_xyz": (C.c_int, [_U64, _U64, C.POINTER(RsbCamera), C.POINTER(RsbRayConfig), C.POINTER(RsbSpectral),
                                        C.POINTER(RsbRngDesc), C.c_int32, C.c_int32, _U64, C.c_int64, c_int32_p, c_double_p,
                                        c_double_p, C.c_int32, c_uint64_p]),
    "rsb_slice_update_xyz_frame": (C.c_int, [_U64, C.c_int32, c_double_p, c_double_p, c_int32_p]),
    "rsb_render_slices_proj": (C.c_int, [_U64, _U64, C.POINTER(RsbCamera), C.POINTER(RsbRayConfig), C.POINTER(RsbSpectral),
                                         C.POINTER(RsbRngDesc), C.c_int32, C.c_int32, _U64, C.c_int64, c_int32_p, C.c_int32, c_int32_p,
                                         c_double_p, c_double_p, C.c_int32, c_uint64_p]),
    "rsb_slice_update_proj_frame": (C.c_int, [_U64, C.c_int32, C.c_int32, C.c_int32, c_double_p, c_double_p, c_int32_p]),
    "rsb_slice_update_bayer_frame": (C.c_int, [_U64, C.c_int32, C.c_int32, c_double_p, c_double_p, c_int32_p]),
    "rsb_comm_create": (C.c_int, [C.c_int32, c_uint64_p, c_uint64_p]),
    "rsb_comm_destroy": (C.c_int, [_U64]),
    "rsb_comm_gather_slices": (C.c_int, [_U64, C.c_int32]),
    "rsb_set_query_reorder": (C.c_int, [_U64, C.c_int32]),
    "rsb_host_pin": (C.c_int, [_U64, _VP, C.c_int64]),
    "rsb_host_unpin": (C.c_int, [_U64, _VP]),
    "rsb_slice_read": (C.c_int, [_U64, c_double_p, c_double_p]),
    "rsb_slice_update_frame": (C.c_int, [_U64, C.c_int32, C.c_int32, C.c_int32, c_double_p, c_double_p, c_int32_p]),
    "rsb_render_dev": (C.c_int, [_U64, _U64, _VP, C.POINTER(RsbCamera), C.POINTER(RsbRayConfig),
                                 C.POINTER(RsbSpectral), C.POINTER(RsbRngDesc), C.c_int64, _VP, _VP, _VP, _VP,
                                 C.c_int32]),
    "rsb_render_passes": (C.c_int, [_U64, _U64, C.POINTER(RsbCamera), C.POINTER(RsbRayConfig), C.POINTER(RsbSpectral),
                                    C.POINTER(RsbRngDesc), C.c_int32, C.c_uint64, C.c_int64, c_int32_p, c_double_p,
                                    c_double_p, c_uint64_p]),
    "rsb_render_passes_dev": (C.c_int, [_U64, _U64, _VP, C.POINTER(RsbCamera), C.POINTER(RsbRayConfig),
                                        C.POINTER(RsbSpectral), C.POINTER(RsbRngDesc), C.c_int32, C.c_uint64, C.c_int64,
                                        _VP, _VP, _VP, _VP, C.c_int32]),
    "rsb_frame_combine_dev": (C.c_int, [_U64, _VP, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int64, _VP,
                                        C.c_int32, _VP, _VP, C.c_int32, _VP, _VP, _VP]),
    "rsb_render_stats": (C.c_int, [_U64, C.POINTER(RsbRenderStats)]),
    "rsb_counters": (C.c_int, [_U64, C.POINTER(RsbCounters)]),
    "rsb_last_kernel_ms": (C.c_int, [_U64, C.POINTER(C.c_float)]),
}

_lib = None


def load():
    """Loads the shared library (once) and declares every entry point; raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s not found: build it with `sh source_b200/csrc/build.sh` (nvcc, sm_100a). "
            "source_b200 has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load().rsb_last_error()
        raise RsbError(status, msg.decode("utf-8", "replace") if msg else "")


def ptr(arr, ctype):
    """ctypes pointer to a C-contiguous numpy array (or NULL for None)."""
    if arr is None:
        return C.cast(None, C.POINTER(ctype))
    assert arr.flags["C_CONTIGUOUS"]
    return arr.ctypes.data_as(C.POINTER(ctype))


def as_f64(a, shape=None):
    out = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        out = out.reshape(shape)
    return out


def as_i32(a, shape=None):
    out = np.ascontiguousarray(a, dtype=np.int32)
    if shape is not None:
        out = out.reshape(shape)
    return out
