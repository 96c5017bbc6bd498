"""Scenegraph -> flat arrays (``RsbSceneDesc``), the host half of ``Accelerator.build``.

Works on any object model that quacks like Raysect's (the real ``raysect`` classes when the plugin is
installed into a Raysect program, or this package's own mirror in ``source_b200.scenegraph``):

* ``world.primitives``  list order = primitive id (raysect/core/acceleration/kdtree.pyx:52-55)
* ``primitive.to_local()/to_root()`` indexable ``m[i, j]``, with ``.inverse()``
* ``primitive.bounding_box()`` -> ``.lower/.upper`` with ``.x .y .z``; ``bounding_sphere()`` -> ``.centre, .radius``
* shape attributes ``radius``, ``height``, ``lower``, ``upper``, ``primitive_a/primitive_b``, ``data`` (mesh)
* ``primitive.material`` with ``importance`` and the spectral functions of the four supported materials

Everything numeric that feeds parity (AABBs, bounding spheres, matrices) is taken from the object
model's own methods, so when driven by real Raysect objects the device sees bit-identical inputs.
"""
import ctypes as C
import io
import struct

import numpy as np

from . import _cabi as cabi

_SHAPES = {
    "Sphere": cabi.PRIM_SPHERE, "Box": cabi.PRIM_BOX, "Cylinder": cabi.PRIM_CYLINDER, "Cone": cabi.PRIM_CONE,
    "Parabola": cabi.PRIM_PARABOLA, "Torus": cabi.PRIM_TORUS,
    "Mesh": cabi.PRIM_MESH, "Union": cabi.PRIM_UNION, "Intersect": cabi.PRIM_INTERSECT, "Subtract": cabi.PRIM_SUBTRACT,
}
_MATERIALS = {
    "AbsorbingSurface": cabi.MAT_ABSORBER, "UniformSurfaceEmitter": cabi.MAT_EMITTER,
    "UnitySurfaceEmitter": cabi.MAT_EMITTER,     # emitter/unity.pyx:65-76: every bin 1.0 = a constant table, scale 1
    "Checkerboard": cabi.MAT_EMITTER,            # emitter/checkerboard.pyx:38-146: two emission spectra, picked by the hit point
    "Lambert": cabi.MAT_LAMBERT, "Dielectric": cabi.MAT_DIELECTRIC,
    # HomogeneousVolumeEmitter subclasses with a direction-independent emission_function (homogeneous.pyx:40-93)
    "UniformVolumeEmitter": cabi.MAT_VOLUME_EMITTER, "UnityVolumeEmitter": cabi.MAT_VOLUME_EMITTER,
    "Conductor": cabi.MAT_CONDUCTOR,             # conductor.pyx:39-147
    "RoughConductor": cabi.MAT_ROUGH_CONDUCTOR,  # conductor.pyx:157-344 (a ContinuousBSDF, not a Conductor subclass)
}

# world tree parameters of _PrimitiveKDTree (raysect/core/acceleration/kdtree.pyx:43)
WORLD_KD = dict(max_depth=0, min_items=1, hit_cost=80.0, empty_bonus=0.2)


_EVALUATED = ("evaluate_surface", "evaluate_volume", "evaluate_shading", "hit", "next_intersection", "contains",
              "bounding_box", "bounding_sphere")


def _classify(obj, table, what):
    """Row type of a primitive / material: its own class, or a base class it inherits every evaluated method from
    unchanged.  A subclass that overrides hit / evaluate_surface / ... is NOT its base class to the device (which
    never calls back into Python): it is rejected like any unknown type -- no silent fallback."""
    for cls in type(obj).__mro__:
        if cls.__name__ in table:
            if cls is not type(obj):
                changed = [m for m in _EVALUATED if getattr(type(obj), m, None) is not getattr(cls, m, None)]
                if changed:
                    raise NotImplementedError(
                        "%s %r overrides %s of %s: the B200 path evaluates the stock %s only and never calls back into "
                        "Python; there is no CPU fallback" % (what, type(obj).__name__, ", ".join(changed), cls.__name__, cls.__name__))
            return table[cls.__name__]
    raise NotImplementedError(
        "%s %r is not supported by the B200 path (supported: %s); there is no CPU fallback"
        % (what, type(obj).__name__, ", ".join(sorted(table))))


def _encapsulated(p):
    """The primitive an ``EncapsulatedPrimitive`` (raysect/primitive/utility.pyx:36-96; the lens library's base class) hides.
    Its hit / next_intersection / contains / bounding_box hand the WORLD ray or point straight to that primitive, which hangs
    under a private local root with ``transform = wrapper.to_root()``, and only re-label the Intersection (utility.pyx:74-96)
    -- so to the device the wrapper IS the inner primitive with the inner primitive's own matrices.  The attribute is a
    private cdef field; the garbage collector's view of the object exposes it."""
    import gc
    base = next((c for c in type(p).__mro__ if c.__name__ == "EncapsulatedPrimitive"), None)
    if base is None:
        return None
    changed = [m for m in _EVALUATED if getattr(type(p), m, None) is not getattr(base, m, None)]
    if changed:
        raise NotImplementedError("primitive %r overrides %s of EncapsulatedPrimitive: the B200 path never calls back into Python; "
                                  "there is no CPU fallback" % (type(p).__name__, ", ".join(changed)))
    inner = [o for o in gc.get_referents(p)
             if o is not p and hasattr(o, "hit") and hasattr(o, "bounding_box") and type(getattr(o, "parent", None)).__name__ == "BridgeNode"]
    if len(inner) != 1:
        raise NotImplementedError("cannot find the primitive encapsulated by %r" % type(p).__name__)
    return inner[0]


def mat34(m):
    """An AffineMatrix3D as the C ABI carries it: rows 0..2, then m33 (13 values).  Raysect's Point3D.transform divides
    by w = m30 x + m31 y + m32 z + m33 (point.pyx:272-281); the bottom row of an affine matrix is (0, 0, 0, m33), and
    AffineMatrix3D.inverse() of a non-rigid chain can leave m33 one ulp off 1.0 -- the device multiplies by 1 / m33
    exactly as the reference does."""
    if [float(m[3, j]) for j in range(3)] != [0.0, 0.0, 0.0]:
        raise NotImplementedError("a primitive transform is not affine (bottom row %r): projective transforms are not "
                                  "supported by the B200 path; there is no CPU fallback" % ([float(m[3, j]) for j in range(4)],))
    return [float(m[i, j]) for i in range(3) for j in range(4)] + [float(m[3, 3])]


def _box6(b):
    return [b.lower.x, b.lower.y, b.lower.z, b.upper.x, b.upper.y, b.upper.z]


def lambert_reflectivity(material):
    """``Lambert.reflectivity`` is a private cdef field in Raysect; its pickle state exposes it."""
    if hasattr(material, "reflectivity"):
        return material.reflectivity
    return material.__reduce__()[2][-1]


def kdtree_build(boxes, max_depth=0, min_items=1, hit_cost=20.0, empty_bonus=0.2):
    """SAH kd-tree over item AABBs -> the reference's serialised stream (bytes).  Host-only C++."""
    lib = cabi.load()
    boxes = cabi.as_f64(boxes).reshape(-1, 6)
    out = C.c_void_p()
    nbytes = C.c_int64()
    cabi.check(lib.rsb_kdtree_build(cabi.ptr(boxes, C.c_double), boxes.shape[0], max_depth, min_items, hit_cost,
                                    empty_bonus, C.byref(out), C.byref(nbytes)))
    try:
        return C.string_at(out, nbytes.value)
    finally:
        lib.rsb_free(out)


def rsm_kdtree_stream(blob):
    """Offset of the KDTree3DCore stream inside an .rsm blob (raysect/primitive/mesh/mesh.pyx:864-931)."""
    if blob[:3] != b"RSM":
        raise ValueError("Specified file is not a Raysect mesh file.")
    major, minor = struct.unpack_from("<BB", blob, 3)
    if (major, minor) != (1, 0):
        raise ValueError("Unsupported Raysect mesh version.")
    nv, nn, nt = struct.unpack_from("<iii", blob, 8)
    width = 6 if nn > 0 else 3
    return 20 + 12 * nv + 12 * nn + 4 * width * nt


class FlatScene:
    """Flat arrays + the ctypes descriptor that points into them (keeps everything alive)."""

    def __init__(self):
        self.primitives = []        # world-level primitive objects, id order
        self.materials = []         # material objects, row order
        self.rows = []
        self.meshes = []
        self.desc = None

    # ---- per-slice spectral tables ------------------------------------------------------------------
    def spectral(self, min_wavelength, max_wavelength, bins):
        """What SpectralFunction.sample()/average() return for every material on this slice
        (raysect/optical/spectralfunction.pyx:140-216); evaluated by the object model itself."""
        n = len(self.materials)
        # materials that sample two spectral functions (Conductor: index, extinction) get an extra table row each
        # (and Checkerboard: the emission of its second kind of square)
        second = [i for i, (m, t) in enumerate(zip(self.materials, self.mat_type))
                  if t in (cabi.MAT_CONDUCTOR, cabi.MAT_ROUGH_CONDUCTOR) or hasattr(m, "emission_spectrum2")]
        table2 = np.full(n, -1, dtype=np.int32)
        for j, i in enumerate(second):
            table2[i] = n + j
        tables = np.zeros((n + len(second), bins), dtype=np.float64)
        scale = np.ones(n, dtype=np.float64)
        index_in = np.ones(n, dtype=np.float64)
        index_out = np.ones(n, dtype=np.float64)
        for i, (m, t) in enumerate(zip(self.materials, self.mat_type)):
            if t == cabi.MAT_LAMBERT:
                tables[i] = np.asarray(lambert_reflectivity(m).sample(min_wavelength, max_wavelength, bins))
            elif t == cabi.MAT_EMITTER and hasattr(m, "emission_spectrum2"):
                # Checkerboard (checkerboard.pyx:101-127): square one -> (row i, scale), square two -> (row table2[i], index_in);
                # index_out carries 1 / width (its _rwidth)
                tables[i] = np.asarray(m.emission_spectrum1.sample(min_wavelength, max_wavelength, bins))
                tables[table2[i]] = np.asarray(m.emission_spectrum2.sample(min_wavelength, max_wavelength, bins))
                scale[i] = m.scale1
                index_in[i] = m.scale2
                index_out[i] = 1.0 / m.width
            elif t in (cabi.MAT_EMITTER, cabi.MAT_VOLUME_EMITTER) and not hasattr(m, "emission_spectrum"):
                tables[i] = 1.0     # Unity{Surface,Volume}Emitter (emitter/unity.pyx:73-75, 98): samples[:] = 1.0; 1.0 * 1.0 is exact
            elif t in (cabi.MAT_EMITTER, cabi.MAT_VOLUME_EMITTER):
                tables[i] = np.asarray(m.emission_spectrum.sample(min_wavelength, max_wavelength, bins))
                scale[i] = m.scale
            elif t == cabi.MAT_DIELECTRIC:
                tables[i] = np.asarray(m.transmission.sample(min_wavelength, max_wavelength, bins))
                index_in[i] = m.index.average(min_wavelength, max_wavelength)
                index_out[i] = m.external_index.average(min_wavelength, max_wavelength)
            elif t in (cabi.MAT_CONDUCTOR, cabi.MAT_ROUGH_CONDUCTOR):
                # conductor.pyx:101-102, 321-322: n and k resampled onto the ray's bins
                tables[i] = np.asarray(m.index.sample(min_wavelength, max_wavelength, bins))
                tables[table2[i]] = np.asarray(m.extinction.sample(min_wavelength, max_wavelength, bins))
                if t == cabi.MAT_ROUGH_CONDUCTOR:
                    scale[i] = m.roughness
        s = cabi.RsbSpectral()
        s.bins = bins
        s.n_materials = n
        s.tables = cabi.ptr(tables, C.c_double)
        s.scale = cabi.ptr(scale, C.c_double)
        s.index_in = cabi.ptr(index_in, C.c_double)
        s.index_out = cabi.ptr(index_out, C.c_double)
        s.n_tables = tables.shape[0]
        s.table2 = cabi.ptr(table2, C.c_int32)
        s._keep = (tables, scale, index_in, index_out, table2)
        return s


def _mesh_desc(data, keep):
    vertices = np.ascontiguousarray(data.vertices, dtype=np.float32)
    triangles = np.ascontiguousarray(data.triangles, dtype=np.int32)
    vnormals = getattr(data, "vertex_normals", None)
    if vnormals is not None:
        vnormals = np.ascontiguousarray(vnormals, dtype=np.float32)
    fnormals = getattr(data, "face_normals", None)
    if fnormals is not None:
        fnormals = np.ascontiguousarray(fnormals, dtype=np.float32)
    stream = getattr(data, "kdtree_stream", None)
    if stream is None:
        buf = io.BytesIO()
        data.save(buf)                      # MeshData.save -> .rsm blob (arrays + the mesh's own kd-tree)
        blob = buf.getvalue()
        stream = blob[rsm_kdtree_stream(blob):]
    stream = np.frombuffer(stream, dtype=np.uint8)
    d = cabi.RsbMeshDesc()
    d.vertices = cabi.ptr(vertices, C.c_float)
    d.triangles = cabi.ptr(triangles, C.c_int32)
    d.vertex_normals = cabi.ptr(vnormals, C.c_float)
    d.face_normals = cabi.ptr(fnormals, C.c_float)
    d.kdtree = cabi.ptr(stream, C.c_uint8)
    d.kdtree_bytes = stream.size
    d.n_vertices = vertices.shape[0]
    d.n_triangles = triangles.shape[0]
    d.tri_stride = triangles.shape[1]
    d.n_vertex_normals = 0 if vnormals is None else vnormals.shape[0]
    d.smoothing = int(bool(data.smoothing))
    d.closed = int(bool(data.closed))
    keep.extend([vertices, triangles, vnormals, fnormals, stream])
    return d


def flatten_world(world, world_kdtree=None):
    """Flattens ``world`` (Raysect ``World`` or this package's mirror) into a ``FlatScene``.

    ``world_kdtree``: optional pre-serialised world tree (bytes); by default the tree is built by this
    package's own bit-exact SAH builder with the reference's parameters.
    """
    flat = FlatScene()
    prims = list(world.primitives)   # may be empty: the reference builds a one-leaf tree and every query misses
    flat.primitives = prims
    rows = []          # dict rows
    mesh_descs, mesh_index, keep = [], {}, []
    material_index = {}
    encapsulated = {}      # id(EncapsulatedPrimitive) -> the primitive it hides

    def material_row(material):
        key = id(material)
        if key not in material_index:
            _classify(material, _MATERIALS, "material")
            material_index[key] = len(flat.materials)
            flat.materials.append(material)
        return material_index[key]

    def add(p, parent_row, top_level):
        inner = _encapsulated(p)
        if inner is not None:
            # (lenses: EncapsulatedPrimitive(Intersect(Intersect(Sphere, Sphere), Cylinder)) and the like)
            encapsulated[id(p)] = inner
            return add(inner, parent_row, top_level)
        t = _classify(p, _SHAPES, "primitive")
        row = dict(type=t, material=-1, a=-1, b=-1, mesh=-1, parent=parent_row, params=[0.0] * 6,
                   to_local=mat34(p.to_local()), to_root=mat34(p.to_root()), root_inv=mat34(p.to_local()),
                   bbox=_box6(p.bounding_box()))
        idx = len(rows) if not top_level else None
        if not top_level:
            # Normal3D.transform(primitive_to_world) re-inverts the matrix (normal.pyx:241)
            row["root_inv"] = mat34(p.to_root().inverse())
            rows.append(row)
        if t == cabi.PRIM_SPHERE:
            row["params"][0] = p.radius
        elif t == cabi.PRIM_BOX:
            row["params"] = [p.lower.x, p.lower.y, p.lower.z, p.upper.x, p.upper.y, p.upper.z]
        elif t == cabi.PRIM_TORUS:
            if not top_level:
                raise NotImplementedError("Torus as a CSG operand (a torus has up to four crossings; operands here carry two)")
            row["params"][0] = p.major_radius
            row["params"][1] = p.minor_radius
        elif t in (cabi.PRIM_CYLINDER, cabi.PRIM_CONE, cabi.PRIM_PARABOLA):
            row["params"][0] = p.radius
            row["params"][1] = p.height
        elif t == cabi.PRIM_MESH:
            key = id(p.data)
            if key not in mesh_index:
                mesh_index[key] = len(mesh_descs)
                mesh_descs.append(_mesh_desc(p.data, keep))
            row["mesh"] = mesh_index[key]
        return row, idx

    # world-level rows first, in World.primitives order
    for p in prims:
        row, _ = add(p, -1, True)
        row["material"] = material_row(p.material)
        rows.append(row)
    # then CSG operands, depth first
    def expand(p, my_row):
        p = encapsulated.get(id(p), p)
        t = rows[my_row]["type"]
        if t < cabi.PRIM_UNION:
            return
        for key, child in (("a", p.primitive_a), ("b", p.primitive_b)):
            _, idx = add(child, my_row, False)
            rows[my_row][key] = idx
            expand(child, idx)

    for i, p in enumerate(prims):
        expand(p, i)

    n = len(rows)
    flat.rows = rows
    flat.prim_type = cabi.as_i32([r["type"] for r in rows])
    flat.prim_material = cabi.as_i32([r["material"] for r in rows])
    flat.prim_child_a = cabi.as_i32([r["a"] for r in rows])
    flat.prim_child_b = cabi.as_i32([r["b"] for r in rows])
    flat.prim_mesh = cabi.as_i32([r["mesh"] for r in rows])
    flat.prim_parent = cabi.as_i32([r["parent"] for r in rows])
    flat.prim_params = cabi.as_f64([r["params"] for r in rows], (n, 6))
    flat.prim_to_local = cabi.as_f64([r["to_local"] for r in rows], (n, 13))
    flat.prim_to_root = cabi.as_f64([r["to_root"] for r in rows], (n, 13))
    flat.prim_root_inv = cabi.as_f64([r["root_inv"] for r in rows], (n, 13))
    flat.prim_bbox = cabi.as_f64([r["bbox"] for r in rows], (n, 6))

    if world_kdtree is None:
        world_kdtree = kdtree_build(flat.prim_bbox[:len(prims)], **WORLD_KD)
    flat.world_kdtree = np.frombuffer(world_kdtree, dtype=np.uint8)

    flat.mat_type = cabi.as_i32([_classify(m, _MATERIALS, "material") for m in flat.materials])
    flat.mat_transmission_only = cabi.as_i32(
        [int(bool(getattr(m, "transmission_only", False))) for m in flat.materials])

    # ImportanceManager._process_primitives (raysect/optical/scenegraph/world.pyx:88-108)
    spheres, weights = [], []
    for p in prims:
        importance = getattr(p.material, "importance", 0.0)
        if importance > 0:
            s = p.bounding_sphere()
            spheres.append([s.centre.x, s.centre.y, s.centre.z, s.radius])
            weights.append(importance)
    flat.imp_sphere = cabi.as_f64(spheres if spheres else np.zeros((0, 4)), (-1, 4))
    flat.imp_weight = cabi.as_f64(weights)

    flat.meshes = (cabi.RsbMeshDesc * max(1, len(mesh_descs)))(*mesh_descs)
    flat._keep = keep

    d = cabi.RsbSceneDesc()
    d.n_primitives = n
    d.n_world = len(prims)
    d.prim_type = cabi.ptr(flat.prim_type, C.c_int32)
    d.prim_material = cabi.ptr(flat.prim_material, C.c_int32)
    d.prim_child_a = cabi.ptr(flat.prim_child_a, C.c_int32)
    d.prim_child_b = cabi.ptr(flat.prim_child_b, C.c_int32)
    d.prim_mesh = cabi.ptr(flat.prim_mesh, C.c_int32)
    d.prim_parent = cabi.ptr(flat.prim_parent, C.c_int32)
    d.prim_params = cabi.ptr(flat.prim_params, C.c_double)
    d.prim_to_local = cabi.ptr(flat.prim_to_local, C.c_double)
    d.prim_to_root = cabi.ptr(flat.prim_to_root, C.c_double)
    d.prim_root_inv = cabi.ptr(flat.prim_root_inv, C.c_double)
    d.prim_bbox = cabi.ptr(flat.prim_bbox, C.c_double)
    d.world_kdtree = cabi.ptr(flat.world_kdtree, C.c_uint8)
    d.world_kdtree_bytes = flat.world_kdtree.size
    d.n_meshes = len(mesh_descs)
    d.n_materials = len(flat.materials)
    d.meshes = C.cast(flat.meshes, C.POINTER(cabi.RsbMeshDesc))
    d.mat_type = cabi.ptr(flat.mat_type, C.c_int32)
    d.mat_transmission_only = cabi.ptr(flat.mat_transmission_only, C.c_int32)
    d.n_important = len(weights)
    d.imp_sphere = cabi.ptr(flat.imp_sphere, C.c_double)
    d.imp_weight = cabi.ptr(flat.imp_weight, C.c_double)
    flat.desc = d
    return flat
