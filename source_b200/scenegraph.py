"""Scene description front-end with Raysect's names and argument meaning.

A thin mirror of the part of ``raysect.core.scenegraph`` / ``raysect.primitive`` that the hot path
consumes: nodes with parent/transform, the five analytic/CSG/mesh primitive families and ``World``
with ``hit``/``contains``.  It exists so that the B200 path can be driven without Raysect installed
(bench, GPU tests); with Raysect installed, use the plugin classes in ``source_b200.plugin`` on the
real objects instead -- both feed the same flattener (``source_b200.flatten``).

Geometry queries are NOT evaluated here: ``World.hit`` / ``World.contains`` go to the device through
``CudaAccelerator``-equivalent batched calls (1-element batches for the scalar API).
"""
import math

import numpy as np

from . import _cabi as cabi
from .math3d import AffineMatrix3D, BoundingBox3D, BoundingSphere3D, Point3D, Vector3D

_BOX_PADDING = 1e-9       # sphere.pyx:38, box.pyx:37, cylinder.pyx:38, cone.pyx:38, csg.pyx:39
_SPHERE_PADDING = 1.000000001   # sphere.pyx:39
_MESH_BOX_PADDING = 1e-6  # mesh.pyx:49


class Node:
    """raysect/core/scenegraph/node.pyx + _nodebase.pyx: parent, transform, root transforms."""

    def __init__(self, parent=None, transform=None, name=None):
        self._parent = None
        self.children = []
        self._transform = transform if transform is not None else AffineMatrix3D()
        self.name = name
        self.root = self
        self._root_transform = AffineMatrix3D()
        self._root_transform_inverse = AffineMatrix3D()
        self.parent = parent

    # -- tree plumbing ---------------------------------------------------------------------------------
    @property
    def parent(self):
        return self._parent

    @parent.setter
    def parent(self, value):
        if value is self._parent:
            return
        if self._parent is not None:
            self._parent.children.remove(self)
        self._parent = value
        if value is not None:
            node = value
            while node is not None:
                if node is self:
                    raise ValueError("A node cannot be parented to itself or one of its descendants.")
                node = node._parent
            value.children.append(self)
        self._update()

    @property
    def transform(self):
        return self._transform

    @transform.setter
    def transform(self, value):
        self._transform = value
        self._update()
        self.root._change(self)

    def _update(self):
        """_nodebase.pyx:83-134"""
        if self._parent is None:
            if self.root is not self:
                self.root._deregister(self)
                self.root = self
            self._root_transform = AffineMatrix3D()
            self._root_transform_inverse = AffineMatrix3D()
        else:
            if self.root is not self._parent.root:
                self.root._deregister(self)
                self.root = self._parent.root
                self.root._register(self)
            self._root_transform = self._parent._root_transform.mul(self._transform)
            self._root_transform_inverse = self._root_transform.inverse()
        for child in self.children:
            child._update()

    def _register(self, node):
        pass

    def _deregister(self, node):
        pass

    def _change(self, node):
        pass

    def to_local(self):
        """node.pyx:173-180: root space -> this node's space"""
        return self._root_transform_inverse

    def to_root(self):
        """node.pyx:182-189"""
        return self._root_transform


class Primitive(Node):
    """raysect/core/scenegraph/primitive.pyx:33-224"""

    def __init__(self, parent=None, transform=None, material=None, name=None):
        self.material = material
        super().__init__(parent, transform, name)

    def notify_geometry_change(self):
        self.root._change(self)

    def bounding_box(self):
        raise NotImplementedError("Primitive surface has not been defined. Virtual method bounding_box() has not been implemented.")

    def bounding_sphere(self):
        """primitive.pyx:166-184: default wraps the bounding box"""
        return self.bounding_box().enclosing_sphere()

    def _world_box(self, lower, upper):
        """box.pyx:361-383 et al.: the 8 local corners -> world, padded extend"""
        box = BoundingBox3D()
        for point in BoundingBox3D(lower, upper).vertices():
            box.extend(point.transform(self.to_root()), _BOX_PADDING)
        return box


class Sphere(Primitive):
    """raysect/primitive/sphere.pyx"""

    def __init__(self, radius=0.5, parent=None, transform=None, material=None, name=None):
        if radius < 0.0:
            raise ValueError("Sphere radius cannot be less than zero.")
        self._radius = float(radius)
        super().__init__(parent, transform, material, name)

    @property
    def radius(self):
        return self._radius

    @radius.setter
    def radius(self, value):
        if value == self._radius:
            return
        if value < 0.0:
            raise ValueError("Sphere radius cannot be less than zero.")
        self._radius = float(value)
        self.notify_geometry_change()

    def bounding_box(self):
        """sphere.pyx:216-229"""
        origin = Point3D(0, 0, 0).transform(self.to_root())
        extent = self._radius + _BOX_PADDING
        return BoundingBox3D(Point3D(origin.x - extent, origin.y - extent, origin.z - extent),
                             Point3D(origin.x + extent, origin.y + extent, origin.z + extent))

    def bounding_sphere(self):
        """sphere.pyx:231-233"""
        return BoundingSphere3D(Point3D(0, 0, 0).transform(self.to_root()), self._radius * _SPHERE_PADDING)


class Box(Primitive):
    """raysect/primitive/box.pyx"""

    def __init__(self, lower=None, upper=None, parent=None, transform=None, material=None, name=None):
        if lower is not None and upper is not None:
            if lower.x > upper.x or lower.y > upper.y or lower.z > upper.z:
                raise ValueError("The lower point coordinates must be less than or equal to the upper point coordinates.")
            self._lower, self._upper = lower.copy(), upper.copy()
        else:
            self._lower, self._upper = Point3D(-0.5, -0.5, -0.5), Point3D(0.5, 0.5, 0.5)
        super().__init__(parent, transform, material, name)

    @property
    def lower(self):
        return self._lower

    @property
    def upper(self):
        return self._upper

    def bounding_box(self):
        return self._world_box(self._lower, self._upper)


class _RadiusHeight(Primitive):
    _what = "primitive"

    def __init__(self, radius=0.5, height=1.0, parent=None, transform=None, material=None, name=None):
        if radius <= 0.0:
            raise ValueError("%s radius cannot be less than or equal to zero." % self._what)
        if height <= 0.0:
            raise ValueError("%s height cannot be less than or equal to zero." % self._what)
        self._radius, self._height = float(radius), float(height)
        super().__init__(parent, transform, material, name)

    @property
    def radius(self):
        return self._radius

    @property
    def height(self):
        return self._height

    def bounding_box(self):
        """cylinder.pyx:369-391, cone.pyx:382-404"""
        return self._world_box(Point3D(-self._radius, -self._radius, 0.0), Point3D(self._radius, self._radius, self._height))


class Cylinder(_RadiusHeight):
    """raysect/primitive/cylinder.pyx"""
    _what = "Cylinder"


class Cone(_RadiusHeight):
    """raysect/primitive/cone.pyx"""
    _what = "Cone"


class Parabola(_RadiusHeight):
    """raysect/primitive/parabola.pyx: paraboloid of revolution, tip at z = height, base disc of ``radius`` at z = 0"""
    _what = "Parabola"


class CSGRoot(Node):
    """raysect/primitive/csg.pyx:258-287: root of the operand sub-graph; operands live in CSG-local space"""

    def __init__(self, csg_primitive):
        super().__init__()
        self.csg_primitive = csg_primitive

    def _change(self, node):
        self.csg_primitive.root._change(node)


class CSGPrimitive(Primitive):
    """raysect/primitive/csg.pyx:42-241"""

    def __init__(self, primitive_a=None, primitive_b=None, parent=None, transform=None, material=None, name=None):
        if primitive_a is None or primitive_b is None:
            raise ValueError("CSG operands must be primitives.")
        self._csgroot = CSGRoot(self)
        self._primitive_a, self._primitive_b = primitive_a, primitive_b
        primitive_a.parent = self._csgroot
        primitive_b.parent = self._csgroot
        super().__init__(parent, transform, material, name)

    @property
    def primitive_a(self):
        return self._primitive_a

    @property
    def primitive_b(self):
        return self._primitive_b

    def _corner_box(self, box):
        out = BoundingBox3D()
        for point in box.vertices():
            out.extend(point.transform(self.to_root()), _BOX_PADDING)
        return out


class Union(CSGPrimitive):
    def bounding_box(self):
        """csg.pyx:354-369"""
        box = BoundingBox3D()
        box.union(self._primitive_a.bounding_box())
        box.union(self._primitive_b.bounding_box())
        return self._corner_box(box)


class Intersect(CSGPrimitive):
    def bounding_box(self):
        """csg.pyx:452-480"""
        a, b = self._primitive_a.bounding_box(), self._primitive_b.bounding_box()
        box = BoundingBox3D()
        box.lower = Point3D(max(a.lower.x, b.lower.x), max(a.lower.y, b.lower.y), max(a.lower.z, b.lower.z))
        box.upper = Point3D(min(a.upper.x, b.upper.x), min(a.upper.y, b.upper.y), min(a.upper.z, b.upper.z))
        return self._corner_box(box)


class Subtract(CSGPrimitive):
    def bounding_box(self):
        """csg.pyx:574-590"""
        return self._corner_box(self._primitive_a.bounding_box())


class MeshData:
    """raysect/primitive/mesh/mesh.pyx:125-504 (MeshData): arrays + kd-tree, built on the host.

    The kd-tree is built by the library's SAH builder over the per-triangle padded boxes and kept as
    the reference's serialised stream (``kdtree_stream``).
    """

    def __init__(self, vertices, triangles, normals=None, smoothing=True, closed=True, tolerant=True,
                 flip_normals=False, max_depth=0, min_items=1, hit_cost=20.0, empty_bonus=0.2, kdtree_stream=None):
        import ctypes as C
        from .flatten import kdtree_build
        self.smoothing = bool(smoothing)
        self.closed = bool(closed)
        vertices = np.array(vertices, dtype=np.float32)
        triangles = np.array(triangles, dtype=np.int32)
        vertex_normals = None if normals is None else np.array(normals, dtype=np.float32)
        if vertices.ndim != 2 or vertices.shape[1] != 3:
            raise ValueError("The vertex array must have dimensions Nx3.")
        if vertex_normals is not None:
            if vertex_normals.ndim != 2 or vertex_normals.shape[1] != 3:
                raise ValueError("The normal array must have dimensions Nx3.")
            if triangles.ndim != 2 or triangles.shape[1] != 6:
                raise ValueError("The triangle array must have dimensions Nx6.")
        else:
            if triangles.ndim != 2 or triangles.shape[1] != 3:
                raise ValueError("The triangle array must have dimensions Nx3.")
        invalid = (triangles[:, 0:3] < 0) | (triangles[:, 0:3] >= vertices.shape[0])
        if invalid.any():
            raise ValueError("The triangle array references non-existent vertices.")
        if vertex_normals is not None:
            invalid = (triangles[:, 3:6] < 0) | (triangles[:, 3:6] >= vertex_normals.shape[0])
            if invalid.any():
                raise ValueError("The triangle array references non-existent normals.")
        if tolerant:
            triangles = self._filter_triangles(vertices, triangles)
        if flip_normals:
            # mesh.pyx:402-422
            triangles = triangles.copy()
            triangles[:, [0, 2]] = triangles[:, [2, 0]]
            if vertex_normals is not None:
                triangles[:, [3, 5]] = triangles[:, [5, 3]]
                vertex_normals = -vertex_normals
        self.vertices = np.ascontiguousarray(vertices)
        self.triangles = np.ascontiguousarray(triangles)
        self.vertex_normals = None if vertex_normals is None else np.ascontiguousarray(vertex_normals)
        lib = cabi.load()
        nt, stride = self.triangles.shape
        self.face_normals = np.zeros((nt, 3), dtype=np.float32)
        cabi.check(lib.rsb_mesh_face_normals(cabi.ptr(self.vertices, C.c_float), self.vertices.shape[0],
                                             cabi.ptr(self.triangles, C.c_int32), nt, stride,
                                             cabi.ptr(self.face_normals, C.c_float)))
        if kdtree_stream is None:
            boxes = np.zeros((nt, 6), dtype=np.float64)
            cabi.check(lib.rsb_mesh_triangle_boxes(cabi.ptr(self.vertices, C.c_float), self.vertices.shape[0],
                                                   cabi.ptr(self.triangles, C.c_int32), nt, stride,
                                                   cabi.ptr(boxes, C.c_double)))
            kdtree_stream = kdtree_build(boxes, max(0, max_depth), min_items, hit_cost, empty_bonus)
        self.kdtree_stream = kdtree_stream

    @staticmethod
    def _filter_triangles(vertices, triangles):
        """mesh.pyx:363-399: drop triangles whose edge cross product has zero length (float64 maths)"""
        v = vertices.astype(np.float64)
        p1, p2, p3 = v[triangles[:, 0]], v[triangles[:, 1]], v[triangles[:, 2]]
        a, b = p2 - p1, p3 - p1
        cx = a[:, 1] * b[:, 2] - b[:, 1] * a[:, 2]
        cy = a[:, 2] * b[:, 0] - b[:, 2] * a[:, 0]
        cz = a[:, 0] * b[:, 1] - b[:, 0] * a[:, 1]
        length = np.sqrt(cx * cx + cy * cy + cz * cz)
        return triangles[length != 0.0]

    def save(self, file):
        """mesh.pyx:864-931 (MeshData.save): the Raysect mesh file (.rsm) -- "RSM", version 1.0, smoothing / closed /
        has-kdtree flags, counts, vertices f32, vertex normals f32, triangles i32, then the kd-tree stream
        (kdtree3d.pyx:864-912).  ``file`` is a binary stream or a file name.  Byte-identical to what Raysect writes
        for the same mesh (tests/test_plugin_host.py)."""
        import io
        import struct
        close = False
        if not isinstance(file, io.IOBase):
            file = open(file, mode="wb")
            close = True
        nn = 0 if self.vertex_normals is None else self.vertex_normals.shape[0]
        file.write(b"RSM")
        file.write(struct.pack("<BB", 1, 0))
        file.write(struct.pack("<???", self.smoothing, self.closed, True))
        file.write(struct.pack("<iii", self.vertices.shape[0], nn, self.triangles.shape[0]))
        file.write(np.ascontiguousarray(self.vertices, dtype="<f4").tobytes())
        if nn:
            file.write(np.ascontiguousarray(self.vertex_normals, dtype="<f4").tobytes())
        file.write(np.ascontiguousarray(self.triangles, dtype="<i4").tobytes())
        file.write(bytes(self.kdtree_stream))
        if close:
            file.close()

    @classmethod
    def from_rsm(cls, blob):
        """mesh.pyx:933-1024: load arrays + kd-tree from a Raysect mesh (.rsm) blob"""
        import struct
        from .flatten import rsm_kdtree_stream
        off = rsm_kdtree_stream(blob)
        smoothing, closed, _ = struct.unpack_from("<???", blob, 5)
        nv, nn, nt = struct.unpack_from("<iii", blob, 8)
        p = 20
        vertices = np.frombuffer(blob, dtype="<f4", count=3 * nv, offset=p).reshape(nv, 3)
        p += 12 * nv
        normals = None
        if nn > 0:
            normals = np.frombuffer(blob, dtype="<f4", count=3 * nn, offset=p).reshape(nn, 3)
            p += 12 * nn
        width = 6 if nn > 0 else 3
        triangles = np.frombuffer(blob, dtype="<i4", count=width * nt, offset=p).reshape(nt, width)
        return cls(vertices, triangles, normals, smoothing=smoothing, closed=closed, tolerant=False,
                   kdtree_stream=blob[off:])


class Mesh(Primitive):
    """raysect/primitive/mesh/mesh.pyx:1027-1308 (Mesh)"""

    def __init__(self, vertices=None, triangles=None, normals=None, smoothing=True, closed=True, tolerant=True,
                 flip_normals=False, kdtree_max_depth=-1, kdtree_min_items=1, kdtree_hit_cost=5.0,
                 kdtree_empty_bonus=0.25, parent=None, transform=None, material=None, name=None, data=None):
        if data is None:
            if vertices is None or triangles is None:
                raise ValueError("Vertices and triangle arrays must be supplied if the mesh is not configured to be an instance.")
            data = MeshData(vertices, triangles, normals, smoothing, closed, tolerant, flip_normals,
                            kdtree_max_depth, kdtree_min_items, kdtree_hit_cost, kdtree_empty_bonus)
        self.data = data
        super().__init__(parent, transform, material, name)

    def instance(self, parent=None, transform=None, material=None, name=None):
        return Mesh(data=self.data, parent=parent, transform=transform, material=material, name=name)

    def save(self, file):
        """Mesh.save (mesh.pyx:1310-1327): write the mesh and its kd-tree as a .rsm file"""
        self.data.save(file)

    @classmethod
    def from_file(cls, file, parent=None, transform=None, material=None, name=None):
        """Mesh.from_file (mesh.pyx:1343-1370): a mesh from a .rsm file or stream -- the kd-tree comes with it, nothing
        is rebuilt"""
        if hasattr(file, "read"):
            blob = file.read()
        else:
            with open(file, "rb") as f:
                blob = f.read()
        return cls(data=MeshData.from_rsm(blob), parent=parent, transform=transform, material=material, name=name)

    def bounding_box(self):
        """mesh.pyx:835-858: every vertex -> world (float32 -> float64), padded extend"""
        m = self.to_root().m
        v = self.data.vertices.astype(np.float64)
        x, y, z = v[:, 0], v[:, 1], v[:, 2]
        w = m[3][0] * x + m[3][1] * y + m[3][2] * z + m[3][3]
        w = 1.0 / w
        wx = (m[0][0] * x + m[0][1] * y + m[0][2] * z + m[0][3]) * w
        wy = (m[1][0] * x + m[1][1] * y + m[1][2] * z + m[1][3]) * w
        wz = (m[2][0] * x + m[2][1] * y + m[2][2] * z + m[2][3]) * w
        pad = _MESH_BOX_PADDING
        return BoundingBox3D(Point3D((wx - pad).min(), (wy - pad).min(), (wz - pad).min()),
                             Point3D((wx + pad).max(), (wy + pad).max(), (wz + pad).max()))


class Ray:
    """raysect/core/ray.pyx: origin, direction (not required to be unit length), max_distance"""

    def __init__(self, origin=None, direction=None, max_distance=math.inf):
        self.origin = origin if origin is not None else Point3D(0, 0, 0)
        self.direction = direction if direction is not None else Vector3D(0, 0, 1)
        self.max_distance = float(max_distance)


class Intersection:
    """raysect/core/intersection.pxd:37-53 (+ MeshIntersection fields, mesh.pxd:37-41)"""

    def __init__(self, ray, ray_distance, primitive, hit_point, inside_point, outside_point, normal, exiting,
                 world_to_primitive, primitive_to_world):
        self.ray = ray
        self.ray_distance = ray_distance
        self.primitive = primitive
        self.hit_point = hit_point
        self.inside_point = inside_point
        self.outside_point = outside_point
        self.normal = normal
        self.exiting = exiting
        self.world_to_primitive = world_to_primitive
        self.primitive_to_world = primitive_to_world
        self.triangle = -1
        self.u = self.v = self.w = 0.0


class World(Node):
    """raysect/core/scenegraph/world.pyx + raysect/optical/scenegraph/world.pyx: scene-graph root that
    owns the accelerator.  ``hit``/``contains`` run on the GPU through ``source_b200.engine.Device``."""

    def __init__(self, name=None, device=None):
        super().__init__(None, None, name)
        self._primitives = []
        self._observers = []
        self._rebuild = True
        self._accel = None
        self._device = device

    @property
    def primitives(self):
        return self._primitives

    @property
    def observers(self):
        return self._observers

    def _register(self, node):
        """world.pyx:196-206"""
        from .observer import Observer
        if isinstance(node, Primitive):
            self._primitives.append(node)
            self._rebuild = True
        if isinstance(node, Observer):
            self._observers.append(node)

    def _deregister(self, node):
        from .observer import Observer
        if isinstance(node, Primitive):
            self._primitives.remove(node)
            self._rebuild = True
        if isinstance(node, Observer):
            self._observers.remove(node)

    def _change(self, node):
        """world.pyx:220-238: a GEOMETRY signal schedules a rebuild on the next query"""
        self._rebuild = True

    @property
    def device(self):
        if self._device is None:
            from .engine import default_device
            self._device = default_device()
        return self._device

    def build_accelerator(self, force=False):
        """world.pyx:170-194"""
        if self._rebuild or force or self._accel is None:
            if self._accel is not None:
                self._accel.close()
            self._accel = self.device.build(self)
            self._rebuild = False
        return self._accel

    def hit(self, ray):
        """world.pyx:125-146"""
        accel = self.build_accelerator()
        return accel.hit(ray)

    def contains(self, point):
        """world.pyx:148-168"""
        accel = self.build_accelerator()
        return accel.contains(point)

    def hit_batch(self, origins, directions, max_distance=None, geometry=False):
        return self.build_accelerator().hit_batch(origins, directions, max_distance, geometry=geometry)

    def contains_batch(self, points, cap=8):
        return self.build_accelerator().contains_batch(points, cap)
