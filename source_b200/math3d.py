"""Host-side 3D maths mirroring the small part of ``raysect.core.math`` the scene front-end needs.

Python floats are IEEE doubles and CPython never fuses a multiply-add, so each method below, written
in the reference's operation order, produces the same bits as the reference's Cython (gcc -O2, no
FMA).  File:line citations point at the reference.
"""
import math

DEG2RAD = 0.017453292519943295   # raysect/core/math/transform.pyx:39


class Point3D:
    """raysect/core/math/point.pyx (Point3D)."""
    __slots__ = ("x", "y", "z")

    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = float(x), float(y), float(z)

    def __repr__(self):
        return "Point3D(%r, %r, %r)" % (self.x, self.y, self.z)

    def __iter__(self):
        return iter((self.x, self.y, self.z))

    def copy(self):
        return Point3D(self.x, self.y, self.z)

    def transform(self, m):
        """point.pyx:253-281"""
        a = m.m
        w = a[3][0] * self.x + a[3][1] * self.y + a[3][2] * self.z + a[3][3]
        if w == 0.0:
            raise ZeroDivisionError("Bad matrix transform, 4th element of homogeneous coordinate is zero.")
        w = 1.0 / w
        return Point3D((a[0][0] * self.x + a[0][1] * self.y + a[0][2] * self.z + a[0][3]) * w,
                       (a[1][0] * self.x + a[1][1] * self.y + a[1][2] * self.z + a[1][3]) * w,
                       (a[2][0] * self.x + a[2][1] * self.y + a[2][2] * self.z + a[2][3]) * w)

    def vector_to(self, p):
        """point.pyx:229"""
        return Vector3D(p.x - self.x, p.y - self.y, p.z - self.z)

    def distance_to(self, p):
        x, y, z = p.x - self.x, p.y - self.y, p.z - self.z
        return math.sqrt(x * x + y * y + z * z)


class Vector3D:
    """raysect/core/math/vector.pyx (Vector3D)."""
    __slots__ = ("x", "y", "z")

    def __init__(self, x=0.0, y=0.0, z=1.0):
        self.x, self.y, self.z = float(x), float(y), float(z)

    def __repr__(self):
        return "Vector3D(%r, %r, %r)" % (self.x, self.y, self.z)

    def __iter__(self):
        return iter((self.x, self.y, self.z))

    @property
    def length(self):
        return math.sqrt(self.x * self.x + self.y * self.y + self.z * self.z)

    def dot(self, v):
        return self.x * v.x + self.y * v.y + self.z * v.z

    def cross(self, v):
        """vector.pyx:306-310"""
        return Vector3D(self.y * v.z - v.y * self.z, self.z * v.x - v.z * self.x, self.x * v.y - v.x * self.y)

    def normalise(self):
        """vector.pyx:313-337"""
        t = self.x * self.x + self.y * self.y + self.z * self.z
        if t == 0.0:
            raise ZeroDivisionError("A zero length vector can not be normalised as the direction of a zero length vector is undefined.")
        t = 1.0 / math.sqrt(t)
        return Vector3D(self.x * t, self.y * t, self.z * t)

    def transform(self, m):
        """vector.pyx:339-366"""
        a = m.m
        return Vector3D(a[0][0] * self.x + a[0][1] * self.y + a[0][2] * self.z,
                        a[1][0] * self.x + a[1][1] * self.y + a[1][2] * self.z,
                        a[2][0] * self.x + a[2][1] * self.y + a[2][2] * self.z)


Normal3D = Vector3D


class AffineMatrix3D:
    """raysect/core/math/affinematrix.pyx (AffineMatrix3D): 4x4, row major, indexable m[i, j]."""
    __slots__ = ("m",)

    def __init__(self, m=None):
        if m is None:
            self.m = [[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]]
        else:
            self.m = [[float(v) for v in row] for row in m]
            if len(self.m) != 4 or any(len(r) != 4 for r in self.m):
                raise TypeError("AffineMatrix3D must be initialised with a 4x4 indexable object.")

    def __repr__(self):
        return "AffineMatrix3D(%r)" % (self.m,)

    def __getitem__(self, ij):
        try:
            i, j = ij[0], ij[1]
        except Exception:
            raise IndexError("Index must be a tuple containing (at least) the row and column indicies e.g. matrix[1,3].")
        if not (0 <= i < 4 and 0 <= j < 4):
            raise IndexError("Row or column index out of range [0, 3].")
        return self.m[i][j]

    def mul(self, other):
        """affinematrix.pyx:254-273: every element a 4-term sum, left to right."""
        a, b = self.m, other.m
        return AffineMatrix3D([[a[r][0] * b[0][c] + a[r][1] * b[1][c] + a[r][2] * b[2][c] + a[r][3] * b[3][c]
                                for c in range(4)] for r in range(4)])

    def __mul__(self, other):
        if isinstance(other, AffineMatrix3D):
            return self.mul(other)
        return NotImplemented

    def inverse(self):
        """affinematrix.pyx:174-252 (Cramer's rule, same temporaries)."""
        m = self.m
        t = [0.0] * 22
        t[0] = m[0][0] * m[1][1] - m[0][1] * m[1][0]
        t[1] = m[0][0] * m[1][2] - m[0][2] * m[1][0]
        t[2] = m[0][0] * m[1][3] - m[0][3] * m[1][0]
        t[3] = m[0][1] * m[1][2] - m[0][2] * m[1][1]
        t[4] = m[0][1] * m[1][3] - m[0][3] * m[1][1]
        t[5] = m[0][2] * m[1][3] - m[0][3] * m[1][2]
        t[18] = m[2][0] * t[3] - m[2][1] * t[1] + m[2][2] * t[0]
        t[19] = m[2][0] * t[4] - m[2][1] * t[2] + m[2][3] * t[0]
        t[20] = m[2][0] * t[5] - m[2][2] * t[2] + m[2][3] * t[1]
        t[21] = m[2][1] * t[5] - m[2][2] * t[4] + m[2][3] * t[3]
        det = t[20] * m[3][1] + t[18] * m[3][3] - t[21] * m[3][0] - t[19] * m[3][2]
        if abs(det) < 1e-14:
            raise ValueError("Matrix is singular and not invertible.")
        idet = 1.0 / det
        t[6] = m[0][0] * m[3][1] - m[0][1] * m[3][0]
        t[7] = m[0][0] * m[3][2] - m[0][2] * m[3][0]
        t[8] = m[0][0] * m[3][3] - m[0][3] * m[3][0]
        t[9] = m[0][1] * m[3][2] - m[0][2] * m[3][1]
        t[10] = m[0][1] * m[3][3] - m[0][3] * m[3][1]
        t[11] = m[0][2] * m[3][3] - m[0][3] * m[3][2]
        t[12] = m[1][0] * m[3][1] - m[1][1] * m[3][0]
        t[13] = m[1][0] * m[3][2] - m[1][2] * m[3][0]
        t[14] = m[1][0] * m[3][3] - m[1][3] * m[3][0]
        t[15] = m[1][1] * m[3][2] - m[1][2] * m[3][1]
        t[16] = m[1][1] * m[3][3] - m[1][3] * m[3][1]
        t[17] = m[1][2] * m[3][3] - m[1][3] * m[3][2]
        return AffineMatrix3D([
            [(m[2][2] * t[16] - m[2][1] * t[17] - m[2][3] * t[15]) * idet,
             (m[2][1] * t[11] - m[2][2] * t[10] + m[2][3] * t[9]) * idet,
             (m[3][1] * t[5] - m[3][2] * t[4] + m[3][3] * t[3]) * idet,
             -t[21] * idet],
            [(m[2][0] * t[17] - m[2][2] * t[14] + m[2][3] * t[13]) * idet,
             (m[2][2] * t[8] - m[2][0] * t[11] - m[2][3] * t[7]) * idet,
             (m[3][2] * t[2] - m[3][0] * t[5] - m[3][3] * t[1]) * idet,
             t[20] * idet],
            [(m[2][1] * t[14] - m[2][0] * t[16] - m[2][3] * t[12]) * idet,
             (m[2][0] * t[10] - m[2][1] * t[8] + m[2][3] * t[6]) * idet,
             (m[3][0] * t[4] - m[3][1] * t[2] + m[3][3] * t[0]) * idet,
             -t[19] * idet],
            [(m[2][0] * t[15] - m[2][1] * t[13] + m[2][2] * t[12]) * idet,
             (m[2][1] * t[7] - m[2][0] * t[9] - m[2][2] * t[6]) * idet,
             (m[3][1] * t[1] - m[3][0] * t[3] - m[3][2] * t[0]) * idet,
             t[18] * idet]])


def translate(x, y, z):
    """raysect/core/math/transform.pyx:40-68"""
    return AffineMatrix3D([[1, 0, 0, x], [0, 1, 0, y], [0, 0, 1, z], [0, 0, 0, 1]])


def rotate_x(angle):
    """transform.pyx:70-112"""
    r = DEG2RAD * angle
    return AffineMatrix3D([[1, 0, 0, 0], [0, math.cos(r), -math.sin(r), 0], [0, math.sin(r), math.cos(r), 0], [0, 0, 0, 1]])


def rotate_y(angle):
    """transform.pyx:114-140"""
    r = DEG2RAD * angle
    return AffineMatrix3D([[math.cos(r), 0, math.sin(r), 0], [0, 1, 0, 0], [-math.sin(r), 0, math.cos(r), 0], [0, 0, 0, 1]])


def rotate_z(angle):
    """transform.pyx:142-168"""
    r = DEG2RAD * angle
    return AffineMatrix3D([[math.cos(r), -math.sin(r), 0, 0], [math.sin(r), math.cos(r), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])


def rotate(yaw, pitch, roll):
    """transform.pyx:215-231: intrinsic (-Y)(-X)'Z''"""
    return rotate_y(-yaw) * rotate_x(-pitch) * rotate_z(roll)


class BoundingBox3D:
    """raysect/core/boundingbox.pyx (BoundingBox3D): the handful of methods scene construction uses."""
    __slots__ = ("lower", "upper")

    def __init__(self, lower=None, upper=None):
        if lower is None or upper is None:
            self.lower = Point3D(math.inf, math.inf, math.inf)
            self.upper = Point3D(-math.inf, -math.inf, -math.inf)
        else:
            if lower.x > upper.x or lower.y > upper.y or lower.z > upper.z:
                raise ValueError("The lower point coordinates must be less than or equal to the upper point coordinates.")
            self.lower, self.upper = lower, upper

    def __repr__(self):
        return "BoundingBox3D(%r, %r)" % (self.lower, self.upper)

    def union(self, box):
        """boundingbox.pyx:265-281"""
        self.lower = Point3D(min(self.lower.x, box.lower.x), min(self.lower.y, box.lower.y), min(self.lower.z, box.lower.z))
        self.upper = Point3D(max(self.upper.x, box.upper.x), max(self.upper.y, box.upper.y), max(self.upper.z, box.upper.z))

    def extend(self, point, padding=0.0):
        """boundingbox.pyx:283-300"""
        self.lower = Point3D(min(self.lower.x, point.x - padding), min(self.lower.y, point.y - padding), min(self.lower.z, point.z - padding))
        self.upper = Point3D(max(self.upper.x, point.x + padding), max(self.upper.y, point.y + padding), max(self.upper.z, point.z + padding))

    def vertices(self):
        """boundingbox.pyx:326-343"""
        l, u = self.lower, self.upper
        return [Point3D(l.x, l.y, l.z), Point3D(l.x, l.y, u.z), Point3D(l.x, u.y, l.z), Point3D(l.x, u.y, u.z),
                Point3D(u.x, l.y, l.z), Point3D(u.x, l.y, u.z), Point3D(u.x, u.y, l.z), Point3D(u.x, u.y, u.z)]

    def get_centre(self):
        """boundingbox.pyx:138-144"""
        return Point3D(0.5 * (self.lower.x + self.upper.x), 0.5 * (self.lower.y + self.upper.y), 0.5 * (self.lower.z + self.upper.z))

    def enclosing_sphere(self):
        """boundingbox.pyx:444-457 (SPHERE_PADDING = 1.000001, :47)"""
        centre = self.get_centre()
        return BoundingSphere3D(centre, self.lower.distance_to(centre) * 1.000001)


class BoundingSphere3D:
    """raysect/core/boundingsphere.pyx:39-60"""
    __slots__ = ("centre", "radius")

    def __init__(self, centre, radius):
        if radius <= 0:
            raise ValueError("The radius of the bounding sphere must be greater than zero.")
        self.centre, self.radius = centre, radius
