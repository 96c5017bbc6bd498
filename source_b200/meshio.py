"""Mesh file importers with the reference's call signatures (raysect/primitive/mesh/obj.py, stl.py).

``import_obj(filename, scaling=1.0, **mesh_kwargs)`` and ``import_stl(filename, scaling=1.0, mode=..., **mesh_kwargs)``
return a ``source_b200.Mesh``; its kd-tree is built by this package's bit-exact SAH builder and the triangles go to
the device as pre-gathered 48-byte rows.
"""
import struct

import numpy as np

from .math3d import Normal3D
from .scenegraph import Mesh


def _face_indices(token):
    """one corner of an OBJ face, ``v``, ``v/vt`` or ``v/vt/vn`` (1-based in the file): (vertex, normal or None)"""
    parts = token.split("/")
    if not 1 <= len(parts) <= 3:
        raise ValueError("The .obj contains an invalid face definition.")
    return int(parts[0]) - 1, (int(parts[2]) - 1 if len(parts) == 3 else None)


def import_obj(filename, scaling=1.0, **kwargs):
    """Wavefront OBJ -> Mesh with the semantics of the reference importer (raysect/primitive/mesh/obj.py:38-142):
    ``v`` records scaled by ``scaling``, ``vn`` records normalised in double precision, ``vt`` ignored, triangular
    ``f`` records only; a face carries normal indices only when all three corners name one."""
    records = {"v": [], "vn": [], "f": []}
    with open(filename) as f:
        for line in f:
            fields = line.split()
            if fields and fields[0] in records:
                records[fields[0]].append(fields[1:])
    vertices = scaling * np.array(records["v"], dtype=np.float64).reshape(-1, 3)
    normals = []
    for x, y, z in records["vn"]:
        n = Normal3D(float(x), float(y), float(z)).normalise()
        normals.append([n.x, n.y, n.z])
    triangles = []
    for corners in records["f"]:
        if len(corners) != 3:
            raise ValueError("The .obj importer only support meshes containing 3 sided faces (triangles).")
        idx = [_face_indices(c) for c in corners]
        row = [v for v, _ in idx]
        if all(n is not None for _, n in idx):
            row += [n for _, n in idx]
        triangles.append(row)
    if normals:
        return Mesh(vertices, triangles, normals, **kwargs)
    return Mesh(vertices, triangles, **kwargs)


STL_AUTOMATIC, STL_ASCII, STL_BINARY = "auto", "ascii", "binary"


def import_stl(filename, scaling=1.0, mode=STL_AUTOMATIC, **kwargs):
    """STLHandler.import_stl (raysect/primitive/mesh/stl.py:46-180): every facet contributes three new vertices
    (no welding), facet normals are ignored (the mesh computes its own face normals), smoothing is off."""
    mode = mode.lower()
    if mode == STL_ASCII:
        vertices, triangles = _load_stl_ascii(filename, scaling)
    elif mode == STL_BINARY:
        vertices, triangles = _load_stl_binary(filename, scaling)
    elif mode == STL_AUTOMATIC:
        try:
            vertices, triangles = _load_stl_ascii(filename, scaling)
        except ValueError:
            vertices, triangles = _load_stl_binary(filename, scaling)
    else:
        raise ValueError("Unrecognised import mode specified: {}".format(mode))
    kwargs.setdefault("smoothing", False)
    return Mesh(vertices, triangles, **kwargs)


def _load_stl_ascii(filename, scaling):
    with open(filename, "r") as f:
        try:
            if not f.readline().startswith("solid"):
                raise ValueError("ASCII STL data does not start with 'solid'.")
        except UnicodeDecodeError:
            raise ValueError("File does not contain valid ascii data.")
        vertices, triangles = [], []
        facet = []
        try:
            for line in f:
                tokens = line.strip().split()
                if not tokens:
                    continue
                if tokens[0] == "vertex":
                    facet.append([scaling * float(tokens[1]), scaling * float(tokens[2]), scaling * float(tokens[3])])
                elif tokens[0] == "endfacet":
                    if len(facet) != 3:
                        raise ValueError("ASCII STL facet does not have three vertices.")
                    base = len(vertices)
                    vertices.extend(facet)
                    triangles.append([base, base + 1, base + 2])
                    facet = []
        except UnicodeDecodeError:
            raise ValueError("File does not contain valid ascii data.")
    if not triangles:
        raise ValueError("ASCII STL file contains no facets.")
    return vertices, triangles


def _load_stl_binary(filename, scaling):
    with open(filename, "rb") as f:
        f.seek(80)                                            # header
        count = struct.unpack("<I", f.read(4))[0]
        data = np.frombuffer(f.read(50 * count), dtype=np.uint8)
    if data.size != 50 * count:
        raise ValueError("Binary STL file is truncated.")
    rec = data.reshape(count, 50)[:, 12:48].copy().view("<f4").reshape(count, 3, 3)   # skip the facet normal
    vertices = (scaling * rec.astype(np.float64)).reshape(-1, 3)
    triangles = np.arange(3 * count, dtype=np.int32).reshape(count, 3)
    return vertices, triangles
