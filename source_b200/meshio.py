"""Mesh file importers with the reference's call signatures (raysect/primitive/mesh/obj.py, stl.py, ply.py, vtk.py).

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


PLY_AUTOMATIC, PLY_ASCII, PLY_BINARY = "auto", "ascii", "binary"
_PLY_SCALARS = {"char": "b", "int8": "b", "uchar": "B", "uint8": "B", "short": "h", "int16": "h", "ushort": "H", "uint16": "H",
                "int": "i", "int32": "i", "uint": "I", "uint32": "I", "float": "f", "float32": "f", "double": "d", "float64": "d"}


def import_ply(filename, scaling=1.0, mode=PLY_AUTOMATIC, **kwargs):
    """PLYHandler.import_ply (raysect/primitive/mesh/ply.py:47-200): vertices scaled by ``scaling`` in double precision,
    triangles as written, smoothing off.  ``mode`` is kept for the signature; the header says which format the file has.
    Reads what the reference reads and writes -- ascii and binary little-endian, float x / y / z, faces as
    ``list uchar int vertex_index(es)`` -- and, unlike it, any other scalar properties on the vertices (skipped), uint / int
    index types and big-endian files.  Faces that are not triangles are refused, as in the reference."""
    mode = mode.lower()
    if mode not in (PLY_AUTOMATIC, PLY_ASCII, PLY_BINARY):
        raise ValueError("Unrecognised import mode, valid values are: {}".format((PLY_AUTOMATIC, PLY_ASCII, PLY_BINARY)))
    with open(filename, "rb") as f:
        blob = f.read()
    end = blob.find(b"end_header")
    if not blob.startswith(b"ply") or end < 0:
        raise ValueError("This file is not a valid PLY file.")
    body = blob[blob.index(b"\n", end) + 1:]
    fmt, elements = None, []
    for raw in blob[:end].decode("ascii", "replace").splitlines()[1:]:
        words = raw.split()
        if not words or words[0] in ("comment", "obj_info"):
            continue
        if words[0] == "format":
            fmt = words[1]
        elif words[0] == "element":
            elements.append((words[1], int(words[2]), []))
        elif words[0] == "property" and elements:
            elements[-1][2].append(words[1:])
    if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
        raise ValueError("This file is not a valid PLY file.")
    if (mode == PLY_ASCII) != (fmt == "ascii") and mode != PLY_AUTOMATIC:
        raise ValueError("This file is not a valid PLY file.")
    vertices = triangles = None
    tokens = iter(body.split()) if fmt == "ascii" else None
    endian, at = "<" if fmt != "binary_big_endian" else ">", 0

    def scalar(kind):
        nonlocal at
        if tokens is not None:
            return float(next(tokens)) if kind in ("f", "d") else int(next(tokens))
        (v,) = struct.unpack_from(endian + kind, body, at)
        at += struct.calcsize(kind)
        return v
    for name, count, props in elements:
        kinds = []
        for p in props:
            if p[0] == "list":
                kinds.append(("list", _PLY_SCALARS[p[1]], _PLY_SCALARS[p[2]], p[3]))
            else:
                kinds.append(("scalar", _PLY_SCALARS[p[0]], None, p[1]))
        if name == "vertex":
            cols = {k[3]: i for i, k in enumerate(kinds)}
            if not all(c in cols and kinds[cols[c]][0] == "scalar" for c in "xyz"):
                raise ValueError("This file is not a valid PLY file.")
            vertices = np.empty((count, 3))
        elif name == "face":
            triangles = np.empty((count, 3), dtype=np.int32)
        for i in range(count):
            row = []
            for k in kinds:
                if k[0] == "scalar":
                    row.append(scalar(k[1]))
                else:
                    row.append([scalar(k[2]) for _ in range(scalar(k[1]))])
            if name == "vertex":
                vertices[i] = [row[cols[c]] for c in "xyz"]
            elif name == "face":
                corners = next((r for r, k in zip(row, kinds) if k[0] == "list" and k[3] in ("vertex_index", "vertex_indices")), None)
                if corners is None:
                    raise ValueError("This file is not a valid PLY file.")
                if len(corners) != 3:
                    raise ValueError("Raysect meshes can only handle triangles.")
                triangles[i] = corners
    if vertices is None or triangles is None:
        raise ValueError("This file is not a valid PLY file.")
    vertices *= scaling
    kwargs.setdefault("smoothing", False)
    return Mesh(vertices, triangles, **kwargs)


VTK_AUTOMATIC, VTK_ASCII, VTK_BINARY = "auto", "ascii", "binary"


def import_vtk(filename, scaling=1.0, mode=VTK_AUTOMATIC, **kwargs):
    """VTKHandler.import_vtk (raysect/primitive/mesh/vtk.py:49-140): legacy ASCII "DataFile Version 2.0" files holding an
    UNSTRUCTURED_GRID of triangular cells (cell type 5); vertices scaled in double precision, smoothing off, the mesh named
    after the file's title line unless ``name`` is given.  Binary .vtk files are not read (the reference does not read them
    either)."""
    mode = mode.lower()
    if mode == VTK_BINARY:
        raise NotImplementedError("The binary .vtk loading routine has not been implemented yet.")
    if mode not in (VTK_AUTOMATIC, VTK_ASCII):
        raise ValueError("Unrecognised import mode, valid values are: {}".format((VTK_ASCII, VTK_BINARY)))
    with open(filename, "r") as f:
        lines = [ln.strip() for ln in f]
    if len(lines) < 5 or lines[0] != "# vtk DataFile Version 2.0" or lines[2] != "ASCII":
        if mode == VTK_AUTOMATIC:
            raise NotImplementedError("The binary .vtk loading routine has not been implemented yet.")
        raise ValueError("This file is not a valid ASCII VTK file.")
    if lines[3] != "DATASET UNSTRUCTURED_GRID":
        raise RuntimeError("Unrecognised dataset encountered in vtk file.")
    words = lines[4].split()
    if len(words) != 3 or words[0] != "POINTS" or words[2] != "float":
        raise RuntimeError("Unrecognised dataset encountered in vtk file.")
    n_points = int(words[1])
    vertices = np.array([[float(c) * scaling for c in lines[5 + i].split()[:3]] for i in range(n_points)]).reshape(n_points, 3)
    at = 5 + n_points
    words = lines[at].split()
    if not words or words[0] != "CELLS":
        raise RuntimeError("Unrecognised dataset encountered in vtk file.")
    n_cells = int(words[1])
    triangles = np.array([[int(c) for c in lines[at + 1 + i].split()[1:4]] for i in range(n_cells)], dtype=np.int32).reshape(n_cells, 3)
    at += 1 + n_cells
    if lines[at].split()[0] != "CELL_TYPES" or any(int(lines[at + 1 + i]) != 5 for i in range(n_cells)):
        raise ValueError("Raysect meshes can only handle triangles.")
    kwargs.setdefault("name", lines[1] or "VTKMesh")
    kwargs.setdefault("smoothing", False)
    return Mesh(vertices, triangles, **kwargs)
