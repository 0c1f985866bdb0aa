"""Mesh ingestion for the scenes the reference's drivers load: Wavefront OBJ and Stanford PLY, with the
semantics of the reference's loaders (`src/shapes/obj.cpp:150-410`, `src/shapes/ply.cpp:130-440`) — what ends up in
`vertex_positions`, `vertex_normals` and `faces`, because those arrays are what the path differentiates:

* OBJ: a vertex of the mesh is a distinct `(v, vt, vn)` index triple, numbered in order of first appearance
  (`obj.cpp:286-321`); polygons are fanned `(0, k−1, k)` (`obj.cpp:323-333`); `to_world` is applied to positions and — as the
  reference does — to normals with the same affine transform followed by normalisation (`obj.cpp:236-253`); texture
  coordinates are read and their v flipped (`flip_tex_coords`, default true).
* PLY: ascii, binary_little_endian and binary_big_endian; the `vertex` element's `x y z` (+ `nx ny nz`, + `u v` / `s t`) and the
  `face` element's `vertex_index` / `vertex_indices` list; triangle meshes only (`ply.cpp:395`: "is this a triangle mesh?").
* Both: without normals in the file (and `face_normals=False`) vertex normals are recomputed angle-weighted
  (`Mesh::recompute_vertex_normals`, `mesh.cpp:257-330`); with `face_normals=True` the mesh is flat-shaded (no normals).

Host-side IO only: returns `scenes.Mesh` objects for `SceneData.add_mesh`."""
from __future__ import annotations

import struct
from typing import Optional

import numpy as np

from .scenes import Mesh, Transform4f, vertex_normals


class MeshIOError(RuntimeError):
    pass


def _apply(to_world: Optional[Transform4f], positions: np.ndarray, normals: Optional[np.ndarray]):
    if to_world is None:
        return positions, normals
    m = np.asarray(to_world.matrix, np.float64)
    p = (positions.astype(np.float64) @ m[:3, :3].T + m[:3, 3]).astype(np.float32)
    n = None
    if normals is not None:
        # transform_affine(InputNormal3f) dispatches to the Normal overload: inverse_transpose · n, no translation
        # (include/mitsuba/core/transform.h:98-100,146-155; obj.cpp:253, ply.cpp:298)
        n = normals.astype(np.float64) @ np.linalg.inv(m[:3, :3])
        ln = np.linalg.norm(n, axis=1, keepdims=True); ln[ln == 0] = 1
        n = (n / ln).astype(np.float32)
    return p, n


def _finish(name, positions, faces, normals, face_normals, kw):
    if face_normals:
        normals = None
    elif normals is None:
        normals = vertex_normals(positions, faces)
    return Mesh(name=name, positions=positions, faces=faces, normals=normals, **kw)


def load_obj(path: str, to_world: Optional[Transform4f] = None, face_normals: bool = False, flip_tex_coords: bool = True,
             name: Optional[str] = None, return_texcoords: bool = False, **mesh_kw):
    vs, vns, vts, tris = [], [], [], []
    binding = {}                                                          # (v, vt, vn) → vertex id, in order of first appearance
    try:
        fh = open(path, "r")
    except OSError:
        raise MeshIOError(f'Error while loading OBJ file "{path}": file not found')
    with fh:
        for line in fh:
            t = line.split()
            if not t:
                continue
            try:
                if t[0] == "v":
                    vs.append([float(t[1]), float(t[2]), float(t[3])])
                elif t[0] == "vn":
                    if not face_normals:
                        vns.append([float(t[1]), float(t[2]), float(t[3])])
                elif t[0] == "vt":
                    uv = [float(t[1]), float(t[2])]
                    if flip_tex_coords:
                        uv[1] = 1.0 - uv[1]
                    vts.append(uv)
                elif t[0] == "f":
                    ids = []
                    for tok in t[1:]:
                        parts = tok.split("/")
                        if len(parts) > 3:
                            raise ValueError
                        key = tuple(int(p) if p else 0 for p in parts) + (0,) * (3 - len(parts))
                        if key[0] < 1 or key[0] > len(vs):
                            raise MeshIOError(f'Error while loading OBJ file "{path}": reference to invalid vertex {key[0]}!')
                        if key not in binding:
                            binding[key] = len(binding)
                        ids.append(binding[key])
                    for k in range(2, len(ids)):
                        tris.append([ids[0], ids[k - 1], ids[k]])
            except (ValueError, IndexError):
                raise MeshIOError(f'Error while loading OBJ file "{path}": could not parse line "{line.strip()}"')
    nv = len(binding)
    positions = np.zeros((nv, 3), np.float32); normals = np.zeros((nv, 3), np.float32); uvs = np.zeros((nv, 2), np.float32)
    vs_a = np.asarray(vs, np.float32).reshape(-1, 3); vn_a = np.asarray(vns, np.float32).reshape(-1, 3); vt_a = np.asarray(vts, np.float32).reshape(-1, 2)
    for (v, vt, vn), i in binding.items():
        positions[i] = vs_a[v - 1]
        if vt:
            if vt > len(vt_a):
                raise MeshIOError(f'Error while loading OBJ file "{path}": reference to invalid texture coordinate {vt}!')
            uvs[i] = vt_a[vt - 1]
        if not face_normals and vn:
            if vn > len(vn_a):
                raise MeshIOError(f'Error while loading OBJ file "{path}": reference to invalid normal {vn}!')
            normals[i] = vn_a[vn - 1]
    has_normals = len(vn_a) > 0 and not face_normals
    positions, n_t = _apply(to_world, positions, normals if has_normals else None)
    faces = np.asarray(tris, np.uint32).reshape(-1, 3)
    mesh = _finish(name or path.split("/")[-1], positions, faces, n_t, face_normals, mesh_kw)
    return (mesh, uvs if len(vt_a) else None) if return_texcoords else mesh


_PLY_TYPES = {"char": "b", "int8": "b", "uchar": "B", "uint8": "B", "short": "h", "int16": "h", "ushort": "H", "uint16": "H",
              "int": "i", "int32": "i", "uint": "I", "uint32": "I", "float": "f", "float32": "f", "double": "d", "float64": "d"}


def load_ply(path: str, to_world: Optional[Transform4f] = None, face_normals: bool = False, flip_tex_coords: bool = True,
             name: Optional[str] = None, return_texcoords: bool = False, **mesh_kw):
    def fail(msg):
        raise MeshIOError(f'Error while loading PLY file "{path}": {msg}!')
    try:
        data = open(path, "rb").read()
    except OSError:
        fail("file not found")
    end = data.find(b"end_header")
    if not data.startswith(b"ply") or end < 0:
        fail("invalid PLY header")
    header = data[:end].decode("ascii", "replace").splitlines()
    body = data[data.index(b"\n", end) + 1:]
    fmt, elements = None, []
    for line in header[1:]:
        t = line.split()
        if not t or t[0] in ("comment", "obj_info"):
            continue
        if t[0] == "format":
            fmt = t[1]
        elif t[0] == "element":
            elements.append({"name": t[1], "count": int(t[2]), "props": []})
        elif t[0] == "property":
            if not elements:
                fail("invalid PLY header: property before element")
            if t[1] == "list":
                if len(t) < 5:
                    fail('invalid PLY header: missing token after "property list"')
                elements[-1]["props"].append(("list", t[2], t[3], t[4]))
            else:
                elements[-1]["props"].append(("scalar", t[1], t[2]))
    if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
        fail("invalid PLY header: unknown format")
    for el in elements:
        for p in el["props"]:
            for ty in p[1:-1]:
                if ty not in _PLY_TYPES:
                    fail(f'invalid PLY header: unknown type "{ty}"')
    positions = normals = uvs = faces = None
    if fmt == "ascii":
        tokens = body.split(); pos = 0
    else:
        endian = "<" if fmt == "binary_little_endian" else ">"; off = 0
    for el in elements:
        rows = []
        scalar_only = all(p[0] == "scalar" for p in el["props"])
        if fmt != "ascii" and scalar_only:
            dt = np.dtype([(p[2], endian + _PLY_TYPES[p[1]]) for p in el["props"]])
            arr = np.frombuffer(body, dt, el["count"], off); off += dt.itemsize * el["count"]
            cols = {p[2]: arr[p[2]] for p in el["props"]}
        else:
            cols = {p[-1]: [] for p in el["props"]}
            for _ in range(el["count"]):
                for p in el["props"]:
                    if p[0] == "scalar":
                        if fmt == "ascii":
                            v = float(tokens[pos]); pos += 1
                        else:
                            (v,) = struct.unpack_from(endian + _PLY_TYPES[p[1]], body, off); off += struct.calcsize(_PLY_TYPES[p[1]])
                        cols[p[2]].append(v)
                    else:
                        if fmt == "ascii":
                            n = int(tokens[pos]); vals = [int(float(x)) for x in tokens[pos + 1:pos + 1 + n]]; pos += 1 + n
                        else:
                            (n,) = struct.unpack_from(endian + _PLY_TYPES[p[1]], body, off); off += struct.calcsize(_PLY_TYPES[p[1]])
                            vals = list(struct.unpack_from(endian + str(n) + _PLY_TYPES[p[2]], body, off)); off += n * struct.calcsize(_PLY_TYPES[p[2]])
                        cols[p[3]].append(vals)
        if el["name"] == "vertex":
            if not all(k in cols for k in ("x", "y", "z")):
                fail("vertex element without x, y, z")
            positions = np.stack([np.asarray(cols[k], np.float32) for k in ("x", "y", "z")], 1)
            if all(k in cols for k in ("nx", "ny", "nz")) and not face_normals:
                normals = np.stack([np.asarray(cols[k], np.float32) for k in ("nx", "ny", "nz")], 1)
            for a, b in (("u", "v"), ("s", "t"), ("texture_u", "texture_v"), ("texture_s", "texture_t")):
                if a in cols and b in cols:
                    uvs = np.stack([np.asarray(cols[a], np.float32), np.asarray(cols[b], np.float32)], 1)
                    if flip_tex_coords:
                        uvs[:, 1] = 1.0 - uvs[:, 1]
                    break
        elif el["name"] == "face":
            key = "vertex_index" if "vertex_index" in cols else ("vertex_indices" if "vertex_indices" in cols else None)
            if key is None:
                fail("vertex_index/vertex_indices property not found")
            if any(len(f) != 3 for f in cols[key]):
                fail("incompatible contents -- is this a triangle mesh?")
            faces = np.asarray(cols[key], np.uint32).reshape(-1, 3)
    if fmt == "ascii":
        if pos != len(tokens):
            fail("invalid file -- trailing content")
    elif off != len(body):
        fail("invalid file -- trailing content")
    if positions is None or faces is None:
        fail("vertex or face element missing")
    if faces.size and faces.max() >= positions.shape[0]:
        fail("face references a vertex out of range")
    positions, normals = _apply(to_world, positions, normals)
    mesh = _finish(name or path.split("/")[-1], positions, faces, normals, face_normals, mesh_kw)
    return (mesh, uvs) if return_texcoords else mesh


def save_ply(path: str, mesh: Mesh, binary: bool = True):
    """Writer used by the tests and for exporting the synthetic scenes (vertex x y z [nx ny nz], triangle faces)."""
    has_n = mesh.normals is not None
    has_uv = getattr(mesh, "texcoords", None) is not None
    hdr = ["ply", "format binary_little_endian 1.0" if binary else "format ascii 1.0", f"element vertex {mesh.positions.shape[0]}",
           "property float x", "property float y", "property float z"]
    if has_n:
        hdr += ["property float nx", "property float ny", "property float nz"]
    if has_uv:
        hdr += ["property float u", "property float v"]
    hdr += [f"element face {mesh.faces.shape[0]}", "property list uchar int vertex_indices", "end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode())
        v = np.concatenate([mesh.positions, mesh.normals], 1) if has_n else mesh.positions
        if has_uv:
            v = np.concatenate([v, np.asarray(mesh.texcoords, np.float32).reshape(-1, 2)], 1)
        if binary:
            f.write(np.ascontiguousarray(v, "<f4").tobytes())
            rec = np.zeros(mesh.faces.shape[0], np.dtype([("n", "u1"), ("i", "<i4", 3)]))
            rec["n"] = 3; rec["i"] = mesh.faces.astype(np.int32)
            f.write(rec.tobytes())
        else:
            for row in v:
                f.write((" ".join(repr(float(x)) for x in row) + "\n").encode())
            for t in mesh.faces:
                f.write(f"3 {t[0]} {t[1]} {t[2]}\n".encode())


# ---------------------------------------------------------------------------------------------------------------------
#  Mitsuba `.serialized` meshes (src/shapes/serialized.cpp:198-400): header 0x041C, version 3 | 4, then one deflate
#  stream per sub-mesh: flags u32, [name\0 (v4)], vertex_count u64, face_count u64, positions, [normals], [texcoords],
#  [colours] (float32 or float64 per flags), uint32 indices.  Multi-mesh files end with the sub-stream offsets
#  (u64 for v4, u32 for v3) and the u32 mesh count.
# ---------------------------------------------------------------------------------------------------------------------
_SER_HEADER, _SER_V3, _SER_V4 = 0x041C, 3, 4
_SER_NORMALS, _SER_TEXCOORDS, _SER_COLORS, _SER_FACENORMALS, _SER_SINGLE, _SER_DOUBLE = 0x1, 0x2, 0x8, 0x10, 0x1000, 0x2000


def load_serialized(path: str, shape_index: int = 0, to_world: Optional[Transform4f] = None, face_normals: bool = False,
                    name: Optional[str] = None, return_texcoords: bool = False, **mesh_kw):
    import struct
    import zlib

    def fail(msg):
        raise MeshIOError(f'Error while loading serialized file "{path}": {msg}!')
    try:
        data = open(path, "rb").read()
    except OSError:
        fail("file not found")
    if len(data) < 4:
        fail("encountered an invalid file format")
    fmt, version = struct.unpack_from("<HH", data, 0)
    if fmt != _SER_HEADER:
        fail("encountered an invalid file format")
    if version not in (_SER_V3, _SER_V4):
        fail("encountered an incompatible file version")
    start = 4
    if shape_index != 0:
        (count,) = struct.unpack_from("<I", data, len(data) - 4)
        if shape_index < 0 or shape_index >= count:
            fail(f"Unable to unserialize mesh, shape index is out of range! (requested {shape_index} out of 0..{count - 1})")
        if version == _SER_V4:
            (off,) = struct.unpack_from("<Q", data, len(data) - 8 * (count - shape_index) - 4)
        else:
            (off,) = struct.unpack_from("<I", data, len(data) - 4 * (count - shape_index + 1))
        start = off + 4                                                   # skip the sub-stream's own header
    try:
        body = zlib.decompressobj().decompress(data[start:])
    except zlib.error as e:
        fail(f"zlib: {e}")
    pos = 0
    (flags,) = struct.unpack_from("<I", body, pos); pos += 4
    ser_name = ""
    if version == _SER_V4:
        end = body.index(b"\0", pos); ser_name = body[pos:end].decode("utf-8", "replace"); pos = end + 1
    nv, nf = struct.unpack_from("<QQ", body, pos); pos += 16
    dt = "<f8" if flags & _SER_DOUBLE else "<f4"
    isz = 8 if flags & _SER_DOUBLE else 4

    def block(dim):
        nonlocal pos
        a = np.frombuffer(body, dt, nv * dim, pos).astype(np.float32).reshape(nv, dim); pos += nv * dim * isz
        return a
    positions = block(3)
    normals = block(3) if flags & _SER_NORMALS else None
    uvs = block(2) if flags & _SER_TEXCOORDS else None
    if flags & _SER_COLORS:
        block(3)
    faces = np.frombuffer(body, "<u4", nf * 3, pos).reshape(nf, 3).copy()
    if faces.size and faces.max() >= nv:
        fail("face references a vertex out of range")
    if face_normals:
        normals = None
    positions, normals = _apply(to_world, positions, normals)
    mesh = _finish(name or ser_name or f"{path.split('/')[-1]}@{shape_index}", positions, faces, normals, face_normals, mesh_kw)
    return (mesh, uvs) if return_texcoords else mesh


def save_serialized(path: str, meshes, version: int = _SER_V4, double: bool = False):
    """Writer for tests / exports: one or several meshes into one `.serialized` file."""
    import struct
    import zlib
    meshes = [meshes] if isinstance(meshes, Mesh) else list(meshes)
    out = bytearray(); offsets = []
    for m in meshes:
        offsets.append(len(out))
        out += struct.pack("<HH", _SER_HEADER, version)
        flags = (_SER_DOUBLE if double else _SER_SINGLE) | (_SER_NORMALS if m.normals is not None else 0)
        raw = bytearray(struct.pack("<I", flags))
        if version == _SER_V4:
            raw += m.name.encode() + b"\0"
        raw += struct.pack("<QQ", m.positions.shape[0], m.faces.shape[0])
        dt = "<f8" if double else "<f4"
        raw += np.ascontiguousarray(m.positions, dt).tobytes()
        if m.normals is not None:
            raw += np.ascontiguousarray(m.normals, dt).tobytes()
        raw += np.ascontiguousarray(m.faces, "<u4").tobytes()
        out += zlib.compress(bytes(raw))
    for o in offsets:
        out += struct.pack("<Q" if version == _SER_V4 else "<I", o)
    out += struct.pack("<I", len(meshes))
    with open(path, "wb") as f:
        f.write(bytes(out))
