"""Mesh ingestion: OBJ / PLY with the vertex-binding, triangulation and normal semantics of the reference's
loaders (src/shapes/obj.cpp, src/shapes/ply.cpp).  The reference's own loader tests read files from the absent mitsuba-data
submodule; the cases below rebuild the situations those tests describe (src/shapes/tests/test_obj.py, test_ply.py)."""
import numpy as np
import pytest

from epsm_mitsuba3_b200 import meshio, scenes as S

CUBE_OBJ = """# two quads sharing an edge, one triangle with its own normals; negative-free, 1-based
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 1 0 1
v 1 1 1
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vn 0 0 1
vn 1 0 0
f 1/1/1 2/2/1 3/3/1 4/4/1
f 2/1/2 5/2/2 6/3/2 3/4/2
f 1 2 5
"""


def test_obj_vertex_binding_and_fan_triangulation(tmp_path):
    p = tmp_path / "cube.obj"; p.write_text(CUBE_OBJ)
    m, uv = meshio.load_obj(str(p), return_texcoords=True)
    # distinct (v, vt, vn) triples in order of first appearance (obj.cpp:286-321): 4 + 4 + 3 — vertex 2 appears with three keys
    assert m.positions.shape == (11, 3) and m.faces.shape == (5, 3)
    assert np.array_equal(m.faces[:2], [[0, 1, 2], [0, 2, 3]])               # quad (a b c d) → (a b c), (a c d)  (obj.cpp:323-333)
    assert np.array_equal(m.faces[2:4], [[4, 5, 6], [4, 6, 7]])
    assert np.array_equal(m.faces[4], [8, 9, 10])
    assert np.array_equal(m.positions[1], [1, 0, 0]) and np.array_equal(m.positions[4], [1, 0, 0]) and np.array_equal(m.positions[9], [1, 0, 0])
    assert np.array_equal(m.normals[0], [0, 0, 1]) and np.array_equal(m.normals[4], [1, 0, 0])
    assert np.array_equal(m.normals[8], [0, 0, 0])                           # a key without vn keeps a zero normal (obj.cpp:346-375)
    assert np.allclose(uv[2], [1, 0]) and np.allclose(uv[0], [0, 1])         # flip_tex_coords: v ↦ 1 − v (obj.cpp:262-263)


def test_obj_without_normals_gets_vertex_normals_and_face_normals_flag(tmp_path):
    p = tmp_path / "quad.obj"; p.write_text("v -1 -1 0\nv 1 -1 0\nv 1 1 0\nv -1 1 0\nf 1 2 3 4\n")
    m = meshio.load_obj(str(p))
    assert m.positions.shape == (4, 3) and np.allclose(m.normals, [[0, 0, 1]] * 4)       # recompute_vertex_normals (obj.cpp:401-406)
    flat = meshio.load_obj(str(p), face_normals=True)
    assert flat.normals is None
    t = meshio.load_obj(str(p), to_world=S.Transform4f().translate([0, 0, 2]).scale([2, 2, 1]))
    assert np.allclose(t.positions[2], [2, 2, 2])


def test_obj_errors(tmp_path):
    p = tmp_path / "bad.obj"; p.write_text("v 0 0 0\nf 1 2 3\n")
    with pytest.raises(meshio.MeshIOError, match="invalid vertex"):
        meshio.load_obj(str(p))
    p.write_text("v 0 zero 0\n")
    with pytest.raises(meshio.MeshIOError, match="could not parse"):
        meshio.load_obj(str(p))
    with pytest.raises(meshio.MeshIOError, match="file not found"):
        meshio.load_obj(str(tmp_path / "missing.obj"))


@pytest.mark.parametrize("binary", [True, False])
def test_ply_round_trip(tmp_path, binary):
    src = S.sphere_mesh((0.5, -1.0, 2.0), 1.5, 2, smooth=True, name="s")
    p = tmp_path / "s.ply"
    meshio.save_ply(str(p), src, binary=binary)
    m = meshio.load_ply(str(p))
    assert np.array_equal(m.faces, src.faces)
    assert np.array_equal(m.positions, src.positions) if binary else np.allclose(m.positions, src.positions, atol=0)
    assert np.allclose(m.normals, src.normals, atol=1e-7)
    flat = S.sphere_mesh((0, 0, 0), 1.0, 1, smooth=False, name="f")
    meshio.save_ply(str(p), flat, binary=binary)
    m2 = meshio.load_ply(str(p))
    assert m2.normals is not None and np.allclose(np.linalg.norm(m2.normals, axis=1), 1, atol=1e-5)      # ply.cpp:435-437
    assert meshio.load_ply(str(p), face_normals=True).normals is None


def test_ply_big_endian_texcoords_and_errors(tmp_path):
    p = tmp_path / "t.ply"
    hdr = ("ply\nformat binary_big_endian 1.0\ncomment made by hand\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
           "property float s\nproperty float t\nelement face 1\nproperty list uchar uint vertex_index\nend_header\n").encode()
    v = np.array([[0, 0, 0, 0, 0], [1, 0, 0, 1, 0], [0, 1, 0, 0, 0.25]], ">f4").tobytes()
    f = bytes([3]) + np.array([0, 1, 2], ">u4").tobytes()
    p.write_bytes(hdr + v + f)
    m, uv = meshio.load_ply(str(p), return_texcoords=True)
    assert np.array_equal(m.faces, [[0, 1, 2]]) and np.allclose(m.positions[2], [0, 1, 0]) and np.allclose(uv[2], [0, 0.75])
    p.write_bytes(hdr + v + bytes([4]) + np.array([0, 1, 2, 2], ">u4").tobytes())
    with pytest.raises(meshio.MeshIOError, match="is this a triangle mesh"):
        meshio.load_ply(str(p))
    p.write_bytes(hdr + v + f + b"x")
    with pytest.raises(meshio.MeshIOError, match="trailing content"):
        meshio.load_ply(str(p))


def test_loaded_mesh_feeds_the_scene_and_the_oracle(tmp_path, orc):
    """A PLY export of a scene mesh, loaded back, gives the same hits as the original arrays."""
    from tests import parity
    sd = S.mirror_plane_scene(16, 2, 8)
    p = tmp_path / "mirror.ply"
    meshio.save_ply(str(p), sd.meshes[0])
    sd2 = S.mirror_plane_scene(16, 2, 8)
    m = meshio.load_ply(str(p), name="mirror", bsdf=sd.meshes[0].bsdf, grad=True)
    sd2.meshes[0] = m
    rays = parity.random_rays(sd, 2000)
    assert np.array_equal(orc.OracleScene(sd).trace(rays).view(np.uint32), orc.OracleScene(sd2).trace(rays).view(np.uint32))


def test_file_normals_use_the_inverse_transpose_and_no_translation(tmp_path):
    """transform_affine on a Normal3f = inverse_transpose · n without the translation (transform.h:146-155; obj.cpp:253,
    ply.cpp:298): a translated, non-uniformly scaled to_world must keep normals perpendicular to the transformed surface."""
    p = tmp_path / "slope.obj"
    # a 45° slope: plane x + z = 0 … normal (1,0,1)/√2, given in the file
    p.write_text("v 0 0 0\nv 1 0 -1\nv 0 1 0\nvn 0.70710678 0 0.70710678\nf 1//1 2//1 3//1\n")
    tw = S.T().translate([5.0, -3.0, 2.0]).scale([2.0, 1.0, 0.5])
    m = meshio.load_obj(str(p), to_world=tw)
    e1, e2 = m.positions[1] - m.positions[0], m.positions[2] - m.positions[0]
    for n in m.normals:
        assert abs(np.linalg.norm(n) - 1) < 1e-6
        assert abs(np.dot(n, e1)) < 1e-6 and abs(np.dot(n, e2)) < 1e-6         # still the surface normal
    expect = np.array([0.5, 0, 2.0]) / np.linalg.norm([0.5, 0, 2.0])            # M^-T (1,0,1) ∝ (1/2, 0, 2)
    assert np.allclose(m.normals[0], expect, atol=1e-6)
    # same through PLY
    src = meshio.load_obj(str(p))
    q = tmp_path / "slope.ply"; meshio.save_ply(str(q), src)
    m2 = meshio.load_ply(str(q), to_world=tw)
    assert np.allclose(m2.normals, m.normals, atol=1e-6) and np.allclose(m2.positions, m.positions, atol=1e-6)
