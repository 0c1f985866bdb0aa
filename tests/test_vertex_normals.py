"""Vertex normals and their adjoint (grad_nrm must reach vertex_positions).

Reference: params.update() → Mesh::parameters_changed re-derives the normals of a mesh from its AD-attached positions
(mesh.cpp:82-87 → recompute_vertex_normals :257-330), so the `si_follow.sh_frame.n · g3` seed (epsm.py:644-645) flows
on into vertex_positions.  Product: k_normals_* (forward) and epsm_vertex_normals_backward (adjoint)."""
import numpy as np
import pytest
import torch

from epsm_mitsuba3_b200 import scenes as S
from oracle.normals_ref import normals_backward, recompute_vertex_normals
from tests.golden import make_reference_vectors as gen

V = np.load(gen.FIXTURE)


@pytest.mark.parametrize("key", ("body", "mirror", "blob"))
def test_oracle_forward_is_the_reference_recompute_vertex_normals(key):
    m = gen.normal_cases()[key]
    ref = V[f"normals/{key}/vertex_normals"]
    got = recompute_vertex_normals(torch.tensor(m.positions, dtype=torch.float64), torch.tensor(m.faces.astype(np.int64))).numpy()
    assert np.abs(got - ref).max() <= 2e-6
    # and the host-side generator used for the synthetic scenes
    assert np.abs(S.vertex_normals(m.positions, m.faces) - ref).max() <= 2e-6


def test_oracle_adjoint_matches_finite_differences():
    m = gen.normal_cases()["blob"]
    rng = np.random.default_rng(0)
    g = rng.normal(0, 1, m.positions.shape)
    jt = normals_backward(m.positions, m.faces, g)
    f = torch.tensor(m.faces.astype(np.int64))
    p0 = m.positions.astype(np.float64)
    for _ in range(8):
        d = rng.normal(0, 1, p0.shape); h = 1e-6
        a = recompute_vertex_normals(torch.tensor(p0 + h * d), f).numpy(); b = recompute_vertex_normals(torch.tensor(p0 - h * d), f).numpy()
        fd = ((a - b) / (2 * h) * g).sum()
        assert abs(fd - (jt * d).sum()) <= 1e-5 * max(1.0, abs(fd))


@pytest.mark.gpu
@pytest.mark.parametrize("key", ("body", "mirror", "blob"))
def test_gpu_normals_forward_and_adjoint(key):
    import epsm_mitsuba3_b200 as ep
    m0 = gen.normal_cases()[key]
    sd = S.SceneData()
    b = sd.add_bsdf(S.Bsdf.diffuse((0.5, 0.5, 0.5)))
    sd.add_mesh(S.grid(S.T().translate([0, 0, -5]).scale(20), 2, 2, name="floor", bsdf=b))                   # a mesh without normals in front
    sd.add_mesh(S.Mesh("m", m0.positions, m0.faces, normals=S.vertex_normals(m0.positions, m0.faces), bsdf=b, grad=True))
    sd.add_mesh(S.sphere_mesh((3, 0, 0), 0.5, 1, name="frozen", bsdf=b))                                      # normals, but no gradient tracking
    sd.sensors = [S.Sensor(S.T.look_at([0, -5, 2], [0, 0, 0], [0, 0, 1]), 40, 8, 8)] * 3
    scene = ep.Scene(sd, device=0)
    rng = np.random.default_rng(1)
    p = (m0.positions + rng.normal(0, 0.003, m0.positions.shape)).astype(np.float32)
    scene.update_vertices(1, p)                                           # device recompute (k_normals_accumulate / finish)
    _, nrm = scene.vertex_buffers(); scene.synchronize()
    v0, v1 = int(scene.voff[1]), int(scene.voff[2])
    ref_n = recompute_vertex_normals(torch.tensor(p, dtype=torch.float64), torch.tensor(m0.faces.astype(np.int64))).numpy()
    dn = np.abs(nrm[v0:v1].cpu().numpy() - ref_n).max()
    assert dn <= 5e-5, dn          # float32 atomics (82 faces meet at the poles of `body`) against the float64 restatement
    # adjoint: grad_pos += Jᵀ grad_nrm on the differentiable mesh only
    gn = rng.normal(0, 1, (sd.n_vertices, 3)).astype(np.float32)
    base = rng.normal(0, 1, (sd.n_vertices, 3)).astype(np.float32)
    for dev in (False, True):
        gp = base.copy()
        if dev:
            t = torch.from_numpy(gp).cuda(); scene.vertex_normals_backward(torch.from_numpy(gn).cuda(), t); scene.synchronize(); gp = t.cpu().numpy()
        else:
            scene.vertex_normals_backward(gn, gp)
        ref = normals_backward(p, m0.faces, gn[v0:v1])
        assert np.array_equal(gp[:v0], base[:v0]) and np.array_equal(gp[v1:], base[v1:])
        err = np.abs((gp[v0:v1] - base[v0:v1]) - ref)
        assert err.max() <= 1e-4 * np.abs(ref).max(), (err.max(), np.abs(ref).max())


@pytest.mark.gpu
def test_gpu_params_receive_the_normal_term_and_expose_vertex_normals():
    """Smooth-shaded mirror: `.grad` of vertex_positions = grad_pos + Jᵀ grad_nrm, and `<mesh>.vertex_normals` is a key."""
    import epsm_mitsuba3_b200 as ep
    from tests import parity
    sd = S.mirror_plane_scene(32, 4, 16)
    scene = ep.Scene(sd, 0)
    params = ep.traverse(scene)
    assert "mirror.vertex_normals" in params and "wall.vertex_normals" not in params and "wall.vertex_positions" in params
    key = "mirror.vertex_positions"
    params[key] = params[key].clone().requires_grad_(True)
    nk = "mirror.vertex_normals"
    params.update()
    params[nk].requires_grad_(True)
    integ = ep.load_dict({"type": "manifold", "max_depth": 6})
    g = torch.from_numpy(parity.synthetic_grad_in(16)).cuda()
    integ.render_backward(scene, params, g, seed=5)
    gp, gn = scene.render_backward_raw("manifold", 2, 5, 8, 6, 5, g); scene.synchronize()
    v0, v1 = int(scene.voff[0]), int(scene.voff[1])
    assert gn[v0:v1].abs().sum() > 0, "the scene must exercise the shading-normal seed"
    assert torch.allclose(params[nk].grad, gn[v0:v1], rtol=1e-4, atol=1e-7)
    pos, _ = scene.vertex_buffers()
    jt = normals_backward(pos[v0:v1].cpu().numpy(), sd.meshes[0].faces, gn[v0:v1].cpu().numpy())
    want = gp[v0:v1].cpu().numpy() + jt
    got = params[key].grad.cpu().numpy()
    assert np.abs(got - want).max() <= 2e-4 * np.abs(want).max()
    assert np.abs(jt).max() > 1e-3 * np.abs(want).max(), "the normal term is not negligible here"
    # writing normals directly: keys = [vertex_normals] → no recomputation (mesh.cpp:82-83)
    n_new = torch.nn.functional.normalize(params[nk].detach() + 0.05, dim=-1)
    params[nk] = n_new; params.update()
    _, nrm = scene.vertex_buffers(); scene.synchronize()
    assert torch.allclose(nrm[v0:v1], n_new, atol=1e-7)
    with pytest.raises(ep.EpsmError):
        scene.update_normals(1, np.zeros((sd.meshes[1].positions.shape[0], 3), np.float32))     # mesh.cpp:258-260
