"""`prb_reparam` reverse pass (hybrid phase 2, EPSM/optim.py:89-92,113-119).

Oracle: oracle/prb_reparam_ref.py — the detached half logged by the scalar C++ oracle (`orc_prb_log`), the attached half
re-evaluated with torch autograd in float64, term by term as prb_reparam.py:329-602 / reparam.py:83-330 / common.py:372-397,
905-945 write them.  The product differentiates the same terms BY HAND in float32 (csrc/epsm_prb.cuh).  PARITY UNPINNED: the
reference's AD variants cannot run in this image; like the reference's own tests (src/python/python/tests/
test_ad_integrators.py) the last test holds the gradient to finite differences of the primal renderer.

Not-GPU tests run the lane code on the host harness; `-m gpu` tests run k_prb_backward through the C ABI."""
import numpy as np
import pytest

from epsm_mitsuba3_b200 import scenes as S

RES = 16


def occluder_scene(smooth: bool, rough_floor: bool, res: int = RES, offset=(0.0, 0.0, 0.0), textured: bool = False) -> S.SceneData:
    """Floor + small area light + an icosphere between them (the differentiable mesh), seen from above at an angle."""
    sd = S.SceneData(); T = S.Transform4f
    dif = sd.add_bsdf(S.Bsdf.diffuse((0.7, 0.6, 0.5))); dif2 = sd.add_bsdf(S.Bsdf.diffuse((0.3, 0.5, 0.8)))
    rc = sd.add_bsdf(S.Bsdf.roughconductor(0.3))
    floor = S.rectangle(T().scale(4.0), name="floor"); floor.bsdf = rc if rough_floor else dif; sd.add_mesh(floor)
    light = S.rectangle(T().translate([0.3, 0.2, 4.0]).rotate([1, 0, 0], 180).scale(0.3), name="light"); light.bsdf = dif
    sd.add_mesh(light, radiance=(40.0, 40.0, 40.0))
    v, f = S.icosphere(1)
    p = v * 0.5 + np.array([0.1, 0.0, 1.2]) + np.asarray(offset)
    occ = S.Mesh("occ", p, f, normals=(v.astype(np.float32) if smooth else None), grad=True); occ.bsdf = dif2
    if textured:
        # uv-mapped occluder with a nearest-filtered bitmap reflectance (EPSM/exp/human.py:101-163) and a uv-mapped floor: the
        # shading frames follow dp/du (mesh.cpp:746-763), the texture is piecewise constant
        rng = np.random.default_rng(4)
        occ.texcoords = np.stack([np.arctan2(v[:, 1], v[:, 0]) / (2 * np.pi) + 0.5 + 0.1 * v[:, 2], 0.5 * v[:, 2] + 0.5 + 0.05 * v[:, 0]], -1).astype(np.float32)
        sd.bsdfs[dif2].texture = S.Texture(rng.uniform(0.2, 0.9, (4, 5, 3)).astype(np.float32), filter="nearest", wrap="repeat")
        floor.texcoords = np.array([[0, 0], [2, 0.3], [2.2, 1.5], [-0.1, 1.2]], np.float32)
    sd.add_mesh(occ)
    sd.sensors = [S.Sensor(T.look_at([0.5, -4.0, 3.0], [0, 0, 0.8], [0, 0, 1]), 45.0, res, res, 0.1, 100.0, 4, True)]
    return sd


def image_weights(res):
    yy, xx = np.mgrid[0:res, 0:res] / res
    return np.stack([np.sin(3 * xx + 1) + 0.5 * yy, np.cos(2 * yy) + xx, 0.3 + xx * yy], -1).astype(np.float32)


CONFIGS = [  # smooth normals, rough-conductor floor, max_depth, spp, kappa
    (False, False, 2, 4, 1e3),
    (True, False, 3, 4, 1e3),
    (True, True, 4, 4, 1e5),
    (False, True, 6, 8, 1e3),
]


def _cfg(md, kappa, rays=16, rmd=None):
    return dict(max_depth=md, rr_depth=5, reparam_max_depth=md if rmd is None else rmd, reparam_rays=rays, reparam_kappa=kappa, reparam_exp=3.0)


def _rel(a, b):
    return float(np.abs(a - b).max() / max(1e-12, np.abs(b).max()))


def _oracle(sd, gin, seed, spp, md, kappa, rays=16, **kw):
    from oracle import oracle as O, prb_reparam_ref as R
    o = O.OracleScene(sd)
    gp, gn = R.render_backward(o, sd, gin, sensor=0, seed=seed, spp=spp, max_depth=md, kappa=kappa, n_rays=rays, **kw)
    voff = sd.vertex_offsets()
    return gp[voff[2]:voff[3]], gn[voff[2]:voff[3]], o


@pytest.mark.parametrize("smooth", (False, True))
def test_hand_adjoint_with_uv_aligned_frames_and_a_bitmap_reflectance(smooth):
    """Meshes with texture coordinates: sh_frame.s follows dp/du = (Δv₁·e1 − Δv₀·e2)/det, which is linear in the edges — its
    adjoint replaces that of coordinate_system(n_g); the nearest-filtered texture is piecewise constant in uv."""
    from tests.hostsim import hostsim as HS
    md, spp, kappa = 3, 4, 1e3
    sd = occluder_scene(smooth, True, textured=True)
    gin = image_weights(RES)
    gp_o, gn_o, o = _oracle(sd, gin, 5, spp, md, kappa)
    h = HS.SimScene(sd)
    film = h.render_raw(0, 5, spp, md, 5)
    assert np.array_equal(film, o.render_raw(0, 5, spp, md, 5))
    gp, gn = h.prb_backward(0, 5, spp, _cfg(md, kappa), film, gin)
    voff = sd.vertex_offsets()
    assert np.abs(gp_o).max() > 1e-3 and _rel(gp[voff[2]:voff[3]], gp_o) < 5e-4
    if smooth:
        assert _rel(gn[voff[2]:voff[3]], gn_o) < 5e-4
    # and the frames do differ from the untextured scene's (the case is not vacuous)
    plain = occluder_scene(smooth, True)
    gp_p, _ = HS.SimScene(plain).prb_backward(0, 5, spp, _cfg(md, kappa), HS.SimScene(plain).render_raw(0, 5, spp, md, 5), gin)
    assert _rel(gp_p[voff[2]:voff[3]], gp_o) > 1e-2


@pytest.mark.parametrize("smooth,rough,md,spp,kappa", CONFIGS)
def test_hand_adjoint_equals_autograd_oracle_on_the_host(smooth, rough, md, spp, kappa):
    from tests.hostsim import hostsim as HS
    sd = occluder_scene(smooth, rough)
    gin = image_weights(RES)
    gp_o, gn_o, o = _oracle(sd, gin, 5, spp, md, kappa)
    h = HS.SimScene(sd)
    film = h.render_raw(0, 5, spp, md, 5)
    assert np.array_equal(film, o.render_raw(0, 5, spp, md, 5))          # same primal film ⇒ same δL
    gp, gn = h.prb_backward(0, 5, spp, _cfg(md, kappa), film, gin)
    voff = sd.vertex_offsets()
    assert np.abs(gp_o).max() > 1e-3
    assert _rel(gp[voff[2]:voff[3]], gp_o) < 5e-4                        # per vertex, float32 by hand vs float64 autograd
    if smooth:
        assert np.abs(gn_o).max() > 1e-4 and _rel(gn[voff[2]:voff[3]], gn_o) < 5e-4
    assert not gp[:voff[2]].any() and not gn[:voff[2]].any()             # only attached meshes receive gradients
    # the replay wavefront (k_prb_pass's policy, emulated per lane) walks the same lanes pass by pass: identical sums
    gp_w, gn_w = h.prb_backward(0, 5, spp, _cfg(md, kappa), film, gin, wavefront=True)
    assert np.array_equal(gp_w, gp) and np.array_equal(gn_w, gn)


def test_lane_windows_and_reparam_depth_on_the_host():
    """Per-lane parity (single-lane windows against the oracle's window) and reparam_max_depth = 1 (prb_reparam.py:258-264:
    only the camera ray and the first vertex's ray are reparameterised)."""
    from tests.hostsim import hostsim as HS
    from oracle import oracle as O, prb_reparam_ref as R
    sd = occluder_scene(True, False)
    gin = image_weights(RES)
    o = O.OracleScene(sd); h = HS.SimScene(sd)
    film = h.render_raw(0, 2, 4, 3, 5)
    voff = sd.vertex_offsets()
    n_lanes = sd.sensors[0].wavefront(4)
    hit = 0
    for lane in range(0, n_lanes, 97):
        gp_o, _ = R.render_backward(o, sd, gin, sensor=0, seed=2, spp=4, max_depth=3, kappa=1e3, lane_begin=lane, lane_end=lane + 1, film_rgbw=film)
        gp, _ = h.prb_backward(0, 2, 4, _cfg(3, 1e3), film, gin, lane, lane + 1)
        a, b = gp[voff[2]:voff[3]], gp_o[voff[2]:voff[3]]
        if np.abs(b).max() > 1e-6:
            hit += 1
            assert _rel(a, b) < 2e-3, lane
        else:
            assert np.abs(a).max() < 1e-5
    assert hit >= 3
    gp_o, _ = R.render_backward(o, sd, gin, sensor=0, seed=2, spp=4, max_depth=3, reparam_max_depth=1, kappa=1e3)
    gp, _ = h.prb_backward(0, 2, 4, _cfg(3, 1e3, rmd=1), film, gin)
    assert _rel(gp[voff[2]:voff[3]], gp_o[voff[2]:voff[3]]) < 5e-4


def test_bsdf_duals_match_autograd_and_finite_differences():
    """The product's 6-wide forward duals of BSDF::eval (bsdf_grad, epsm_prb.cuh) against torch autograd of the oracle's
    restatement and against central differences of the product's own eval."""
    import torch
    from tests.hostsim import hostsim as HS
    from oracle import prb_reparam_ref as R
    sd = occluder_scene(False, False)
    bsdfs = [S.Bsdf.diffuse((0.2, 0.5, 0.9)), S.Bsdf.roughconductor(0.25), S.Bsdf.roughconductor(0.4, ggx=True),
             S.Bsdf.roughconductor(0.3, twosided=True), S.Bsdf.diffuse((0.5, 0.5, 0.5), twosided=True), S.Bsdf.dielectric(1.5), S.Bsdf.conductor()]
    first = len(sd.bsdfs)
    for b in bsdfs:
        sd.add_bsdf(b)
    h = HS.SimScene(sd)
    rng = np.random.default_rng(3)
    a = np.array([0.3, 0.5, 0.7])
    for i, b in enumerate(bsdfs):
        for _ in range(20):
            wi = rng.normal(size=3); wi[2] = abs(wi[2]) + 0.3; wi /= np.linalg.norm(wi)
            wo = rng.normal(size=3); wo[2] = abs(wo[2]) + 0.3; wo /= np.linalg.norm(wo)
            if b.twosided and rng.random() < 0.5:
                wi[2], wo[2] = -wi[2], -wo[2]
            val, g_wi, g_wo = h.bsdf_grad(first + i, wi, wo, a)
            x = torch.tensor(np.concatenate([wi, wo])[None].astype(np.float32).astype(np.float64), requires_grad=True)
            f = R._bsdf_eval_one(b, x[:, :3], x[:, 3:])
            np.testing.assert_allclose(val, f.detach().numpy()[0], rtol=2e-5, atol=1e-7)
            if f.requires_grad:
                (f * torch.tensor(a)).sum().backward()
                g = x.grad[0].numpy()
            else:
                g = np.zeros(6)                                              # delta BSDFs evaluate to zero (dielectric, conductor)
            np.testing.assert_allclose(np.concatenate([g_wi, g_wo]), g, rtol=2e-4, atol=2e-5 * max(1.0, np.abs(g).max()))
            eps = 2e-3
            for k in range(6):
                d = np.zeros(6); d[k] = eps
                fp = h.bsdf_eval_pdf(first + i, wi + d[:3], wo + d[3:])[0].astype(np.float64)
                fm = h.bsdf_eval_pdf(first + i, wi - d[:3], wo - d[3:])[0].astype(np.float64)
                fd = ((fp - fm) * a).sum() / (2 * eps)
                assert abs(fd - g[k]) < 3e-2 * max(1.0, np.abs(g).max()), (i, k, fd, g[k])


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("smooth,rough,md,spp,kappa", CONFIGS)
def test_gpu_kernel_equals_oracle(smooth, rough, md, spp, kappa):
    import torch
    import epsm_mitsuba3_b200 as ep
    sd = occluder_scene(smooth, rough)
    gin = image_weights(RES)
    gp_o, gn_o, _ = _oracle(sd, gin, 5, spp, md, kappa)
    scene = ep.Scene(sd)
    integ = ep.load_dict({"type": "prb_reparam", "max_depth": md, "reparam_kappa": kappa})
    gp, gn = integ.vertex_gradients(scene, torch.from_numpy(gin).cuda(), 0, 5, spp)
    voff = sd.vertex_offsets()
    gp, gn = gp.cpu().numpy().astype(np.float64), gn.cpu().numpy().astype(np.float64)
    assert _rel(gp[voff[2]:voff[3]], gp_o) < 1e-3                        # atomics reorder the float32 sums
    if smooth:
        assert _rel(gn[voff[2]:voff[3]], gn_o) < 1e-3
    assert not gp[:voff[2]].any()
    # host pointers through the C ABI give the same buffers; lane windows partition the pass
    g2, n2 = np.zeros((sd.n_vertices, 3), np.float32), np.zeros((sd.n_vertices, 3), np.float32)
    n_lanes = sd.sensors[0].wavefront(spp)
    for lb, le in ((0, n_lanes // 3), (n_lanes // 3, n_lanes)):
        scene.prb_reparam_backward_raw(0, 5, spp, integ, gin, g2, n2, lb, le)
    assert _rel(g2[voff[2]:voff[3]].astype(np.float64), gp[voff[2]:voff[3]]) < 1e-3


@pytest.mark.gpu
def test_gpu_gradient_matches_finite_differences_of_the_primal_render():
    """What the reference's own tests check for this integrator (test_ad_integrators.py): d/dθ Σ image·w for a rigid
    translation θ of the occluder, against central differences of converged primal renders.  The warp-field estimator is
    biased for few auxiliary rays (reparam.py:33-36), so the comparison uses 64 rays and a κ matched to this film's pixel size."""
    import torch
    import epsm_mitsuba3_b200 as ep
    res, md = 32, 3
    gin = image_weights(res)
    sd = occluder_scene(False, False, res)
    scene = ep.Scene(sd)
    integ = ep.load_dict({"type": "prb_reparam", "max_depth": md, "reparam_kappa": 2e3, "reparam_rays": 64})
    voff = sd.vertex_offsets()
    g = torch.from_numpy(gin).cuda()
    grads = []
    for seed in range(8):
        gp, _ = integ.vertex_gradients(scene, g, 0, seed, 256)
        grads.append(gp[int(voff[2]):int(voff[3])].sum(0).cpu().numpy())
    grads = np.array(grads); mean, sem = grads.mean(0), grads.std(0) / np.sqrt(len(grads))
    path = ep.load_dict({"type": "path", "max_depth": md})
    base = sd.meshes[2].positions.copy()
    eps = 0.01
    fd = np.zeros(3); fd_sem = np.zeros(3)
    for ax in range(3):
        ds = []
        for seed in range(6):
            vals = []
            for sgn in (1, -1):
                off = np.zeros(3, np.float32); off[ax] = sgn * eps
                scene.update_vertices(2, base + off)
                img = path.render(scene, sensor=0, seed=100 + seed, spp=8192)
                vals.append(float((img.double() * g.double()).sum()))
            ds.append((vals[0] - vals[1]) / (2 * eps))
        fd[ax], fd_sem[ax] = np.mean(ds), np.std(ds) / np.sqrt(len(ds))
    scene.update_vertices(2, base)
    print("prb_reparam gradient", mean, "+-", sem, " finite differences", fd, "+-", fd_sem)
    tol = 4 * np.sqrt(sem ** 2 + fd_sem ** 2) + 0.15 * np.abs(fd).max()
    assert np.all(np.abs(mean - fd) < tol), (mean, fd, tol)
    assert np.abs(fd).max() > 1.0                                            # the test scene has a real signal
