"""The per-iteration scene update: params.update() → vertex normals + BVH refresh.

The reference rebuilds its acceleration structure on every update (scene.cpp:304-336); the product refits the BVH4 on the
device.  Hit records do not depend on the structure, so the bar is the same as for the build: bit-exact against the
brute-force oracle after every update.  The CPU half runs the refit's per-node code on the test-only host harness;
the GPU half (-m gpu) runs k_refit_tris / k_refit_nodes through the C ABI and holds them to the harness bit for bit."""
import numpy as np
import pytest

from epsm_mitsuba3_b200 import scenes as S
from tests import parity
from tests.hostsim.hostsim import SimScene

CASES = {
    "cbox": lambda: S.cornell_box(16, 4, 12),
    "glass": lambda: S.glass_sphere_caustic(16, 4, 16, subdiv=3),
    "room": lambda: S.indoor_room(6000, 16, 4, 16),
    "smpl": lambda: S.articulated_stage(16, 4, 16),
}


def moved(sd, mesh_id, it, amp=0.03, base=None):
    """Vertex positions of iteration `it` of a fake optimisation: smooth drift + per-vertex noise, O(amp · scene extent)."""
    rng = np.random.default_rng(100 + it)
    p = (sd.meshes[mesh_id].positions if base is None else base).astype(np.float64)
    ext = float(np.ptp(np.concatenate([m.positions for m in sd.meshes]), axis=0).max())
    drift = amp * ext * np.array([np.sin(0.7 * it), np.cos(0.5 * it), 0.3 * np.sin(0.3 * it)])
    wobble = amp * ext * 0.5 * np.sin(3.0 * p / ext + it)
    return (p + drift + wobble + rng.normal(0, 0.05 * amp * ext, p.shape)).astype(np.float32)


def biggest_mesh(sd):
    return int(np.argmax([m.positions.shape[0] for m in sd.meshes]))


@pytest.fixture(scope="module", params=sorted(CASES))
def case(request, orc):
    sd = CASES[request.param]()
    sd.original = [m.positions.copy() for m in sd.meshes]          # OracleScene.update_vertices rebinds sd.meshes[i].positions
    return sd, SimScene(sd), orc.OracleScene(sd)


def test_refit_of_unchanged_positions_reproduces_the_build(case):
    sd, sim, _ = case
    n0, t0 = sim.accel()
    cost_built = sim.accel_cost()
    mid = biggest_mesh(sd)
    cost = sim.refit_vertices(mid, sd.original[mid])
    n1, t1 = sim.accel()
    assert np.array_equal(n0.view(np.uint32), n1.view(np.uint32))
    assert np.array_equal(t0.view(np.uint32), t1.view(np.uint32))
    assert cost == pytest.approx(cost_built, rel=1e-6)


def test_parent_links_are_consistent(case):
    _, sim, _ = case
    nodes, _ = sim.accel()
    link = nodes[:, 28].view(np.uint32); n_inner = nodes[:, 29].view(np.uint32); refs = nodes[:, 24:28].view(np.int32)
    assert link[0] == 0xffffffff and (link[1:] != 0xffffffff).all()
    for i in range(1, nodes.shape[0]):
        assert refs[link[i] >> 2, link[i] & 3] == i
    assert np.array_equal(n_inner, ((refs >= 0) & (refs != 0x7fffffff)).sum(1).astype(np.uint32))


def test_traversal_after_refit_is_bit_exact(case):
    sd, sim, o = case
    mid = biggest_mesh(sd)
    built = sim.accel_cost()
    for it in range(1, 4):
        p = moved(sd, mid, it, base=sd.original[mid])
        cost = sim.refit_vertices(mid, p)
        o.update_vertices(mid, p)
        frac = parity.check_trace(sim, o, parity.random_rays(sd, 4000, seed=it))
        assert frac > 0.02
        parity.check_trace(sim, o, parity.camera_rays(o, 1, it, 4, 512))
        assert cost > 0 and np.isfinite(cost)
    # restore: the refit of the original positions is again the original build
    sim.refit_vertices(mid, sd.original[mid]); o.update_vertices(mid, sd.original[mid])
    assert sim.accel_cost() == pytest.approx(built, rel=1e-6)


def test_refit_equals_rebuild_downstream(orc):
    """Gradients after a refit equal those after a rebuild: the path depends on hit records only."""
    sd = S.mirror_plane_scene(16, 4, 12)
    a, b = SimScene(sd), SimScene(sd)
    mid = biggest_mesh(sd)
    p = moved(sd, mid, 2, amp=0.01)
    a.refit_vertices(mid, p); b.update_vertices(mid, p)
    g = parity.synthetic_grad_in(16)
    ga, gb = a.backward("manifold", 2, 5, 4, 6, 5, g), b.backward("manifold", 2, 5, 4, 6, 5, g)
    assert np.array_equal(ga[0], gb[0]) and np.array_equal(ga[1], gb[1])
    assert np.array_equal(a.render_raw(1, 3, 2), b.render_raw(1, 3, 2))


# ------------------------------------------------------------------------------------------------
#  GPU: the kernels through the C ABI
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module", params=sorted(CASES))
def gpu_case(request, orc):
    import epsm_mitsuba3_b200 as ep
    sd = CASES[request.param]()
    sd.original = [m.positions.copy() for m in sd.meshes]
    return sd, ep.Scene(sd, device=0), SimScene(sd), orc.OracleScene(sd)


@pytest.mark.gpu
def test_gpu_refit_matches_harness_bit_for_bit(gpu_case):
    import torch
    sd, scene, sim, o = gpu_case
    mid = biggest_mesh(sd)
    nodes0 = scene.accel_raw()[0].cpu().numpy().copy()
    assert np.array_equal(nodes0.view(np.uint32), sim.accel()[0].view(np.uint32))      # same host builder on both sides
    for it in range(1, 4):
        p = moved(sd, mid, it, base=sd.original[mid])
        scene.update_vertices(mid, torch.from_numpy(p).cuda() if it % 2 else p, commit=False)
        scene.refit()
        cost_sim = sim.refit_vertices(mid, p)
        o.update_vertices(mid, p)
        gn, gt = (x.cpu().numpy() for x in scene.accel_raw())
        sn, st = sim.accel()
        assert np.array_equal(gt.view(np.uint32), st.view(np.uint32))
        assert np.array_equal(gn.view(np.uint32), sn.view(np.uint32))
        built, now, refits = scene.accel_cost()
        assert refits == it and now == pytest.approx(cost_sim, rel=1e-5)
        parity.check_trace(scene, o, parity.random_rays(sd, 20000, seed=it))
    scene.commit()                                                                       # rebuild of the moved scene
    assert scene.accel_cost()[2] == 0
    parity.check_trace(scene, o, parity.random_rays(sd, 20000, seed=9))
    # identity refit after the rebuild reproduces it
    before = scene.accel_raw()[0].cpu().numpy().copy()
    scene.update_vertices(mid, p, commit=False); scene.refit()
    assert np.array_equal(before.view(np.uint32), scene.accel_raw()[0].cpu().numpy().view(np.uint32))


@pytest.mark.gpu
def test_gpu_gradients_after_refit_equal_rebuild():
    import torch
    import epsm_mitsuba3_b200 as ep
    sd = S.glass_sphere_caustic(32, 4, 32, subdiv=3)
    a, b = ep.Scene(sd, device=0), ep.Scene(sd, device=0)
    mid = biggest_mesh(sd)
    p = moved(sd, mid, 1, amp=0.01)
    a.update_vertices(mid, p, commit=False); a.refit()
    b.update_vertices(mid, p, commit=False); b.commit()
    g = torch.from_numpy(parity.synthetic_grad_in(32)).cuda()
    outs = []
    for s in (a, b):
        gp = torch.zeros(s.n_vertices, 3, device="cuda"); gn = torch.zeros_like(gp)
        torch.cuda.synchronize()                              # the library runs on its own stream: order torch's fills before it
        s.render_backward_raw("manifold_caustic", 2, 5, 4, 6, 5, g, gp, gn)
        s.synchronize()
        outs.append((gp.cpu().numpy(), gn.cpu().numpy()))
    # same paths, same per-path values; only the atomic accumulation order differs between the two runs
    for x, y in zip(outs[0], outs[1]):
        assert np.allclose(x, y, rtol=1e-4, atol=1e-6 * max(1.0, float(np.abs(y).max())))
    assert np.abs(outs[1][0]).sum() > 0


@pytest.mark.gpu
def test_gpu_auto_policy_rebuilds_when_the_tree_degrades():
    import epsm_mitsuba3_b200 as ep
    sd = S.indoor_room(6000, 16, 4, 16)
    scene = ep.Scene(sd, device=0)
    params = ep.traverse(scene)
    scene.accel_update = "auto"; scene.rebuild_cost_ratio = 1.5
    key = [k for k, i in params._mesh_of.items() if i == biggest_mesh(sd)][0]
    params[key] = params[key] * 1.001
    params.update()
    assert scene.accel_cost()[2] == 1                        # small move: refitted
    rng = np.random.default_rng(0)
    import torch
    params[key] = params[key] + torch.from_numpy(rng.normal(0, 2.0, tuple(params[key].shape)).astype(np.float32)).cuda()
    params.update()
    assert scene.accel_cost()[2] == 0                        # scrambled: cost ratio exceeded → rebuilt
