"""Parity of the product's per-lane device code (compiled for the host by the test-only harness tests/hostsim)
against the CPU oracle.  Runs without a GPU; the same checks run through the C ABI on the B200 in test_gpu_parity.py."""
import numpy as np
import pytest

from epsm_mitsuba3_b200 import scenes as S
from tests import parity
from tests.hostsim.hostsim import SimScene

CASES = {
    "mirror": (lambda: S.mirror_plane_scene(16, 4, 12), "manifold", 6),
    "rough": (lambda: S.mirror_plane_scene(16, 4, 12, rough=True), "manifold", 6),
    "glass": (lambda: S.glass_sphere_caustic(16, 4, 16, subdiv=2), "manifold_caustic", 6),
    "cbox_d3": (lambda: S.cornell_box(16, 4, 12), "manifold", 3),
    "cbox_mirror": (lambda: S.cornell_box(16, 4, 12, box_bsdf=S.Bsdf.conductor()), "manifold", 6),
    "cbox_glass": (lambda: S.cornell_box(16, 4, 12, box_bsdf=S.Bsdf.dielectric(1.5)), "manifold_caustic", 5),
    "room": (lambda: S.indoor_room(6000, 16, 4, 16), "manifold", 6),
    "stage": (lambda: S.articulated_stage(16, 4, 16), "manifold_caustic", 3),      # BASELINE configs[3] at test size
    # uv-mapped meshes + bitmap reflectances (EPSM/exp/human.py:101-163): si.uv, uv-aligned shading frame, texture lookup
    "textured": (lambda: S.textured_scene(16, 4, 12, "nearest", "repeat"), "manifold", 6),
    "textured_bilinear": (lambda: S.textured_scene(16, 4, 12, "bilinear", "mirror"), "manifold", 4),
}


@pytest.fixture(scope="module", params=sorted(CASES))
def case(request, orc):
    make, variant, max_depth = CASES[request.param]
    sd = make()
    return sd, SimScene(sd), orc.OracleScene(sd), variant, max_depth


def test_rng_stream_bit_exact(orc):
    sim = SimScene(S.cornell_box(8, 1, 8))
    for seed in (0, 1, 7):
        assert np.array_equal(sim.sampler_stream(seed, 0, 512, 16).view(np.uint32), orc.sampler_stream(seed, 0, 512, 16).view(np.uint32))
    assert np.array_equal(sim.sampler_stream(3, 4000000000, 4000000064, 4).view(np.uint32),
                          orc.sampler_stream(3, 4000000000, 4000000064, 4).view(np.uint32))


def test_traversal_bit_exact(case):
    sd, sim, o, _, _ = case
    frac = parity.check_trace(sim, o, parity.random_rays(sd, 6000))
    assert 0.02 < frac <= 1.0
    parity.check_trace(sim, o, parity.camera_rays(o, 1, 5, 4, 1024))


def test_ray_generation(case):
    sd, sim, o, _, _ = case
    a, b = sim.sample_rays(1, 9, 4, 0, 1024), o.sample_rays(1, 9, 4, 0, 1024)
    assert np.allclose(a, b, rtol=1e-6, atol=1e-7)


def test_primal_film(case):
    sd, sim, o, _, max_depth = case
    o.set_accel(True)
    parity.check_film(sim.render_raw(1, 3, 4, max_depth, 5), o.render_raw(1, 3, 4, max_depth, 5))
    o.set_accel(False)


def test_staged_primal_equals_path_loop(case):
    """The wavefront stage of the primal render (pr_stage: NEE term parked until its shadow ray answers) must give the
    radiance of the plain path loop bit for bit — same operands, same order (epsm.py:466)."""
    sd, sim, o, _, max_depth = case
    for md in sorted({max_depth, 2, 12}):
        a, b = sim.render_raw(1, 3, 4, md, 5), sim.render_raw(1, 3, 4, md, 5, staged=True)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"max_depth={md}"
    assert np.abs(a[..., :3]).max() > 0


def test_per_path_gradients_and_scatter(case):
    sd, sim, o, variant, max_depth = case
    o.set_accel(True)
    res = sd.sensors[2].width
    g = parity.synthetic_grad_in(res)
    n = res * res * 8
    paths = o.path_log(2, 11, 8, max_depth, 5, g)
    rec = sim.path_records(variant, 2, 11, 8, max_depth, 5, g, 0, n)
    stats, (pg, lg, dg) = parity.check_path_records(rec, paths, variant, max_depth)
    assert stats["nonzero"] > 0
    ref_p, ref_n = parity.scatter_gradients(paths, variant, parity.n_logged(max_depth), pg, lg, dg, sd.n_vertices, max_depth)
    gp, gn = sim.backward(variant, 2, 11, 8, max_depth, 5, g)
    parity.check_accumulated(gp, gn, ref_p, ref_n)
    o.set_accel(False)


def test_lane_ranges_partition_the_wavefront(case):
    sd, sim, o, variant, max_depth = case
    res = sd.sensors[2].width; g = parity.synthetic_grad_in(res); n = res * res * 8
    full = sim.backward(variant, 2, 5, 8, max_depth, 5, g)
    a = sim.backward(variant, 2, 5, 8, max_depth, 5, g, 0, n // 3)
    b = sim.backward(variant, 2, 5, 8, max_depth, 5, g, n // 3, n)
    assert np.allclose(a[0] + b[0], full[0], rtol=1e-5, atol=1e-7) and np.allclose(a[1] + b[1], full[1], rtol=1e-5, atol=1e-7)


def test_pruned_replay_equals_full_replay(case):
    """Gradient-liveness pruning (epsm_set_full_replay) must not change the accumulated gradients: the host harness adds
    the per-path seeds in lane order, so the two passes are required to agree bit for bit."""
    sd, sim, o, variant, max_depth = case
    res = sd.sensors[2].width; g = parity.synthetic_grad_in(res)
    for md in sorted({max_depth, 3, 8}):
        full = sim.backward(variant, 2, 13, 8, md, 5, g, prune=False)
        pruned = sim.backward(variant, 2, 13, 8, md, 5, g, prune=True)
        assert np.array_equal(full[0], pruned[0]) and np.array_equal(full[1], pruned[1]), f"max_depth={md}"
        # the production schedule: first-vertex fast path of k_wf_stage_a + the chain-prefix log of k_wf_finish
        fast = sim.backward(variant, 2, 13, 8, md, 5, g, prune=2)
        assert np.array_equal(full[0], fast[0]) and np.array_equal(full[1], fast[1]), f"fast path, max_depth={md}"
    assert np.abs(full[0]).max() > 0 or variant == "manifold_caustic"


def test_update_vertices_rebuilds(orc):
    sd = S.mirror_plane_scene(8, 2, 8)
    sim, o = SimScene(sd), orc.OracleScene(sd)
    p = sd.meshes[0].positions.copy(); p[:, 2] += 0.3
    sim.update_vertices(0, p); o.update_vertices(0, p)
    parity.check_trace(sim, o, parity.random_rays(sd, 3000))
