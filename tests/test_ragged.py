"""Edge cases of the wavefront layout: films whose size is not a multiple of the 8×4 work tile (ragged tiles), non-square films,
one-pixel films, one-sample wavefronts, lane windows that start and end inside a tile.  Host harness on the CPU; the same
checks through the C ABI on the B200 (`-m gpu`)."""
import numpy as np
import pytest

from epsm_mitsuba3_b200 import scenes as S
from tests import parity
from tests.hostsim.hostsim import SimScene


def ragged_scene(w, h, match=None):
    sd = S.mirror_plane_scene(16, 2, 12)
    for k, s in enumerate(sd.sensors):
        s.width, s.height = (w, h) if k < 2 else ((match or w), (match or w))      # sensor 2 is square (epsm.py:238-240 crops [:res,:res])
    return sd


SIZES = [(30, 17), (9, 5), (1, 1), (7, 33)]


@pytest.mark.parametrize("size", SIZES)
def test_ragged_primal_film_hostsim(orc, size):
    sd = ragged_scene(*size, match=12)
    sim, o = SimScene(sd), orc.OracleScene(sd, accel=True)
    for spp in (1, 3):
        parity.check_film(sim.render_raw(1, 5, spp, 6, 5), o.render_raw(1, 5, spp, 6, 5))
        parity.check_film(sim.render_raw(1, 5, spp, 6, 5, staged=True), o.render_raw(1, 5, spp, 6, 5))


@pytest.mark.parametrize("match", [10, 13, 1])
def test_ragged_gradient_sensor_hostsim(orc, match):
    sd = ragged_scene(16, 16, match=match)
    sim, o = SimScene(sd), orc.OracleScene(sd, accel=True)
    g = parity.synthetic_grad_in(match)
    n = match * match * 8
    paths = o.path_log(2, 11, 8, 6, 5, g)
    rec = sim.path_records("manifold", 2, 11, 8, 6, 5, g, 0, n)
    stats, (pg, lg, dg) = parity.check_path_records(rec, paths, "manifold", 6, min_same=0.99 if match > 1 else 0.0)
    ref_p, ref_n = parity.scatter_gradients(paths, "manifold", parity.n_logged(6), pg, lg, dg, sd.n_vertices, 6)
    gp, gn = sim.backward("manifold", 2, 11, 8, 6, 5, g)
    parity.check_accumulated(gp, gn, ref_p, ref_n)
    # a lane window that starts and ends inside a tile
    if n > 40:
        a = sim.backward("manifold", 2, 11, 8, 6, 5, g, 0, 37); b = sim.backward("manifold", 2, 11, 8, 6, 5, g, 37, n)
        assert np.allclose(a[0] + b[0], gp, rtol=1e-4, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("size", SIZES)
def test_ragged_primal_film_gpu(orc, size):
    import epsm_mitsuba3_b200 as ep
    sd = ragged_scene(*size, match=12)
    scene, o = ep.Scene(sd, device=0), orc.OracleScene(sd, accel=True)
    for spp in (1, 3):
        scene.film_clear(1); scene.render_accumulate(1, 5, spp, 6, 5); scene.synchronize()
        parity.check_film(scene.film_raw(1).cpu().numpy(), o.render_raw(1, 5, spp, 6, 5))
    img = ep.load_dict({"type": "manifold"}).render(scene, sensor=1, seed=5, spp=3)
    assert tuple(img.shape) == (size[1], size[0], 5)


@pytest.mark.gpu
@pytest.mark.parametrize("match", [10, 13, 1])
def test_ragged_gradient_sensor_gpu(orc, match):
    import epsm_mitsuba3_b200 as ep
    sd = ragged_scene(16, 16, match=match)
    scene, o = ep.Scene(sd, device=0), orc.OracleScene(sd, accel=True)
    g = parity.synthetic_grad_in(match)
    n = match * match * 8
    paths = o.path_log(2, 11, 8, 6, 5, g)
    rec = scene.path_records("manifold", 2, 11, 8, 6, 5, g, 0, n)
    stats, (pg, lg, dg) = parity.check_path_records(rec, paths, "manifold", 6, min_same=0.99 if match > 1 else 0.0)
    ref_p, ref_n = parity.scatter_gradients(paths, "manifold", parity.n_logged(6), pg, lg, dg, sd.n_vertices, 6)
    gp, gn = scene.render_backward_raw("manifold", 2, 11, 8, 6, 5, g)
    parity.check_accumulated(gp, gn, ref_p, ref_n)
    if n > 40:
        a = scene.render_backward_raw("manifold", 2, 11, 8, 6, 5, g, lane_begin=0, lane_end=37)
        b = scene.render_backward_raw("manifold", 2, 11, 8, 6, 5, g, lane_begin=37, lane_end=n)
        assert np.allclose(a[0] + b[0], gp, rtol=1e-4, atol=1e-6)


@pytest.mark.gpu
def test_gradient_pass_odd_sample_counts_and_sample_shards_gpu(orc):
    """The gradient pass enumerates a tile pixel-major (a warp = consecutive samples of one pixel): sample counts that are not a
    multiple of the warp size, on a ragged sensor, against the host harness; and sample-index shards (the multi-GPU partition,
    epsm_set_sample_range) of unequal, odd sizes must add up to the whole."""
    import epsm_mitsuba3_b200 as ep
    sd = ragged_scene(16, 16, match=13)
    scene, sim = ep.Scene(sd, device=0), SimScene(sd)
    g = parity.synthetic_grad_in(13)
    for spp in (5, 100):
        gp, gn = scene.render_backward_raw("manifold", 2, 7, spp, 6, 5, g)
        hp, hn = sim.backward("manifold", 2, 7, spp, 6, 5, g)
        gp, gn = np.asarray(gp), np.asarray(gn)                 # numpy grad_in → numpy buffers back
        scale = max(1e-6, float(np.abs(hp).max()))
        assert np.abs(gp - hp).max() < 2e-4 * scale and np.abs(gn - hn).max() < 2e-4 * max(1e-6, float(np.abs(hn).max())), spp
    parts = []
    for s0, s1 in ((0, 37), (37, 100)):
        scene.set_sample_range(s0, s1)
        try:
            a, _ = scene.render_backward_raw("manifold", 2, 7, 100, 6, 5, g)
        finally:
            scene.set_sample_range(0, 0)
        parts.append(np.asarray(a))
    assert np.allclose(parts[0] + parts[1], gp, rtol=2e-4, atol=2e-4 * scale)


@pytest.mark.gpu
def test_gradient_pass_on_the_border_sensor_gpu(orc):
    """The C ABI takes any sensor: the gradient pass on sensor 0 (hdrfilm sample_border: (W+4)·(H+4)·spp lanes, the border's samples
    carry no pixel-motion gradient) against the host harness, at an odd sample count."""
    import epsm_mitsuba3_b200 as ep
    sd = ragged_scene(14, 11, match=12)
    assert sd.sensors[0].sample_border
    scene, sim = ep.Scene(sd, device=0), SimScene(sd)
    g = parity.synthetic_grad_in(14)[:11]                       # (H, W, 5) of sensor 0's film
    for spp in (3, 40):
        gp, gn = scene.render_backward_raw("manifold", 0, 9, spp, 6, 5, np.ascontiguousarray(g))
        hp, hn = sim.backward("manifold", 0, 9, spp, 6, 5, g)
        scale = max(1e-6, float(np.abs(hp).max()))
        assert np.abs(hp).max() > 0 and np.abs(np.asarray(gp) - hp).max() < 2e-4 * scale, spp
