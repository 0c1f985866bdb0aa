"""GPU parity tests proper: the CUDA path, called through the C ABI of libepsm_b200.so, against the CPU oracle on the
same seeded inputs.  Run on the B200 box with `pytest -m gpu`."""
import ctypes as C

import numpy as np
import pytest
import torch

from epsm_mitsuba3_b200 import capi, scenes as S
from tests import parity

pytestmark = pytest.mark.gpu


class GpuBackend:
    """Adapts epsm_mitsuba3_b200.Scene (C ABI, device) to the backend interface of tests/parity.py."""

    def __init__(self, sd):
        import epsm_mitsuba3_b200 as ep
        self.scene = ep.Scene(sd, device=0); self.data = sd

    def trace(self, rays, any_hit=False, counters=None):
        return self.scene.trace(np.ascontiguousarray(rays, np.float32), any_hit, counters)

    def sampler_stream(self, *a):
        return self.scene.sampler_stream(*a)

    def sample_rays(self, *a):
        return self.scene.sample_rays(*a)

    def render_raw(self, sensor, seed, spp, max_depth=6, rr_depth=5, lb=0, le=0):
        self.scene.film_clear(sensor)
        self.scene.render_accumulate(sensor, seed, spp, max_depth, rr_depth, lb, le)
        self.scene.synchronize()
        return self.scene.film_raw(sensor).cpu().numpy()

    def path_records(self, *a):
        return self.scene.path_records(*a)

    def backward(self, variant, sensor, seed, spp, max_depth, rr_depth, g, lb=0, le=0):
        gp, gn = self.scene.render_backward_raw(variant, sensor, seed, spp, max_depth, rr_depth, np.ascontiguousarray(g, np.float32),
                                                lane_begin=lb, lane_end=le)
        return gp, gn

    def update_vertices(self, mesh_id, p):
        self.scene.update_vertices(mesh_id, p)


from tests.cases import CASES  # noqa: E402


@pytest.fixture(scope="module", params=sorted(CASES))
def case(request, orc):
    make, variant, max_depth = CASES[request.param]
    sd = make()
    return sd, GpuBackend(sd), orc.OracleScene(sd), variant, max_depth


def test_library_is_loaded_and_counts_launches():
    b = GpuBackend(S.cornell_box(8, 1, 8))
    n0 = b.scene.launch_count()
    b.sampler_stream(0, 0, 8, 2)
    assert b.scene.launch_count() == n0 + 1
    with open("/proc/self/maps") as f:
        assert "libepsm_b200.so" in f.read()


def test_rng_stream_bit_exact(orc):
    b = GpuBackend(S.cornell_box(8, 1, 8))
    for seed in (0, 1, 7):
        assert np.array_equal(b.sampler_stream(seed, 0, 4096, 64).view(np.uint32), orc.sampler_stream(seed, 0, 4096, 64).view(np.uint32))
    assert np.array_equal(b.sampler_stream(3, 4000000000, 4000000064, 4).view(np.uint32),
                          orc.sampler_stream(3, 4000000000, 4000000064, 4).view(np.uint32))


def test_traversal_bit_exact(case):
    sd, b, o, _, _ = case
    n = 20000 if sd.n_triangles < 5000 else 6000          # the oracle side is a brute-force loop
    frac = parity.check_trace(b, o, parity.random_rays(sd, n))
    assert 0.02 < frac <= 1.0
    parity.check_trace(b, o, parity.camera_rays(o, 1, 5, 4, 2048))


def test_traversal_counters_and_device_buffers(case):
    sd, b, o, _, _ = case
    rays = parity.random_rays(sd, 4096)
    cnt = np.zeros(4, np.uint64)
    h1 = b.trace(rays, counters=cnt)
    assert cnt[2] == 4096 and cnt[0] >= 4096 and cnt[1] > 0
    d_rays = torch.from_numpy(rays).cuda()
    h2 = b.scene.trace(d_rays)
    b.scene.synchronize()
    assert np.array_equal(h1, h2.cpu().numpy().view(np.uint32))


def test_ray_generation(case):
    sd, b, o, _, _ = case
    a, r = b.sample_rays(1, 9, 4, 0, 4096), o.sample_rays(1, 9, 4, 0, 4096)
    assert np.allclose(a, r, rtol=1e-6, atol=1e-7)


def test_primal_film(case):
    sd, b, o, _, max_depth = case
    o.set_accel(True)
    ref = o.render_raw(1, 3, 4, max_depth, 5)
    parity.check_film(b.render_raw(1, 3, 4, max_depth, 5), ref)
    o.set_accel(False)


def test_render_entry_point_matches_reference_layout(case):
    import epsm_mitsuba3_b200 as ep
    sd, b, o, variant, max_depth = case
    integ = ep.load_dict({"type": variant, "max_depth": max_depth})
    img = integ.render(b.scene, sensor=1, seed=3, spp=4)
    assert img.shape == (sd.sensors[1].height, sd.sensors[1].width, 5) and img.dtype == torch.float32
    img = img.cpu().numpy()
    assert np.all(img[..., 3:] == 0)                              # epsm.py:79-80
    o.set_accel(True)
    ref = o.render(1, 3, 4, max_depth, 5)
    o.set_accel(False)
    rms = np.sqrt(np.mean(ref[..., :3].astype(np.float64) ** 2))
    assert np.sqrt(np.mean((img[..., :3] - ref[..., :3]).astype(np.float64) ** 2)) <= 1e-5 * rms


def test_per_path_gradients_and_scatter(case):
    sd, b, o, variant, max_depth = case
    o.set_accel(True)
    res = sd.sensors[2].width
    g = parity.synthetic_grad_in(res)
    n = res * res * 8
    paths = o.path_log(2, 11, 8, max_depth, 5, g)
    rec = b.path_records(variant, 2, 11, 8, max_depth, 5, g, 0, n)
    stats, (pg, lg, dg) = parity.check_path_records(rec, paths, variant, max_depth)
    parity.record_stats("gpu_path_records_vs_oracle", f"{variant}/{sd.name}/{sd.n_triangles}tris", stats)
    assert stats["nonzero"] > 0
    ref_p, ref_n = parity.scatter_gradients(paths, variant, parity.n_logged(max_depth), pg, lg, dg, sd.n_vertices, max_depth)
    gp, gn = b.backward(variant, 2, 11, 8, max_depth, 5, g)
    parity.check_accumulated(gp, gn, ref_p, ref_n)
    # device-resident buffers give the same result
    gd = torch.from_numpy(g).cuda()
    gp2, gn2 = b.scene.render_backward_raw(variant, 2, 11, 8, max_depth, 5, gd)
    b.scene.synchronize()
    parity.check_accumulated(gp2.cpu().numpy(), gn2.cpu().numpy(), ref_p, ref_n)
    o.set_accel(False)


def test_pruned_replay_equals_full_replay(case):
    """The default (gradient-liveness pruned) pass against epsm_set_full_replay(1): same gradients up to the order of
    the float atomics."""
    sd, b, o, variant, max_depth = case
    res = sd.sensors[2].width; g = parity.synthetic_grad_in(res)
    for md in sorted({max_depth, 3, 8}):
        b.scene.set_full_replay(True)
        full = b.backward(variant, 2, 13, 8, md, 5, g)
        b.scene.set_full_replay(False)
        pruned = b.backward(variant, 2, 13, 8, md, 5, g)
        for x, y in zip(full, pruned):
            assert np.allclose(x, y, rtol=1e-5, atol=1e-7 * max(1.0, float(np.abs(x).max()))), f"max_depth={md}"


def test_lane_ranges_partition_the_wavefront(case):
    sd, b, o, variant, max_depth = case
    res = sd.sensors[2].width; g = parity.synthetic_grad_in(res); n = res * res * 8
    full = b.backward(variant, 2, 5, 8, max_depth, 5, g)
    cut = (res // 8) * 4 * res * 8                               # a tile-row boundary, as band_lanes produces
    a1 = b.backward(variant, 2, 5, 8, max_depth, 5, g, 0, cut)
    a2 = b.backward(variant, 2, 5, 8, max_depth, 5, g, cut, n)
    assert np.allclose(a1[0] + a2[0], full[0], rtol=1e-4, atol=1e-6) and np.allclose(a1[1] + a2[1], full[1], rtol=1e-4, atol=1e-6)
    # an arbitrary (non-tile-aligned) split must work too
    a1 = b.backward(variant, 2, 5, 8, max_depth, 5, g, 0, n // 3)
    a2 = b.backward(variant, 2, 5, 8, max_depth, 5, g, n // 3, n)
    assert np.allclose(a1[0] + a2[0], full[0], rtol=1e-4, atol=1e-6)


def test_update_vertices_rebuilds_and_recomputes_normals(orc):
    sd = S.mirror_plane_scene(16, 2, 8)
    b, o = GpuBackend(sd), orc.OracleScene(sd)
    p = sd.meshes[0].positions.copy(); p[:, 2] += 0.3 * np.sin(p[:, 0])
    b.update_vertices(0, p); o.update_vertices(0, p)
    parity.check_trace(b, o, parity.random_rays(sd, 5000))
    # vertex normals were recomputed on the device (mesh.cpp:257-330): compare shading through a render
    sd2 = S.mirror_plane_scene(16, 2, 8)
    sd2.meshes[0].positions = p.copy(); sd2.meshes[0].normals = S.vertex_normals(p, sd2.meshes[0].faces)
    o2 = orc.OracleScene(sd2, accel=True)
    ref = o2.render_raw(1, 3, 2, 6, 5)
    got = b.render_raw(1, 3, 2, 6, 5)
    rms = np.sqrt(np.mean(ref.astype(np.float64) ** 2))
    assert np.sqrt(np.mean((got - ref).astype(np.float64) ** 2)) <= 2e-3 * rms      # normals differ in the last bits → rare path changes


def test_errors_do_not_cross_the_abi_as_crashes():
    b = GpuBackend(S.cornell_box(8, 1, 8))
    lib, ctx = b.scene.lib, b.scene.ctx
    out = np.zeros((8, 8, 5), np.float32)
    assert lib.epsm_render(ctx, 99, 0, 1, 6, 5, out.ctypes.data, 0) == -1
    assert b"sensor" in lib.epsm_last_error(ctx)
    assert lib.epsm_render(ctx, 1, 0, 1, 6, 0, out.ctypes.data, 0) == -1           # rr_depth must be > 0 (common.py:40-41)
    assert lib.epsm_render_backward(ctx, 7, 2, 0, 8, 6, 5, out.ctypes.data, out.ctypes.data, out.ctypes.data, 0, 0, 0, None) == -1


def test_differentiable_render_op_accumulates_into_params():
    import epsm_mitsuba3_b200 as ep
    sd = S.mirror_plane_scene(32, 4, 16)
    scene = ep.Scene(sd, 0)
    params = ep.traverse(scene)
    key = "mirror.vertex_positions"
    params[key] = params[key].clone().requires_grad_(True)
    params.update()
    integ = ep.load_dict({"type": "manifold", "max_depth": 6})
    img = ep.render(scene, params, sensor=1, integrator=integ, seed=4, spp=4)
    g = torch.from_numpy(np.tile(parity.synthetic_grad_in(16), (2, 2, 1))).cuda()   # optim.py tiles the low-res gradient
    (img * g).sum().backward()
    grad = params[key].grad
    assert grad is not None and grad.shape == params[key].shape and torch.isfinite(grad).all() and grad.abs().sum() > 0
    # same numbers as the raw entry point with seed_grad = TEA(seed, 1)[0]
    gp, gn = scene.render_backward_raw("manifold", 2, ep.sample_tea_32(4, 1)[0], 8, 6, 5, g[:16, :16].contiguous())
    scene.synchronize()
    # the mirror is smooth-shaded: its normals are re-derived from the positions by params.update() (mesh.cpp:82-87), so the
    # sh_frame.n seed continues into the positions — grad = grad_pos + Jᵀ grad_nrm (tests/test_vertex_normals.py holds Jᵀ)
    scene.vertex_normals_backward(gn, gp); scene.synchronize()
    v0, v1 = int(scene.voff[0]), int(scene.voff[1])
    assert torch.allclose(grad, gp[v0:v1], rtol=1e-4, atol=1e-6 * float(gp.abs().max()))


def test_primal_render_at_full_film_size_conserves_filter_weight():
    """BASELINE-size film (1024², more than one wavefront chunk): every sample splats its whole 5×5 footprint — also the
    samples whose fp32 position px + ξ rounds up to px + 1 (≈ 3e-5 of them at px ≈ 1000; they once fell off the tile-local
    film block).  Size-independent property: Σ film weight = Σ over samples of the footprint weight inside the film."""
    import epsm_mitsuba3_b200 as ep
    sd = S.cornell_box(1024, 8, 64)
    scene = ep.Scene(sd, device=0)
    scene.film_clear(1)
    scene.render_accumulate(1, 3, 8, 6, 5)
    scene.synchronize()
    film = scene.film_raw(1).cpu().numpy().astype(np.float64)
    assert np.isfinite(film).all()
    inner = film[8:-8, 8:-8, 3]
    assert abs(inner.mean() / film[..., 3].mean() - 1) < 0.01 and inner.min() > 0       # only the cropped border loses weight
    o_total = film[..., 3].sum()
    # second seed: same expected total within sampling noise of the border pixels only
    scene.film_clear(1); scene.render_accumulate(1, 4, 8, 6, 5); scene.synchronize()
    assert abs(scene.film_raw(1).cpu().numpy().astype(np.float64)[..., 3].sum() / o_total - 1) < 1e-3
    # and the single-kernel variant agrees on the weight channel to accumulation-order precision
    img = ep.load_dict({"type": "manifold"}).render(scene, sensor=1, seed=3, spp=8)
    assert tuple(img.shape) == (1024, 1024, 5) and bool(torch.isfinite(img).all())


def test_optimisation_loop_recovers_a_translation():
    """End to end, the way EPSM/optim.py drives the path: params.update() (device refit) → render → Matcher.match_Sinkhorn →
    backward of sum(img · grad) → Adam.  The wall seen in the mirror starts 0.29 units off; the matcher's pixel-motion gradient
    through the manifold integrator's vertex gradients must bring it back (descent, not parity: no reference run exists)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("optim_translation", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                                      "examples", "optim_translation.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    h = mod.run("mirror", iters=40, res=128, spp=16, match_res=64, quiet=True)
    e0 = float(np.linalg.norm([0.25, -0.15, 0.0])); e1 = float(np.mean([r["err"] for r in h[-5:]]))
    assert e1 < 0.5 * e0, (e0, e1)
    assert np.mean([r["sinkhorn"] for r in h[-5:]]) < 0.7 * np.mean([r["sinkhorn"] for r in h[:3]])
