"""BASELINE.json's full sizes on the GPU (configs[2]: 1 M triangles, 1024², 32 spp; configs[1]: glass sphere, 512², 64 spp),
held through size-independent properties — the oracle cannot finish 33 M paths, so it checks samples of the same
wavefront and the rest is closed by identities the path must satisfy:

  * hit records of sampled rays bit-exact against the oracle (brute-force semantics through its own BVH);
  * per-path records of sampled lane windows of the FULL wavefront against the oracle + the reference's torch calc_grad;
  * additivity of the gradient pass over images with disjoint pixel support (each path reads its own pixel: epsm.py:250-257);
  * the pruned replay equals the full replay, lane / sample shards sum to the whole, the wavefront chunk size does not
    matter, refit equals rebuild — all on the full wavefront."""
import os

import numpy as np
import pytest
import torch

from epsm_mitsuba3_b200 import scenes as S
from tests import parity

pytestmark = pytest.mark.gpu

RES, SPP, MD = 1024, 32, 8


@pytest.fixture(scope="module")
def room(orc):
    import epsm_mitsuba3_b200 as ep
    sd = S.indoor_room(1_000_000, RES, SPP, RES)
    return sd, ep.Scene(sd, device=0), orc.OracleScene(sd, accel=True)


def close(x, y, rtol=2e-4, atol_rel=2e-6):
    x, y = x.double(), y.double()
    return bool(((x - y).abs() <= rtol * y.abs() + atol_rel * max(1.0, float(y.abs().max()))).all())


def backward(scene, variant, sensor, seed, spp, md, g):
    buf = torch.zeros(2, scene.n_vertices, 3, device="cuda"); torch.cuda.synchronize()
    scene.render_backward_raw(variant, sensor, seed, spp, md, 5, g, buf[0], buf[1]); scene.synchronize()
    return buf


def test_fullsize_scene_and_hits_bit_exact(room):
    sd, scene, o = room
    st = scene.stats()
    assert abs(st["n_triangles"] - 1_000_000) <= 10_000
    rays = np.concatenate([parity.random_rays(sd, 60000, seed=1), parity.camera_rays(o, 1, 3, SPP, 40000)])
    parity.check_trace(scene, o, rays)


def test_fullsize_path_records_of_sampled_windows(room):
    sd, scene, o = room
    g = parity.synthetic_grad_in(RES)
    total = RES * RES * SPP
    for lb in (0, (total // 2 // SPP) * SPP, total - 4096):
        le = lb + 4096
        paths = o.path_log(1, 21, SPP, MD, 5, g, lb, le)
        rec = scene.path_records("manifold", 1, 21, SPP, MD, 5, g, lb, le)
        stats, _ = parity.check_path_records(rec, paths, "manifold", MD)
        parity.record_stats("gpu_fullsize_window_vs_oracle", f"room/lanes{lb}-{le}", stats)
        assert stats["nonzero"] >= 0


def test_fullsize_gradient_pass_is_additive_over_pixels(room):
    """Each path reads grad_in at its own pixel only (epsm.py:250-257) and everything downstream is per path, so the pass
    is additive over images with disjoint pixel support — exactly, ±0.1 outlier zeroing (epsm.py:933-944) included
    (plain linearity in grad_in does not hold because of that zeroing)."""
    sd, scene, o = room
    g = torch.from_numpy(parity.synthetic_grad_in(RES, seed=7)).cuda()
    yy, xx = torch.meshgrid(torch.arange(RES, device="cuda"), torch.arange(RES, device="cuda"), indexing="ij")
    mask = (((yy // 8) + (xx // 8)) % 2 == 0)[..., None].float()
    whole = backward(scene, "manifold", 1, 31, SPP, MD, g)
    a, b = backward(scene, "manifold", 1, 31, SPP, MD, g * mask), backward(scene, "manifold", 1, 31, SPP, MD, g * (1 - mask))
    assert close(a + b, whole) and float(a.abs().sum()) > 0 and float(b.abs().sum()) > 0
    # and a zero image gradient reaches no parameter
    assert float(backward(scene, "manifold", 1, 31, SPP, MD, torch.zeros_like(g)).abs().sum()) == 0.0


def test_fullsize_pruned_equals_full_and_shards_sum(room):
    sd, scene, o = room
    g = torch.from_numpy(parity.synthetic_grad_in(RES)).cuda()
    whole = backward(scene, "manifold", 1, 41, SPP, MD, g)
    scene.set_full_replay(True); full = backward(scene, "manifold", 1, 41, SPP, MD, g); scene.set_full_replay(False)
    assert close(whole, full)
    acc = torch.zeros_like(whole)
    for r in range(4):                                          # sample shards of 4 ranks (epsm_set_sample_range)
        scene.set_sample_range(SPP * r // 4, SPP * (r + 1) // 4)
        acc += backward(scene, "manifold", 1, 41, SPP, MD, g)
    scene.set_sample_range(0, 0)
    assert close(acc, whole)


def test_fullsize_refit_equals_rebuild(room):
    sd, scene, o = room
    g = torch.from_numpy(parity.synthetic_grad_in(RES)).cuda()
    mid = max((i for i, m in enumerate(sd.meshes) if m.grad), key=lambda i: sd.meshes[i].positions.shape[0])
    p0 = sd.meshes[mid].positions.copy()
    p = (p0 + 0.02 * np.sin(5.0 * p0 + 1.0)).astype(np.float32)
    scene.update_vertices(mid, p, commit=False); scene.refit()
    a = backward(scene, "manifold", 1, 51, 8, MD, g)
    built, now, refits = scene.accel_cost()
    assert refits == 1 and now > 0
    scene.commit()
    b = backward(scene, "manifold", 1, 51, 8, MD, g)
    assert close(a, b) and float(b.abs().sum()) > 0
    o.update_vertices(mid, p)
    parity.check_trace(scene, o, parity.random_rays(sd, 40000, seed=5))
    scene.update_vertices(mid, p0); o.update_vertices(mid, p0)


def test_glass_config_chunk_size_does_not_matter():
    """configs[1] (512², 64 spp, manifold_caustic): the wavefront cut into 2²⁰-lane chunks equals the default 2²²."""
    import epsm_mitsuba3_b200 as ep
    sd = S.glass_sphere_caustic(512, 64, 512)
    g = torch.from_numpy(parity.synthetic_grad_in(512)).cuda()
    outs = []
    for chunk in (None, 1 << 20):
        if chunk:
            os.environ["EPSM_WF_CHUNK"] = str(chunk)
        try:
            scene = ep.Scene(sd, device=0)
            outs.append(backward(scene, "manifold_caustic", 1, 61, 64, 6, g))
        finally:
            os.environ.pop("EPSM_WF_CHUNK", None)
    assert close(outs[0], outs[1]) and float(outs[0].abs().sum()) > 0


@pytest.mark.parametrize("name,make,variant,res,spp,md", [
    ("glass", lambda: S.glass_sphere_caustic(512, 64, 512), "manifold_caustic", 512, 64, 6),          # BASELINE configs[1]
    ("stage", lambda: S.articulated_stage(512, 64, 512), "manifold_caustic", 512, 64, 3),             # BASELINE configs[3]: shadow_term branch
])
def test_fullsize_path_records_of_configs_1_and_3(orc, name, make, variant, res, spp, md):
    """Per-path records of lane windows of the FULL configs[1] / configs[3] wavefronts against the oracle + the reference's
    torch calc_grad (windows: image start, centre — where the object is — and end), and the accumulated gradient of a centre
    band of pixels against the oracle's scatter."""
    import epsm_mitsuba3_b200 as ep
    sd = make()
    scene, o = ep.Scene(sd, device=0), orc.OracleScene(sd, accel=True)
    g = parity.synthetic_grad_in(res)
    total = res * res * spp
    centre = ((res // 2) * res + res // 2 - 16) * spp
    n_grad = 0
    for lb in (0, centre, total - 4096):
        le = lb + 4096
        paths = o.path_log(1, 21, spp, md, 5, g, lb, le)
        rec = scene.path_records(variant, 1, 21, spp, md, 5, g, lb, le)
        # the centre window lies entirely on the object: every path carries a gradient and the share of ill-conditioned ones
        # (held to the float64 evaluation within the reference's own float32 error, parity.cmp_ill) is above the scene average
        # of ≤ 1 % — measured 1.7 % on `stage` (profiles/r02_e_summary.md)
        stats, _ = parity.check_path_records(rec, paths, variant, md, max_ill_frac=0.03)
        parity.record_stats("gpu_fullsize_window_vs_oracle", f"{name}/lanes{lb}-{le}", stats)
        n_grad += stats["nonzero"]
    assert n_grad > 0
    # accumulated gradients of a 4-row band through the image centre
    from oracle.calc_grad_ref import calc_grad, scatter_gradients
    lb, le = (res // 2 - 2) * res * spp, (res // 2 + 2) * res * spp
    paths = o.path_log(1, 23, spp, md, 5, g, lb, le)
    nl = min(min(md, 6), 5)
    pg, lg, dg = calc_grad(paths, variant, nl)
    rp, rn = scatter_gradients(paths, variant, nl, pg, lg, dg, sd.n_vertices, md)
    buf = torch.zeros(2, scene.n_vertices, 3, device="cuda"); torch.cuda.synchronize()
    scene.render_backward_raw(variant, 1, 23, spp, md, 5, torch.from_numpy(g).cuda(), buf[0], buf[1], lane_begin=lb, lane_end=le); scene.synchronize()
    parity.check_accumulated(buf[0].cpu().numpy(), buf[1].cpu().numpy(), rp, rn)
