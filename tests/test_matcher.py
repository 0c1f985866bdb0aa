"""The matcher at the tensor boundary: Matcher.match_Sinkhorn (EPSM/utils/matcher.py:51-63).

PARITY UNPINNED: geomloss is absent (unpinned third-party dependency, no golden vector in the reference).  The oracle is
the literal torch restatement of geomloss' published sinkhorn_loop (oracle/sinkhorn_ref.py) in float64; CPU tests hold
its autograd gradient to the closed form the kernel evaluates, GPU tests hold the kernel to the oracle."""
import numpy as np
import pytest
import torch

from oracle import sinkhorn_ref as R

# float32 kernel vs float64 oracle: 54 ε-steps down to ε = 1e-4, exponents up to ~1e4 → stated tolerance on g
G_RTOL = 2e-3          # relative to max |g|


def images(res, seed, shift=0.08):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:res, 0:res] / (res - 1)
    def blob(cx, cy):
        return np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / 0.02)
    a = np.stack([blob(0.4, 0.5), 0.5 * blob(0.6, 0.4), 0.2 + 0 * xx], -1) + rng.uniform(0, 0.05, (res, res, 3))
    b = np.stack([blob(0.4 + shift, 0.5), 0.5 * blob(0.6, 0.4 + shift), 0.2 + 0 * xx], -1) + rng.uniform(0, 0.05, (res, res, 3))
    a[0, 0] = (-0.3, 1.4, 0.5)                       # out-of-range values: the clamp of matcher.py:53,56
    return a.reshape(-1, 3).astype(np.float32), b.reshape(-1, 3).astype(np.float32)


def test_oracle_autograd_equals_closed_form():
    for res, seed in ((8, 0), (12, 1)):
        a, b = images(res, seed)
        g, loss = R.match_sinkhorn(a, b, res)
        assert np.allclose(g, R.grad_barycentric(a, b, res), rtol=0, atol=1e-12)
        assert loss > 0 and np.abs(g).max() > 0


def test_oracle_schedule_and_identities():
    a, b = images(8, 2)
    x, y = R.clouds(a, b, 8)
    eps = R.epsilon_schedule(2, R.max_diameter(x, y), 0.01, 0.9)
    assert eps[0] == pytest.approx(R.max_diameter(x, y) ** 2) and eps[-1] == pytest.approx(1e-4)
    assert all(e1 > e2 for e1, e2 in zip(eps[1:-1], eps[2:-1])) and 40 < len(eps) < 70
    g, loss = R.match_sinkhorn(a, a, 8)              # identical clouds: divergence and gradient vanish (debiasing)
    assert abs(loss) < 1e-12 and np.abs(g).max() < 1e-9
    # pixel grid layout of matcher.py:14-18: point index i·res + j ↦ (x, y) = (j, i)/(res − 1)
    pos = R.grid_positions(4).numpy()
    assert np.allclose(pos[1], (1 / 3, 0)) and np.allclose(pos[4], (0, 1 / 3))


@pytest.mark.gpu
@pytest.mark.parametrize("res", [8, 16, 32, 48])
def test_gpu_matcher_against_oracle(res):
    import epsm_mitsuba3_b200 as ep
    a, b = images(res, res)
    m = ep.Matcher(res, "cuda:0")
    g = m.match_Sinkhorn(torch.from_numpy(a), torch.from_numpy(b), want_loss=True).cpu().numpy()
    ref, loss = R.match_sinkhorn(a, b, res)
    x, y = R.clouds(a, b, res)
    assert m.n_steps == len(R.epsilon_schedule(2, R.max_diameter(x.float(), y.float()), 0.01, 0.9))
    scale = np.abs(ref).max()
    assert np.abs(g - ref).max() <= G_RTOL * scale, (np.abs(g - ref).max(), scale)
    assert m.last_loss == pytest.approx(loss, rel=2e-3, abs=1e-7)
    # geomloss itself would run this in float32 with the |x|² + |y|² − 2x·y distance: the kernel must be at least as close
    ref32, _ = R.match_sinkhorn(a, b, res, dtype=torch.float32)
    assert np.abs(g - ref).max() <= max(2.0 * np.abs(ref32 - ref).max(), 1e-4 * scale)


@pytest.mark.gpu
@pytest.mark.parametrize("res,oracle_device", [(64, "cpu"), (128, "cuda")])
def test_gpu_matcher_at_production_sizes(res, oracle_device):
    """match_res of the reference's experiments (EPSM/exp/*.py: 128; 64 as the step below).  At 128² the cloud spans
    sixteen 1024-column tiles, so the far-tile skip of k_sink_softmin is on the hot path.  The float64 dense oracle runs on
    the CPU at 64² and — the same torch code — on the CUDA device at 128² (2 GB per cost matrix)."""
    import epsm_mitsuba3_b200 as ep
    a, b = images(res, res)
    m = ep.Matcher(res, "cuda:0")
    g = m.match_Sinkhorn(torch.from_numpy(a), torch.from_numpy(b), want_loss=True).cpu().numpy()
    ref, loss = R.match_sinkhorn(a, b, res, device=oracle_device)
    scale = np.abs(ref).max()
    err = np.abs(g - ref).max()
    assert err <= G_RTOL * scale, (err, scale)
    assert m.last_loss == pytest.approx(loss, rel=2e-3, abs=1e-7)
    # per-row relative error where the gradient is not negligible: no tile is skipped that matters to any row
    big = np.abs(ref).max(1) > 0.05 * scale
    assert (np.abs(g - ref)[big].max(1) / np.abs(ref)[big].max(1)).max() < 0.05


@pytest.mark.gpu
def test_gpu_matcher_identities_and_host_buffers():
    import ctypes as C
    import epsm_mitsuba3_b200 as ep
    from epsm_mitsuba3_b200 import capi
    res = 16
    a, b = images(res, 5)
    m = ep.Matcher(res, "cuda:0")
    g_same = m.match_Sinkhorn(torch.from_numpy(a), torch.from_numpy(a), want_loss=True)
    assert float(g_same.abs().max()) < 1e-4 and abs(m.last_loss) < 1e-6
    g_dev = m.match_Sinkhorn(torch.from_numpy(a), torch.from_numpy(b)).cpu().numpy()
    # host pointers through the C ABI give the same numbers
    r, t = (z.cpu().numpy() for z in m.clouds(torch.from_numpy(a), torch.from_numpy(b)))
    out = np.zeros_like(r); loss = C.c_double()
    rc = m.lib.epsm_sinkhorn_grad(m.ctx, r.ctypes.data, r.shape[0], t.ctypes.data, t.shape[0], 0.01, 0.9, float(res * res), out.ctypes.data,
                                  C.byref(loss), None, capi.MEM_HOST)
    assert rc == 0 and np.array_equal(out, g_dev)
    # unequal cloud sizes (n ≠ m) run through the same kernels
    ref, _ = None, None
    x, y = R.clouds(a, b, res)
    y2 = y[: res * res // 2]
    xg = x.clone().requires_grad_(True)
    l2 = R.sinkhorn_loss(xg, y2)
    (gref,) = torch.autograd.grad(l2 * res * res, [xg])
    out2 = np.zeros_like(r)
    t2 = np.ascontiguousarray(t[: res * res // 2])
    assert m.lib.epsm_sinkhorn_grad(m.ctx, r.ctypes.data, r.shape[0], t2.ctypes.data, t2.shape[0], 0.01, 0.9, float(res * res), out2.ctypes.data,
                                    None, None, capi.MEM_HOST) == 0
    assert np.abs(out2 - gref.numpy()).max() <= G_RTOL * np.abs(gref.numpy()).max()
    # errors stay error codes
    assert m.lib.epsm_sinkhorn_grad(m.ctx, r.ctypes.data, 0, t.ctypes.data, 4, 0.01, 0.9, 1.0, None, None, None, capi.MEM_HOST) != 0
    assert m.lib.epsm_sinkhorn_grad(m.ctx, r.ctypes.data, 4, t.ctypes.data, 4, 0.01, 1.5, 1.0, None, None, None, capi.MEM_HOST) != 0
