"""TEST INFRASTRUCTURE (checker only — nothing in epsm_mitsuba3_b200 imports this).

CPU restatement of the matcher that produces `grad_in` for the manifold integrators:
`Matcher.match_Sinkhorn` (EPSM/utils/matcher.py:51-63) = geomloss `SamplesLoss("sinkhorn", blur=0.01, scaling=0.9)`
on two 5-D point clouds (clamped rgb + pixel grid), differentiated w.r.t. the rendered cloud.

PARITY UNPINNED.  geomloss is a third-party dependency that is absent from /root/reference and from this image
(README.md:11 `pip install geomloss[full]`, no version pin), and the reference holds no golden vector at this boundary
(SURVEY §8c).  What follows restates the published algorithm of geomloss 0.2.x (`geomloss/sinkhorn_divergence.py`:
`epsilon_schedule`, `scaling_parameters`, `sinkhorn_loop`, `sinkhorn_cost`; `geomloss/sinkhorn_samples.py`:
`softmin_tensorized`, `sinkhorn_tensorized`; `geomloss/utils.py`: `squared_distances`, `max_diameter`) for the
configuration the reference uses — p = 2, debias = True, reach = None (balanced, damping 1), uniform weights — in torch
on the CPU, dense N×N, with autograd providing the gradient exactly as the reference obtains it
(`torch.autograd.grad(sum(loss)·res², [render_point_5d])`, matcher.py:60).  float64 by default.

The analytic form the CUDA kernel uses is restated next to it (`grad_barycentric`) and the tests hold the two to each
other: with the dual potentials detached before the last extrapolation step (sinkhorn_loop's `.detach()`), and x detached
on the column side of every cost matrix (`C_xx = cost(x, x.detach())`, `C_yx = cost(y, x.detach())`), the gradient is
  g_i = N·a_i·[ Σ_j w^xx_ij x_j − Σ_j w^xy_ij y_j ],   w_ij = softmax_j(h_j − C_ij/ε)."""
import numpy as np
import torch


def grid_positions(res: int, dtype=torch.float64) -> torch.Tensor:
    """Matcher.__init__ (matcher.py:14-18): meshgrid(linspace, linspace) with 'ij' indexing; pos = (pos[1], pos[0])."""
    x = torch.linspace(0, 1, res, dtype=torch.float32).to(dtype)
    p0, p1 = torch.meshgrid(x, x, indexing="ij")
    return torch.cat([p1[..., None], p0[..., None]], dim=2).reshape(-1, 2)


def clouds(render_rgb, target_rgb, res, dtype=torch.float64, device="cpu"):
    """matcher.py:52-58: (r, g, b, x, y) with rgb clamped to [0, 1].  `device`: where the dense N×N restatement is
    evaluated — the production sizes (128² points = 2 GB per float64 cost matrix) are checked with the oracle on a CUDA
    device in float64; it stays the same torch code."""
    pos = grid_positions(res, dtype)
    r = torch.as_tensor(np.asarray(render_rgb), dtype=dtype).reshape(-1, 3).clamp(0, 1)
    t = torch.as_tensor(np.asarray(target_rgb), dtype=dtype).reshape(-1, 3).clamp(0, 1)
    return torch.cat([r, pos], 1).to(device), torch.cat([t, pos], 1).to(device)


def max_diameter(x, y):
    mins = torch.stack((x.min(dim=0)[0], y.min(dim=0)[0])).min(dim=0)[0]
    maxs = torch.stack((x.max(dim=0)[0], y.max(dim=0)[0])).max(dim=0)[0]
    return float((maxs - mins).norm())


def epsilon_schedule(p, diameter, blur, scaling):
    return ([diameter ** p]
            + [float(np.exp(e)) for e in np.arange(p * np.log(diameter), p * np.log(blur), p * np.log(scaling))]
            + [blur ** p])


def squared_distances(x, y):
    d_xx = (x * x).sum(-1)[:, None]
    d_yy = (y * y).sum(-1)[None, :]
    return d_xx - 2 * (x @ y.T) + d_yy


def cost(x, y):
    return squared_distances(x, y) / 2            # p = 2


def softmin(eps, c, h):
    return -eps * torch.logsumexp(h[None, :] - c / eps, dim=1)


def sinkhorn_potentials(x, y, blur=0.01, scaling=0.9):
    """sinkhorn_loop with debias, balanced, uniform weights.  Returns the four potentials after the last (attached)
    extrapolation, the final ε and the ε schedule."""
    n, m = x.shape[0], y.shape[0]
    a_log = torch.full((n,), -np.log(n), dtype=x.dtype, device=x.device); b_log = torch.full((m,), -np.log(m), dtype=x.dtype, device=x.device)
    diameter = max_diameter(x.detach(), y.detach())
    eps_list = epsilon_schedule(2, diameter, blur, scaling)
    c_xx, c_yy = cost(x, x.detach()), cost(y, y.detach())
    c_xy, c_yx = cost(x, y.detach()), cost(y, x.detach())
    with torch.no_grad():
        eps = eps_list[0]
        g_ab = softmin(eps, c_yx, a_log); f_ba = softmin(eps, c_xy, b_log)
        f_aa = softmin(eps, c_xx, a_log); g_bb = softmin(eps, c_yy, b_log)
        for eps in eps_list:
            ft_ba = softmin(eps, c_xy, b_log + g_ab / eps); gt_ab = softmin(eps, c_yx, a_log + f_ba / eps)
            ft_aa = softmin(eps, c_xx, a_log + f_aa / eps); gt_bb = softmin(eps, c_yy, b_log + g_bb / eps)
            f_ba, g_ab = 0.5 * (f_ba + ft_ba), 0.5 * (g_ab + gt_ab)
            f_aa, g_bb = 0.5 * (f_aa + ft_aa), 0.5 * (g_bb + gt_bb)
    pre = dict(f_ba=f_ba, g_ab=g_ab, f_aa=f_aa, g_bb=g_bb)
    # last extrapolation, gradients enabled, potentials detached
    f_ba2, g_ab2 = (softmin(eps, c_xy, (b_log + g_ab / eps).detach()), softmin(eps, c_yx, (a_log + f_ba / eps).detach()))
    f_aa2 = softmin(eps, c_xx, (a_log + f_aa / eps).detach()); g_bb2 = softmin(eps, c_yy, (b_log + g_bb / eps).detach())
    return dict(f_ba=f_ba2, g_ab=g_ab2, f_aa=f_aa2, g_bb=g_bb2), pre, eps, eps_list


def sinkhorn_loss(x, y, blur=0.01, scaling=0.9):
    pot, _, _, _ = sinkhorn_potentials(x, y, blur, scaling)
    n, m = x.shape[0], y.shape[0]
    return (pot["f_ba"] - pot["f_aa"]).sum() / n + (pot["g_ab"] - pot["g_bb"]).sum() / m       # sinkhorn_cost, debias


def match_sinkhorn(render_rgb, target_rgb, res, blur=0.01, scaling=0.9, dtype=torch.float64, device="cpu"):
    """Matcher.match_Sinkhorn: returns (g (N,5) float64 numpy, loss)."""
    x, y = clouds(render_rgb, target_rgb, res, dtype, device)
    x.requires_grad_(True)
    loss = sinkhorn_loss(x, y, blur, scaling)
    (g,) = torch.autograd.grad(loss * res * res, [x])
    return g.cpu().numpy(), float(loss.detach())


def grad_barycentric(render_rgb, target_rgb, res, blur=0.01, scaling=0.9, dtype=torch.float64):
    """The closed form of the same gradient (what the CUDA kernel evaluates)."""
    x, y = clouds(render_rgb, target_rgb, res, dtype)
    with torch.no_grad():
        _, pre, eps, _ = sinkhorn_potentials(x, y, blur, scaling)
        n, m = x.shape[0], y.shape[0]
        w_xy = torch.softmax((-np.log(m) + pre["g_ab"] / eps)[None, :] - cost(x, y) / eps, dim=1)
        w_xx = torch.softmax((-np.log(n) + pre["f_aa"] / eps)[None, :] - cost(x, x) / eps, dim=1)
        g = (w_xx @ x - w_xy @ y) * (res * res / n)
    return g.numpy()
