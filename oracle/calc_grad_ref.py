"""Gradient oracle — TEST INFRASTRUCTURE ONLY (see oracle/epsm_oracle.cpp header for the rules).

Restates `ManifoldIntegrator.calc_grad` (src/python/python/ad/integrators/epsm.py:745-946) and
`ManifoldCausticIntegrator.calc_grad` (epsm.py:952-1200) with the reference's own numerical method:
every constraint row is obtained by reverse-mode autograd through the half-vector expression, the
system `constraint[:, :2id, 2:2id+2]` is inverted densely with `torch.linalg.inv`, and
`dL/dθ = dlduv[:, :, :2id] · (−C⁻¹ · ∂c/∂θ)` is accumulated with the reference's masks,
`nan_to_num` and the final ±0.1 outlier zeroing.  Runs on CPU (the reference hard-codes `.cuda()`,
epsm.py:768-773) in float32 or float64.

PINNED TO THE REFERENCE ITSELF: the reference has no test, golden vector or fixture for calc_grad
(SURVEY.md F5, §8c), but its calc_grad source needs only torch, so tests/golden/make_calc_grad_vectors.py
`exec`s epsm.py:744-946 / :951-1200 where they lie under /root/reference (with `.cuda()` and the four
`mi.` symbols stubbed) on the path logs of the parity scenes and commits what the reference returns
(tests/golden/calc_grad_reference.npz, float32 and float64).  tests/test_calc_grad_pinned.py holds this
restatement to those vectors (float64: 1e-12 relative) and, where /root/reference is present, re-executes
the reference live.  The restatement uses the reference's own method (autograd rows + dense inverse), so it
is methodologically independent of the CUDA path (analytic Jacobians + block solves).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

BF_NULL, BF_DIFFUSE = 0x1, 0x6


# epsm.py:746-762 (both classes define the same helpers)
def _orthogonal(nrm):
    v = torch.cat([torch.zeros_like(nrm[:, 0:1]), -nrm[:, 2:], nrm[:, 1:2]], dim=-1)
    return v / torch.norm(v, dim=-1, keepdim=True)


def _local_frame(n):
    nn = n / torch.norm(n, dim=-1, keepdim=True)
    t = _orthogonal(nn)
    b = torch.cross(nn, t, dim=-1)
    return torch.stack([t, b, nn], dim=1)            # rows t, b, n  (N,3,3)


def _bary(a, b, c, u0, u1):
    return a * u0[..., None] + b * u1[..., None] + c * (1 - u0 - u1)[..., None]


class _Vertex:
    """Leaf tensors of one logged vertex (the dict built at epsm.py:648-654)."""

    def __init__(self, rec, dtype):
        f = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=dtype)
        self.active = torch.tensor(rec["active"] > 0)
        self.flags = torch.tensor(rec["bsdf_flags"].astype(np.int64))
        self.ismesh = torch.tensor(rec["ismesh"] > 0)
        self.active_em = torch.tensor(rec["active_em"] > 0)
        self.light = f(rec["light"])
        self.p = [f(rec["p0"]), f(rec["p1"]), f(rec["p2"])]
        self.uv = [f(rec["b0"]), f(rec["b1"])]
        self.n = [f(rec["n0"]), f(rec["n1"]), f(rec["n2"])]
        self.eta = f(rec["eta"])
        self.hf = f(rec["hf"])

    def point(self):
        return _bary(self.p[0], self.p[1], self.p[2], self.uv[0], self.uv[1])

    def normal(self):
        return _bary(self.n[0], self.n[1], self.n[2], self.uv[0], self.uv[1])


def _halfvector(frame, prev, cur, nxt, eta):
    wi = prev - cur
    wo = nxt - cur
    wi = wi / torch.norm(wi, dim=-1, keepdim=True)
    wo = wo / torch.norm(wo, dim=-1, keepdim=True)
    wi2 = torch.bmm(frame, wi[..., None])[..., 0]
    wo2 = torch.bmm(frame, wo[..., None])[..., 0]
    res = wi2 + wo2 * eta[..., None]
    return res / torch.norm(res, dim=-1, keepdim=True), wo2


def calc_grad(paths: np.ndarray, variant: str, n_logged: int, dtype=torch.float32, keep_dtype: bool = False):
    """paths: structured array from OracleScene.path_log; n_logged = len(path_info) − 1.

    Returns (path_grad [5 per vertex (manifold) / 5 per vertex but 3 for the last (caustic)],
             light_grad [n_logged], diffuse_grad [n_logged]) as lists of (N,3) numpy arrays,
    exactly the three lists the reference's calc_grad returns."""
    assert variant in ("manifold", "manifold_caustic")
    caustic = variant == "manifold_caustic"
    N = paths.shape[0]
    Lp = n_logged + 1                                   # len(path_info)
    V: Dict[int, _Vertex] = {k: _Vertex(paths["v"][:, k - 1], dtype) for k in range(1, Lp)}
    cam = torch.tensor(np.ascontiguousarray(paths["cam"]), dtype=dtype)
    dlduv = torch.zeros((N, 1, 2 * Lp), dtype=dtype)                      # epsm.py:256
    dlduv[:, 0, 0:2] = torch.tensor(np.ascontiguousarray(paths["dlduv"]), dtype=dtype)   # epsm.py:268-269
    dldp1 = torch.tensor(np.ascontiguousarray(paths["dldp1"]), dtype=dtype).clone()       # epsm.py:270

    constraint = torch.zeros((N, 2 * Lp, 2 * Lp), dtype=dtype)
    diffuse_pos = torch.zeros(N, dtype=dtype)
    hasdiffuse = torch.zeros(N, dtype=dtype)
    params: List[torch.Tensor] = []
    rows: List[torch.Tensor] = []                        # param_grad_list
    final: List[torch.Tensor] = []
    light_grad: List[torch.Tensor] = []
    diffuse_grad: List[torch.Tensor] = []

    def add(v):
        v.requires_grad_(True); v.retain_grad()
        params.append(v); rows.append(torch.zeros((N, 2 * Lp, 3), dtype=dtype)); final.append(torch.zeros((N, 3), dtype=dtype))

    def take(t):
        g = t.grad; t.grad = None; return g

    eye = lambda k: torch.eye(k, dtype=dtype)
    valid = None
    for i_d in range(1, Lp):
        v = V[i_d]
        isdiffuse = (v.flags & BF_DIFFUSE) > 0
        for u in v.uv:
            u.requires_grad_(True); u.retain_grad()
        for p in v.p:
            add(p)
        if i_d == 1:
            prev = cam
            dldp1[~isdiffuse] = 0
            if caustic:
                dlduv[~isdiffuse] = 0                                       # epsm.py:999
            diffuse_grad.append(dldp1)
            valid = v.ismesh.clone()
        else:
            prev = V[i_d - 1].point()
            valid = valid & v.ismesh
        hasdiffuse = hasdiffuse + isdiffuse.to(dtype)
        valid = valid & (hasdiffuse < 2)
        diffuse_pos[isdiffuse] = i_d
        nolight = ~v.active_em
        cur = v.point()
        # ---- section A: connect to the sampled light point (epsm.py:804-866 / 1012-1094) ----
        v.light.requires_grad_(True); v.light.retain_grad()
        nxt = v.light
        light_rows = torch.zeros((N, 2 * Lp, 3), dtype=dtype)
        n = v.normal()
        if not caustic:
            m = v.hf
            add(n); add(m)
            frame = _local_frame(n)
        else:
            frame = _local_frame(n).detach()                                # epsm.py:1022
        res, wo2 = _halfvector(frame, prev, cur, nxt, v.eta)
        res2 = wo2 - wo2.detach()
        r0 = 2 * i_d - 2
        for i in range(2):
            res[:, i].sum().backward(retain_graph=True)
            if i_d > 1:
                constraint[:, r0 + i, 2 * i_d - 2] = take(V[i_d - 1].uv[0])
                constraint[:, r0 + i, 2 * i_d - 1] = take(V[i_d - 1].uv[1])
            constraint[:, r0 + i, 2 * i_d + 0] = take(v.uv[0])
            constraint[:, r0 + i, 2 * i_d + 1] = take(v.uv[1])
            for idx, para in enumerate(params):
                if para.grad is not None:
                    rows[idx][:, r0 + i, :] = take(para)
                else:
                    rows[idx][:, r0 + i, :] = 0
            light_rows[:, r0 + i, :] = take(nxt)
            if caustic:                                                     # epsm.py:1051-1070
                res2[:, i].sum().backward(retain_graph=True)
                for j in range(1, i_d + 1):
                    sel = diffuse_pos == j
                    constraint[sel, 2 * j - 2 + i] = 0
                    constraint[sel, 2 * j - 2 + i, 2 * i_d + 0] = v.uv[0].grad[sel]
                    constraint[sel, 2 * j - 2 + i, 2 * i_d + 1] = v.uv[1].grad[sel]
                    for idx, para in enumerate(params):
                        if para.grad is not None:
                            rows[idx][sel, 2 * j - 2 + i, :] = para.grad[sel]
                        else:
                            rows[idx][sel, 2 * j - 2 + i, :] = 0
                    if nxt.grad is not None:
                        light_rows[sel, 2 * j - 2 + i, :] = nxt.grad[sel]
                        nxt.grad = None
                    else:
                        light_rows[sel, 2 * j - 2 + i, :] = 0
                v.uv[0].grad = None; v.uv[1].grad = None
                for para in params:
                    para.grad = None
        C = constraint[:, :2 * i_d, 2:2 * i_d + 2].clone()
        C[~valid] = eye(2 * i_d)
        C[~v.active] = eye(2 * i_d)
        C[nolight] = eye(2 * i_d)
        Cinv = torch.linalg.inv(C)
        for idx, prow in enumerate(rows):
            duvdp = -torch.bmm(Cinv, prow[:, :2 * i_d, :])
            g = torch.bmm(dlduv[..., :2 * i_d], duvdp)[:, 0, :]
            g[~valid] = 0; g[~v.active] = 0; g[nolight] = 0; g[hasdiffuse > 0] = 0
            final[idx] += torch.nan_to_num(g)
        duvdlp = -torch.bmm(Cinv, light_rows[:, :2 * i_d, :])
        g = torch.bmm(dlduv[..., :2 * i_d], duvdlp)[:, 0, :]
        g[~valid] = 0; g[hasdiffuse > 0] = 0; g[~v.active] = 0; g[nolight] = 0
        light_grad.append(torch.nan_to_num(g))

        # ---- section B: connect to the next path vertex (epsm.py:868-930 / 1096-1184) ----
        if i_d < Lp - 1:
            w = V[i_d + 1]
            for u in w.uv:
                u.requires_grad_(True); u.retain_grad()
            nxt = w.point()
            if caustic:
                n = v.normal(); m = v.hf
                add(n); add(m)
            nxt.retain_grad()
            frame = _local_frame(n)
            res, wo2 = _halfvector(frame, prev, cur, nxt, v.eta)
            res = res - m
            res2 = wo2 - wo2.detach()
            next_rows = torch.zeros((N, 2 * Lp, 3), dtype=dtype)
            for i in range(2):
                res[:, i].sum().backward(retain_graph=True)
                if i_d > 1:
                    constraint[:, r0 + i, 2 * i_d - 2] = take(V[i_d - 1].uv[0])
                    constraint[:, r0 + i, 2 * i_d - 1] = take(V[i_d - 1].uv[1])
                constraint[:, r0 + i, 2 * i_d + 0] = take(v.uv[0])
                constraint[:, r0 + i, 2 * i_d + 1] = take(v.uv[1])
                constraint[:, r0 + i, 2 * i_d + 2] = take(w.uv[0])
                constraint[:, r0 + i, 2 * i_d + 3] = take(w.uv[1])
                for idx, para in enumerate(params):
                    if para.grad is not None:
                        rows[idx][:, r0 + i, :] = take(para)
                next_rows[:, r0 + i, :] = take(nxt)
                if caustic:                                                 # epsm.py:1141-1163
                    res2[:, i].sum().backward(retain_graph=True)
                    for j in range(1, i_d + 1):
                        sel = diffuse_pos == j
                        constraint[sel, 2 * j - 2 + i] = 0
                        constraint[sel, 2 * j - 2 + i, 2 * i_d + 0] = v.uv[0].grad[sel]
                        constraint[sel, 2 * j - 2 + i, 2 * i_d + 1] = v.uv[1].grad[sel]
                        constraint[sel, 2 * j - 2 + i, 2 * i_d + 2] = w.uv[0].grad[sel]
                        constraint[sel, 2 * j - 2 + i, 2 * i_d + 3] = w.uv[1].grad[sel]
                        for idx, para in enumerate(params):
                            if para.grad is not None:
                                rows[idx][sel, 2 * j - 2 + i, :] = para.grad[sel]
                            else:
                                rows[idx][sel, 2 * j - 2 + i, :] = 0
                        if nxt.grad is not None:
                            next_rows[sel, 2 * j - 2 + i, :] = nxt.grad[sel]
                        else:
                            next_rows[sel, 2 * j - 2 + i, :] = 0
                    v.uv[0].grad = None; v.uv[1].grad = None
                    w.uv[0].grad = None; w.uv[1].grad = None
                    for para in params:
                        para.grad = None
            C = constraint[:, :2 * i_d, 2:2 * i_d + 2].clone()
            C[~valid] = eye(2 * i_d)
            C[~w.active] = eye(2 * i_d)
            Cinv = torch.linalg.inv(C)
            next_diffuse = (w.flags & BF_DIFFUSE) > 0
            next_null = (w.flags & BF_NULL) > 0
            for idx, prow in enumerate(rows):
                duvdp = -torch.bmm(Cinv, prow[:, :2 * i_d, :])
                g = torch.bmm(dlduv[..., :2 * i_d], duvdp)[:, 0, :]
                g[~valid] = 0; g[~w.active] = 0; g[~next_diffuse] = 0
                g = torch.nan_to_num(g)
                if not caustic:
                    g[hasdiffuse > 0] = 0                                   # epsm.py:920 (absent at :1169-1176)
                final[idx] += g
            duvddp = -torch.bmm(Cinv, next_rows[:, :2 * i_d, :])
            g = torch.bmm(dlduv[..., :2 * i_d], duvddp)[:, 0, :]
            g[~valid] = 0; g[~w.active] = 0
            if caustic:
                g[(~next_null) & (~next_diffuse)] = 0                       # epsm.py:1182
            else:
                g[~next_diffuse] = 0; g[hasdiffuse > 0] = 0                 # epsm.py:927-928
            diffuse_grad.append(torch.nan_to_num(g))

    def outlier(t):                                                         # epsm.py:933-944 / 1187-1198
        t = t.detach().clone()
        t[t > 0.1] = 0; t[t < -0.1] = 0
        return t.numpy() if keep_dtype else t.to(torch.float32).numpy()

    return [outlier(t) for t in final], [outlier(t) for t in light_grad], [outlier(t) for t in diffuse_grad]


def scatter_gradients(paths: np.ndarray, variant: str, n_logged: int, path_grad, light_grad, diffuse_grad,
                      n_vertices: int, max_depth: int):
    """The Backward-mode replay of sample_path (epsm.py:559-562, 609-627, 644-645) as NumPy scatter-adds.
    Returns (grad_pos, grad_nrm), each (n_vertices, 3) float64.  Only meshes with gradient tracking receive
    values.  Seeds: (i) si.p0..p2 · g0..g2, (ii) si_follow.p · diffuse_grad,
    (iii) si_direct.p · light_grad · ΣLr_dir (+ the max_depth ≤ 3 shadow term), (iv) si_follow.sh_frame.n · g3
    (bsdf_sample.hf · g4 has no parameter dependence in the EPSM fork: microfacet.h:245 samples D(m) irrespective of wi)."""
    MESH_GRAD = 2
    gp = np.zeros((n_vertices, 3), np.float64); gn = np.zeros((n_vertices, 3), np.float64)
    n_pg, n_lg, n_dg = len(path_grad), len(light_grad), len(diffuse_grad)
    for k in range(n_logged):
        v = paths["v"][:, k]
        hit = v["active"] > 0
        grad_on = (v["mesh_flags"] & MESH_GRAD) > 0
        sel = hit & grad_on
        fi = v["fi"].astype(np.int64); bw = v["bw"].astype(np.float64)
        if 5 * k + 4 < n_pg:
            for c in range(3):                                              # (i)
                np.add.at(gp, fi[sel, c], path_grad[5 * k + c][sel].astype(np.float64))
            g3 = path_grad[5 * k + 3].astype(np.float64)                    # (iv)
            has_n = v["has_normals"] > 0
            s = sel & has_n
            n0, n1, n2 = (v[x].astype(np.float64) for x in ("n0", "n1", "n2"))
            # si.n0.. are zero for non-ismesh shapes in the log; FollowShape normals come from the mesh itself,
            # so recompute from the interpolated shading normal: d normalize(ñ)/dñ = (I − n nᵀ)/|ñ|
            nt = n0 * bw[:, 0:1] + n1 * bw[:, 1:2] + n2 * bw[:, 2:3]
            ln = np.linalg.norm(nt, axis=1, keepdims=True); ln[ln == 0] = 1
            nh = nt / ln
            flip = np.where((v["mesh_flags"] & 4) > 0, -1.0, 1.0)[:, None]
            gt = flip * (g3 - nh * (nh * g3).sum(1, keepdims=True)) / ln
            for c in range(3):
                np.add.at(gn, fi[s, c], gt[s] * bw[s, c:c + 1])
            # flat-shaded meshes: sh_frame.n = normalize(cross(p1-p0, p2-p0)) → positions (mesh.cpp:729,811)
            s = sel & ~has_n & (v["ismesh"] > 0)
            if s.any():
                p0, p1, p2 = (v[x].astype(np.float64) for x in ("p0", "p1", "p2"))
                e1, e2 = p1 - p0, p2 - p0
                c_ = np.cross(e1, e2); lc = np.linalg.norm(c_, axis=1, keepdims=True); lc[lc == 0] = 1
                nh = c_ / lc
                flip = np.where((v["mesh_flags"] & 4) > 0, -1.0, 1.0)[:, None]
                gc = flip * (g3 - nh * (nh * g3).sum(1, keepdims=True)) / lc
                ge1 = np.cross(e2, gc); ge2 = np.cross(gc, e1)
                np.add.at(gp, fi[s, 1], ge1[s]); np.add.at(gp, fi[s, 2], ge2[s]); np.add.at(gp, fi[s, 0], -(ge1 + ge2)[s])
        if k < n_dg:                                                        # (ii)
            g = diffuse_grad[k].astype(np.float64)
            for c in range(3):
                np.add.at(gp, fi[sel, c], g[sel] * bw[sel, c:c + 1])
        dfi = v["direct_fi"].astype(np.int64); dbw = v["direct_bw"].astype(np.float64)
        d_grad = (v["direct_any"] & 0x100) > 0
        if k == 0 and min(max_depth, 6) <= 3 and k < n_dg:                  # epsm.py:609-620
            s = ((v["direct_any"] & 1) > 0) & d_grad
            g = diffuse_grad[0].astype(np.float64) * v["direct_dis"].astype(np.float64)[:, None]
            for c in range(3):
                np.add.at(gp, dfi[s, c], g[s] * dbw[s, c:c + 1])
        if k < n_lg:                                                        # (iii) epsm.py:622-627
            s = (v["direct_valid"] > 0) & d_grad
            g = light_grad[k].astype(np.float64) * v["lr_dir_sum"].astype(np.float64)[:, None]
            for c in range(3):
                np.add.at(gp, dfi[s, c], g[s] * dbw[s, c:c + 1])
    return gp, gn
