"""CPU oracle of `prb_reparam`'s reverse-mode pass — TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's CPU legs).

Restates RBIntegrator.render_backward (src/python/python/ad/integrators/common.py:803-975) with
PRBReparamIntegrator.sample (prb_reparam.py:277-607), sample_rays (common.py:291-397) and the reparameterisation of
Bangaru et al. (src/python/python/ad/reparam.py:10-463).  Split in two:

  * the DETACHED half — the random walk, all sampled quantities, the auxiliary rays and their harmonic weights — comes from
    `orc_prb_log` in oracle/epsm_oracle.cpp (scalar float32 C++, same sampling code as the pinned primal oracle);
  * the ATTACHED half — everything the reference evaluates under `dr.resume_grad()` — is re-evaluated here on those records
    with torch autograd in float64, term by term in the reference's order, and back-propagated to the vertex buffers.

PARITY UNPINNED: the reference's AD variants cannot run in this image (no LLVM / OptiX) and its tests hold
no golden vector for this integrator (they compare against finite differences, src/python/python/tests/
test_ad_integrators.py); tests/test_prb_reparam.py does the same finite-difference check on the GPU.  The order in which
the two custom ops of one iteration consume the wrapper's random generator (emitter ray first) follows Dr.Jit's reverse
traversal order and is an assumption; it only permutes which auxiliary directions go to which ray.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import oracle as _o

F64 = torch.float64
AUX, HDR, IT = 8, 16, 72


def _filter_coeffs():
    """GaussianFilter (src/rfilters/gaussian.cpp:57-96), stddev 0.5, radius 2: scaled Remez polynomial in x²."""
    c = [9.992604880e-1, -4.977025247e-1, 1.222248550e-1, -1.932406282e-2, 2.136713061e-3,
         -1.679873860e-4, 9.202145248e-6, -3.329417433e-7, 7.128382794e-9, -6.821193280e-11]
    out, scale = [], 1.0
    for v in c:
        out.append(np.float32(v * scale)); scale /= 0.25
    x = np.float32(4.0)
    acc = np.float32(0)
    for v in reversed(out):
        acc = np.float32(acc * x + v)
    # the reference subtracts estrin(radius²) evaluated in float32; Horner differs by rounding only
    out[0] = np.float32(out[0] - acc)
    return [float(v) for v in out]


_COEFF = _filter_coeffs()


def filter_eval(x):
    x2 = x * x
    acc = torch.zeros_like(x2)
    for v in reversed(_COEFF):
        acc = acc * x2 + v
    return torch.clamp(acc, min=0.0)


def normalize(v):
    return v / torch.linalg.norm(v, dim=-1, keepdim=True)


def dot(a, b):
    return (a * b).sum(-1)


def coordinate_system(n):
    """include/mitsuba/core/vector.h:  Duff et al. branchless ONB (frame.h:21-26 uses it for Frame3f(n))."""
    sign = torch.where(torch.signbit(n[..., 2]), -torch.ones_like(n[..., 2]), torch.ones_like(n[..., 2]))
    a = -1.0 / (sign + n[..., 2])
    b = n[..., 0] * n[..., 1] * a
    s = torch.stack([sign * (n[..., 0] * n[..., 0] * a) + 1.0, sign * b, -sign * n[..., 0]], -1)
    t = torch.stack([b, sign + n[..., 1] * n[..., 1] * a, -n[..., 1]], -1)
    return s, t


def replace_grad(primal, attached):
    return primal.detach() + (attached - attached.detach())


class Geometry:
    """Vertex buffers as autograd leaves + the static tables of the scene."""

    def __init__(self, sd):
        self.sd = sd
        self.P = torch.tensor(np.concatenate([m.positions for m in sd.meshes]), dtype=F64, requires_grad=True)
        nrm = [m.normals if m.normals is not None else np.zeros_like(m.positions) for m in sd.meshes]
        self.N = torch.tensor(np.concatenate(nrm), dtype=F64, requires_grad=True)
        voff, foff = sd.vertex_offsets(), sd.face_offsets()
        self.faces = torch.tensor(np.concatenate([m.faces.astype(np.int64) + int(voff[i]) for i, m in enumerate(sd.meshes)]))
        fm = np.concatenate([np.full(m.faces.shape[0], i, np.int64) for i, m in enumerate(sd.meshes)])
        self.face_mesh = torch.tensor(fm)
        self.has_normals = torch.tensor([m.normals is not None for m in sd.meshes])
        self.flip = torch.tensor([bool(m.flip_normals) for m in sd.meshes])
        self.mesh_bsdf = torch.tensor([m.bsdf for m in sd.meshes])
        self.mesh_emitter = torch.tensor([m.emitter for m in sd.meshes])
        self.bsdfs = sd.bsdfs
        # vertex texture coordinates (constants: Mesh::vertex_texcoord is not a differentiable parameter here)
        uv = [np.asarray(m.texcoords, np.float64).reshape(-1, 2) if getattr(m, "texcoords", None) is not None else np.zeros((m.positions.shape[0], 2)) for m in sd.meshes]
        self.UV = torch.tensor(np.concatenate(uv), dtype=F64)
        self.has_uv = torch.tensor([getattr(m, "texcoords", None) is not None for m in sd.meshes])


def surface_interaction(G: Geometry, prim, b1, b2, valid, ray_o, ray_d, mode):
    """Mesh::compute_surface_interaction (src/render/mesh.cpp:632-822) + finalize (interaction.h:482-505) for the lanes'
    hit triangles.  mode: 'all' (Möller–Trumbore re-derives t, uv differentiably from ray and vertices, :676-686),
    'follow' (FollowShape: uv detached, p glued to the triangle), 'detach_shape' (vertices detached, :669-673)."""
    prim = prim.clamp(min=0)
    f = G.faces[prim]
    mesh = G.face_mesh[prim]
    P, N = (G.P.detach(), G.N.detach()) if mode == "detach_shape" else (G.P, G.N)
    p0, p1, p2 = P[f[:, 0]], P[f[:, 1]], P[f[:, 2]]
    u, v = b1, b2
    if mode in ("all", "detach_shape"):
        e1, e2 = p1 - p0, p2 - p0                                         # mesh.h:347-364
        pvec = torch.linalg.cross(ray_d, e2)
        inv_det = 1.0 / dot(e1, pvec)
        tvec = ray_o - p0
        u_d = dot(tvec, pvec) * inv_det
        qvec = torch.linalg.cross(tvec, e1)
        v_d = dot(ray_d, qvec) * inv_det
        u, v = replace_grad(b1, u_d), replace_grad(b2, v_d)
    b0 = 1.0 - u - v
    dp0, dp1 = p1 - p0, p2 - p0
    p = p0 * b0[:, None] + p1 * u[:, None] + p2 * v[:, None]
    n = normalize(torch.linalg.cross(dp0, dp1))
    dp_du, _ = coordinate_system(n)
    # meshes with texture coordinates: interpolated uv and the uv-aligned dp/du (mesh.cpp:731-763)
    t0, t1, t2 = G.UV[f[:, 0]], G.UV[f[:, 1]], G.UV[f[:, 2]]
    huv = G.has_uv[mesh]
    uv = torch.where(huv[:, None], t0 * b0[:, None] + t1 * u[:, None] + t2 * v[:, None], torch.stack([u, v], -1))
    duv0, duv1 = t1 - t0, t2 - t0
    det = duv0[:, 0] * duv1[:, 1] - duv0[:, 1] * duv1[:, 0]
    use = huv & (det != 0)
    inv_det = 1.0 / torch.where(use, det, torch.ones_like(det))
    dp_du = torch.where(use[:, None], (duv1[:, 1:2] * dp0 - duv0[:, 1:2] * dp1) * inv_det[:, None], dp_du)
    hn = G.has_normals[mesh]
    n0, n1, n2 = N[f[:, 0]], N[f[:, 1]], N[f[:, 2]]
    ns = n0 * b0[:, None] + n1 * u[:, None] + n2 * v[:, None]
    ns = torch.where(hn[:, None], ns, n)                                  # guards the normalisation of all-zero rows
    sh_n = torch.where(hn[:, None], normalize(ns), n)
    flip = G.flip[mesh][:, None]
    n = torch.where(flip, -n, n); sh_n = torch.where(flip, -sh_n, sh_n)
    sh_s = normalize(dp_du - sh_n * dot(sh_n, dp_du)[:, None])            # interaction.h:262-275
    sh_t = torch.linalg.cross(sh_n, sh_s)
    wi_w = -ray_d
    wi = torch.stack([dot(wi_w, sh_s), dot(wi_w, sh_t), dot(wi_w, sh_n)], -1)
    return dict(p=p, n=n, sh_n=sh_n, sh_s=sh_s, sh_t=sh_t, wi=wi, mesh=mesh, valid=valid, f=f, bw=torch.stack([b0, u, v], -1), uv=uv)


def to_local(si, v):
    return torch.stack([dot(v, si["sh_s"]), dot(v, si["sh_t"]), dot(v, si["sh_n"])], -1)


# ---- BSDF::eval (value·cosθo), differentiable in wi, wo -------------------------------------------------------------
def _fresnel_conductor(c, eta, k):                                        # include/mitsuba/render/fresnel.h:93-117
    c2 = c * c; s2 = 1.0 - c2; s4 = s2 * s2
    t1 = eta * eta - k * k - s2
    a2pb2 = torch.sqrt(torch.clamp(t1 * t1 + 4.0 * k * k * eta * eta, min=0.0))
    a = torch.sqrt(torch.clamp(0.5 * (a2pb2 + t1), min=0.0))
    term1 = a2pb2 + c2; term2 = 2.0 * c * a
    rs = (term1 - term2) / (term1 + term2)
    term3 = a2pb2 * c2 + s4; term4 = term2 * s2
    rp = rs * (term3 - term4) / (term3 + term4)
    return 0.5 * (rs + rp)


def _bsdf_eval_one(b, wi, wo, uv=None):
    from epsm_mitsuba3_b200 import capi
    zero = torch.zeros(wi.shape[0], 3, dtype=F64)
    if b.twosided:                                                        # twosided.cpp:226-232
        neg = wi[:, 2] < 0
        wo = torch.stack([wo[:, 0], wo[:, 1], torch.where(neg, -wo[:, 2], wo[:, 2])], -1)
        wi = torch.stack([wi[:, 0], wi[:, 1], wi[:, 2].abs()], -1)
    refl = torch.tensor(b.reflectance, dtype=F64)
    if getattr(b, "texture", None) is not None and uv is not None:
        # bitmap reflectance at the hit's uv, detached (exact for `nearest`, the filter of EPSM/exp/human.py; a bilinear lookup's
        # ∂ρ/∂uv term is not carried by the product either — stated in DESIGN §7)
        refl = torch.tensor(_o.texture_eval(b.texture, uv.detach().numpy().astype(np.float32)).astype(np.float64))
    else:
        refl = refl[None, :]
    if b.kind == capi.BSDF_DIFFUSE:                                       # diffuse.cpp:158-175
        act = (wi[:, 2] > 0) & (wo[:, 2] > 0)
        return torch.where(act[:, None], refl * (wo[:, 2:3] / math.pi), zero)
    if b.kind == capi.BSDF_ROUGHCONDUCTOR:                                # roughconductor.cpp:417-492, microfacet.h:183-210,365-392
        H = normalize(wo + wi)
        act = (wi[:, 2] > 0) & (wo[:, 2] > 0) & (dot(wi, H) > 0) & (dot(wo, H) > 0)
        au, av = float(b.alpha_u), float(b.alpha_v)
        ct2 = H[:, 2] ** 2
        if not b.ggx:
            D = torch.exp(-((H[:, 0] / au) ** 2 + (H[:, 1] / av) ** 2) / ct2) / (math.pi * au * av * ct2 * ct2)
        else:
            D = 1.0 / (math.pi * au * av * ((H[:, 0] / au) ** 2 + (H[:, 1] / av) ** 2 + ct2) ** 2)
        D = torch.where(D * H[:, 2] > 1e-20, D, torch.zeros_like(D))
        act = act & (D != 0)

        def g1(vv):
            xy = (au * vv[:, 0]) ** 2 + (av * vv[:, 1]) ** 2
            tan2 = xy / vv[:, 2] ** 2
            if not b.ggx:
                a = torch.rsqrt(tan2); a2 = a * a
                r = torch.where(a >= 1.6, torch.ones_like(a), (3.535 * a + 2.181 * a2) / (1.0 + 2.276 * a + 2.577 * a2))
            else:
                r = 2.0 / (1.0 + torch.sqrt(1.0 + tan2))
            r = torch.where(xy == 0, torch.ones_like(r), r)
            return torch.where(dot(vv, H) * vv[:, 2] <= 0, torch.zeros_like(r), r)
        val = D * g1(wi) * g1(wo) / (4.0 * wi[:, 2])
        c = dot(wi, H)
        Fr = torch.stack([_fresnel_conductor(c, float(b.eta[i]), float(b.k[i])) for i in range(3)], -1)
        return torch.where(act[:, None], Fr * refl * val[:, None], zero)
    return zero                                                           # delta BSDFs evaluate to zero


def bsdf_eval(G: Geometry, mesh, wi, wo, mask, uv=None):
    out = torch.zeros(wi.shape[0], 3, dtype=F64)
    bid = G.mesh_bsdf[mesh]
    for i, b in enumerate(G.bsdfs):
        sel = (bid == i) & mask
        if bool(sel.any()):
            idx = sel.nonzero()[:, 0]
            out = out.index_put((idx,), _bsdf_eval_one(b, wi[idx], wo[idx], None if uv is None else uv[idx]))
    return out


# ---- the pass ---------------------------------------------------------------------------------------------------------
def log_records(o_scene, sensor, seed, spp, max_depth, rr_depth, reparam_max_depth, n_rays, kappa, exponent, lane_begin, lane_end, n_iter):
    L = _o.lib()
    L.orc_prb_record_floats.restype = C.c_uint32
    L.orc_prb_record_floats.argtypes = [C.c_uint32, C.c_uint32]
    L.orc_prb_log.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                              C.c_float, C.c_float, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
    stride = L.orc_prb_record_floats(n_iter, n_rays)
    out = np.zeros((lane_end - lane_begin, stride), np.float32)
    L.orc_prb_log(o_scene.h, sensor, seed, spp, max_depth, rr_depth, min(reparam_max_depth, 0xffffffff), n_rays, kappa, exponent,
                  lane_begin, lane_end, n_iter, out.ctypes.data_as(C.c_void_p))
    return out


def render_backward(o_scene, sd, grad_in, sensor=0, seed=0, spp=4, max_depth=6, rr_depth=5, reparam_max_depth=None, n_rays=16,
                    kappa=1e5, exponent=3.0, lane_begin=0, lane_end=None, film_rgbw=None, terms=None):
    """Gradients of Σ image·grad_in w.r.t. (vertex positions, vertex normals), (V,3) float64 each.
    `film_rgbw`: the accumulated primal film of ALL lanes (H,W,4) when only a lane window is differentiated; by default it is
    accumulated from the lanes given.  `terms`: optional dict that receives per-lane scalars for debugging."""
    s = sd.sensors[sensor]
    reparam_max_depth = max_depth if reparam_max_depth is None else reparam_max_depth
    W, H = s.width, s.height
    border = 2 if s.sample_border else 0
    n_lanes = (W + 2 * border) * (H + 2 * border) * spp
    lane_end = n_lanes if lane_end is None else lane_end
    n_iter = min(max_depth, 8)
    rec = torch.tensor(log_records(o_scene, sensor, seed, spp, max_depth, rr_depth, reparam_max_depth, n_rays, kappa, exponent,
                                   lane_begin, lane_end, n_iter), dtype=F64)
    G = Geometry(sd)
    Nl = rec.shape[0]
    cam_o, cam_d, pos0 = rec[:, 0:3], rec[:, 3:6], rec[:, 6:8]
    Ltot = rec[:, 8:11]
    cam_aux = rec[:, HDR:HDR + AUX * n_rays].reshape(Nl, n_rays, AUX)
    cam_active = rec[:, 13] > 0.5

    # -- sample_rays with the reparameterisation (common.py:372-397): the film position follows the warp field
    d_r, det_cam = reparam(G, cam_o, cam_d, cam_aux, cam_active)
    to_world = torch.tensor(s.to_world.matrix, dtype=F64)
    inv_tw = torch.linalg.inv(to_world)
    c2s = torch.linalg.inv(_sample_to_camera(s))
    it_p = cam_o + d_r
    ref_p = it_p @ inv_tw[:3, :3].T + inv_tw[:3, 3]
    hp = ref_p @ c2s[:3, :3].T + c2s[:3, 3]
    hw = ref_p @ c2s[3, :3] + c2s[3, 3]
    uv = hp[:, :2] / hw[:, None]
    pos = replace_grad(pos0, uv * torch.tensor([W, H], dtype=F64))          # perspective.cpp:282-320 (ds.uv · resolution)

    # -- film: block.put(pos, L·det, weight=det) → develop (common.py:919-945; imageblock.cpp:414-494; hdrfilm.cpp:357-383)
    L_leaf = Ltot.clone().requires_grad_(True)
    pix = torch.floor(pos0).long() - 2
    rel = pix.to(F64) + 0.5 - pos
    film = torch.zeros(H * W, 4, dtype=F64)
    vals = torch.cat([L_leaf * det_cam[:, None], det_cam[:, None]], 1)
    wx = [filter_eval(rel[:, 0] + i) for i in range(5)]
    wy = [filter_eval(rel[:, 1] + i) for i in range(5)]
    contribs = []
    for ys in range(5):
        for xs in range(5):
            x, y = pix[:, 0] + xs, pix[:, 1] + ys
            ok = (x >= 0) & (y >= 0) & (x < W) & (y < H)
            idx = (y.clamp(0, H - 1) * W + x.clamp(0, W - 1))
            contribs.append((idx, ok, (wy[ys] * wx[xs])))
    if film_rgbw is None:
        for idx, ok, wgt in contribs:
            film = film.index_add(0, idx, vals * (wgt * ok)[:, None])
        film_total = film
    else:
        # lanes outside the window are constants of the differentiation: add their (detached) film
        own = torch.zeros(H * W, 4, dtype=F64)
        for idx, ok, wgt in contribs:
            own = own.index_add(0, idx, vals * (wgt * ok)[:, None])
        film_total = torch.tensor(np.asarray(film_rgbw, np.float64).reshape(H * W, 4)) - own.detach() + own
    wch = film_total[:, 3:4]
    image = film_total[:, :3] / torch.where(wch == 0, torch.ones_like(wch), wch)
    gin = torch.tensor(np.asarray(grad_in, np.float64)[..., :3].reshape(H * W, 3))
    S0 = (image * gin).sum()
    S0.backward()
    dL = L_leaf.grad.clone()                                               # δL = dr.grad(L), common.py:944
    if terms is not None:
        terms["dL"] = dL.numpy().copy(); terms["film_grad_pos"] = G.P.grad.numpy().copy()

    # -- sample(mode=Backward) (prb_reparam.py:329-602): one self-contained AD graph per iteration
    base = HDR + AUX * n_rays
    stride = IT + 2 * AUX * n_rays
    n_it = rec[:, 12].long()
    prev = None                                                            # detached record of the previous vertex
    for it in range(n_iter):
        I = rec[:, base + it * stride: base + (it + 1) * stride]
        lane_on = n_it > it
        if not bool(lane_on.any()):
            break
        valid = (I[:, 0] > 0.5) & lane_on
        prim, b1, b2 = I[:, 1].long(), I[:, 2], I[:, 3]
        ray_o, ray_d = I[:, 5:8], I[:, 8:11]
        beta, mis_em = I[:, 11:14], I[:, 14]
        active_next, active_em = (I[:, 16] > 0.5) & lane_on, (I[:, 17] > 0.5) & lane_on
        ds_p, em_weight, mis_direct = I[:, 18:21], I[:, 25:28], I[:, 28]
        bs_pdf, bs_weight, ray_next_d = I[:, 32], I[:, 33:36], I[:, 37:40]
        L_prev, L_cur = I[:, 40:43], I[:, 43:46]
        nvalid, nprim, nb1, nb2 = (I[:, 46] > 0.5) & lane_on, I[:, 47].long(), I[:, 48], I[:, 49]
        L_next, bsn_wo = I[:, 50:53], I[:, 53:56]
        main_act, em_act = (I[:, 57] > 0.5) & lane_on, (I[:, 58] > 0.5) & lane_on
        em_o0, em_d0 = I[:, 59:62], I[:, 62:65]
        Le, depth = I[:, 65:68], I[:, 71]
        aux_main = I[:, IT:IT + AUX * n_rays].reshape(Nl, n_rays, AUX)
        aux_em = I[:, IT + AUX * n_rays:].reshape(Nl, n_rays, AUX)
        first = torch.full((Nl,), it == 0)

        # reparameterise the ray that arrives at this vertex; its origin follows the previous shape (prb_reparam.py:352-366)
        if prev is None:
            o_in = ray_o
        else:
            si_prev_f = surface_interaction(G, prev["prim"], prev["b1"], prev["b2"], prev["valid"], prev["ray_o"], prev["ray_d"], "follow")
            o_in = replace_grad(ray_o, si_prev_f["p"])                     # spawn_ray: p + detached offset (interaction.h:159-165)
        d_rp, det_rp = reparam(G, o_in, ray_d, aux_main, main_act)
        if it == 0:
            det_rp = torch.ones_like(det_rp)                               # ray_reparam_det[first_vertex] = 1

        # si_cur with derivative tracking (prb_reparam.py:373-375)
        si = surface_interaction(G, prim, b1, b2, valid, ray_o, d_rp, "all")
        # emitter sampling is detached; the shadow ray is reparameterised from a point that follows the shape (:410-424)
        si_f = surface_interaction(G, prim, b1, b2, valid, ray_o, ray_d, "follow")
        o_e = replace_grad(em_o0, si_f["p"])
        d_e_in = ds_p - o_e
        d_e_in = torch.where(active_em[:, None], d_e_in, em_d0)
        d_e_in = replace_grad(em_d0, normalize(d_e_in))
        d_e, det_e = reparam(G, o_e, d_e_in, aux_em, em_act)
        wo_em = to_local(si, d_e)
        bsdf_em = bsdf_eval(G, si["mesh"], si["wi"], wo_em, active_em, si["uv"])
        Lr_dir = beta * mis_direct[:, None] * bsdf_em * em_weight * det_e[:, None]          # :431-435 (em_weight: constant radiance)
        Lr_dir = torch.where(active_em[:, None], Lr_dir, torch.zeros_like(Lr_dir))

        # indirect term: L · bsdf_val / detach(bsdf_val) (:553-571)
        wo_next = to_local(si, ray_next_d)
        bsdf_val = bsdf_eval(G, si["mesh"], si["wi"], wo_next, active_next, si["uv"])
        bv_det = bs_weight * bs_pdf[:, None]
        inv_bv = torch.where(bv_det != 0, 1.0 / bv_det, torch.zeros_like(bv_det))
        ratio = inv_bv * bsdf_val
        Lr_ind = L_cur * (ratio - ratio.detach() + 1.0)

        # adjacent vertices, derivatives only through the reparameterised ray (:522-551)
        si_ro = surface_interaction(G, prim, b1, b2, valid, ray_o, d_rp, "detach_shape")
        extra = torch.zeros(Nl, 3, dtype=F64)
        if prev is not None:
            si_p = prev["si"]
            wo_prev = normalize(torch.where(valid[:, None], si_ro["p"] - si_p["p"], torch.ones_like(si_ro["p"])))
            bv_prev = bsdf_eval(G, si_p["mesh"], si_p["wi"], to_local(si_p, wo_prev), prev["valid"] & valid, si_p["uv"])
            extra = extra + torch.where((prev["valid"] & valid)[:, None], L_prev * bv_prev / bv_prev.detach().clamp(min=1e-8), torch.zeros_like(extra))
        si_n = surface_interaction(G, nprim, nb1, nb2, nvalid, ray_o, ray_next_d, "follow")
        si_n = {k: (v.detach() if torch.is_tensor(v) and v.dtype == F64 else v) for k, v in si_n.items()}
        wi_next = normalize(torch.where((valid & nvalid)[:, None], si_ro["p"] - si_n["p"], torch.ones_like(si_ro["p"])))
        bv_next = bsdf_eval(G, si_n["mesh"], to_local(si_n, wi_next), bsn_wo, valid & nvalid, si_n["uv"])
        extra = extra + torch.where((valid & nvalid)[:, None], L_next * bv_next / bv_next.detach().clamp(min=1e-8), torch.zeros_like(extra))

        Lo = (Le + Lr_dir + Lr_ind) * det_rp[:, None] + extra               # :576
        Lo = torch.where(lane_on[:, None], Lo, torch.zeros_like(Lo))
        S = (dL * Lo).sum()
        if S.requires_grad:
            S.backward()
        if terms is not None:
            terms[f"Lo{it}"] = Lo.detach().numpy().copy()
        # detached record for the next iteration (dr.disable_grad(si_prev), :366)
        si_d = {k: (v.detach() if torch.is_tensor(v) and v.dtype == F64 else v) for k, v in
                surface_interaction(G, prim, b1, b2, valid, ray_o, ray_d, "follow").items()}
        prev = dict(prim=prim, b1=b1, b2=b2, valid=valid, ray_o=ray_o, ray_d=ray_d, si=si_d)
    gp = G.P.grad if G.P.grad is not None else torch.zeros_like(G.P)
    gn = G.N.grad if G.N.grad is not None else torch.zeros_like(G.N)
    return gp.numpy().copy(), gn.numpy().copy()


def reparam(G, ray_o, ray_d, aux, active):
    """reparameterize_ray (reparam.py:83-118, 236-330).  Returns (direction, det) with the values (ray.d, 1) and the
    gradients of normalize(ray.d + V/Z), (div − V/Z·dZ)/Z evaluated at V = 0 (backward_symbolic, reparam.py:276-287), where
    V = Σ w_i·V_direct_i and div = Σ dot(d_w_omega_i, V_direct_i) are attached through the hit points (FollowShape) and
    through the ray: its origin, and its direction on auxiliary rays that miss (V_direct = ray.d, reparam.py:101)."""
    n = aux.shape[1]
    hit = aux[:, :, 0] > 0.5
    prim = aux[:, :, 1].long()
    b1, b2 = aux[:, :, 2], aux[:, :, 3]
    w, dw = aux[:, :, 4], aux[:, :, 5:8]
    f = G.faces[prim.clamp(min=0)]
    p = (G.P[f[..., 0]] * (1 - b1 - b2)[..., None] + G.P[f[..., 1]] * b1[..., None] + G.P[f[..., 2]] * b2[..., None])
    dexp = ray_d[:, None, :].expand(-1, n, -1)
    rel = torch.where(hit[..., None], p - ray_o[:, None, :], dexp.detach())
    V_direct = torch.where(hit[..., None], normalize(rel), dexp)
    Z = w.sum(1).clamp(min=1e-8)
    dZ = dw.sum(1)
    V = (w[..., None] * V_direct).sum(1)
    div = (dw * V_direct).sum(-1).sum(1)
    V0 = (V - V.detach()) / Z[:, None]
    direction = normalize(ray_d.detach() + V0)
    det = 1.0 + ((div - div.detach()) - dot(V0, dZ)) / Z
    direction = torch.where(active[:, None], direction, ray_d.detach())
    det = torch.where(active, det, torch.ones_like(det))
    return direction, det


def _sample_to_camera(s):
    """ProjectiveCamera: sample_to_camera = (perspective · translate · scale)⁻¹ (include/mitsuba/render/sensor.h:227-262,
    src/sensors/perspective.cpp:173-186), as a float64 4×4."""
    aspect = s.width / s.height
    n, f = s.near_clip, s.far_clip
    recip = 1.0 / (f - n); cot = 1.0 / math.tan(0.5 * math.radians(s.fov))
    persp = np.array([[cot, 0, 0, 0], [0, cot, 0, 0], [0, 0, f * recip, -n * f * recip], [0, 0, 1, 0]], np.float64)
    tr = np.eye(4); tr[0, 3] = -1.0; tr[1, 3] = -1.0 / aspect
    sc = np.diag([-0.5, -0.5 * aspect, 1.0, 1.0])
    c2s = sc @ tr @ persp
    return torch.tensor(np.linalg.inv(c2s), dtype=F64)
