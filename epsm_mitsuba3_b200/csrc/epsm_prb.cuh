// epsm_prb.cuh — per-lane code of `prb_reparam`'s reverse-mode pass (hybrid phase 2 of the EPSM drivers, optim.py:89-92,113-119).
//
// Reference (paths relative to the reference root):
//   RBIntegrator.render_backward          src/python/python/ad/integrators/common.py:803-975
//   sample_rays with reparameterisation   common.py:291-397   (film position follows the warp field; _ReparamWrapper :1154-1217)
//   PRBReparamIntegrator.sample           src/python/python/ad/integrators/prb_reparam.py:277-607
//   reparameterize_ray / warp field       src/python/python/ad/reparam.py:10-118 (one auxiliary ray), :236-330 (backward_symbolic)
//   boundary tests                        src/render/mesh.cpp:823-886, src/shapes/rectangle.cpp:319-320
//   von Mises–Fisher warp                 include/mitsuba/core/warp.h:546-583
//   differentiable hit record             src/render/mesh.cpp:632-822 (DetachShape / FollowShape), include/mitsuba/render/mesh.h:347-364
//
// The reference obtains the derivatives with Dr.Jit's reverse-mode AD, one self-contained graph per path vertex.  Here the
// same terms are differentiated by hand: every block below names the reference expression it is the adjoint of.  The
// detached half (random walk, sampling, auxiliary rays and their harmonic weights) is the primal code of epsm_device.cuh.
// BSDF values are differentiated with 6-wide forward duals (wi, wo) — exact, and independent of the BSDF's formula.
// Host + device, like epsm_device.cuh (tests/hostsim runs it on the CPU against oracle/prb_reparam_ref.py).
#pragma once
#include "epsm_device.cuh"


// ------------------------------------------------------------------------------------------------
//  Forward duals over (wi.x, wi.y, wi.z, wo.x, wo.y, wo.z)
// ------------------------------------------------------------------------------------------------
struct D6 { float v, d[6]; };
EPSM_HD D6 d6c(float v) { D6 r; r.v = v; for (int i = 0; i < 6; ++i) r.d[i] = 0.f; return r; }
EPSM_HD D6 d6v(float v, int k) { D6 r = d6c(v); r.d[k] = 1.f; return r; }
EPSM_HD D6 operator+(D6 a, D6 b) { D6 r; r.v = a.v + b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
EPSM_HD D6 operator-(D6 a, D6 b) { D6 r; r.v = a.v - b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
EPSM_HD D6 operator-(D6 a) { D6 r; r.v = -a.v; for (int i = 0; i < 6; ++i) r.d[i] = -a.d[i]; return r; }
EPSM_HD D6 operator*(D6 a, D6 b) { D6 r; r.v = a.v * b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
EPSM_HD D6 operator*(D6 a, float s) { D6 r; r.v = a.v * s; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * s; return r; }
EPSM_HD D6 operator+(D6 a, float s) { a.v += s; return a; }
EPSM_HD D6 operator/(D6 a, D6 b) {
    D6 r; const float ib = 1.0f / b.v; r.v = a.v * ib;
    for (int i = 0; i < 6; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
    return r;
}
EPSM_HD D6 d6_sqrt(D6 a) { D6 r; r.v = sqrtf(a.v); const float k = r.v > 0.f ? 0.5f / r.v : 0.f; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * k; return r; }
EPSM_HD D6 d6_exp(D6 a) { D6 r; r.v = expf(a.v); for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * r.v; return r; }
EPSM_HD D6 d6_safe_sqrt(D6 a) { if (!(a.v > 0.f)) return d6c(0.f); return d6_sqrt(a); }
struct D6v3 { D6 x, y, z; };
EPSM_HD D6 d6_dot(const D6v3 &a, const D6v3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// fresnel.h:93-117 (fresnel_cond in epsm_device.cuh), dual argument
EPSM_HD D6 d6_fresnel_cond(D6 c, float eta_r, float eta_i) {
    D6 c2 = c * c, s2 = d6c(1.f) - c2, s4 = s2 * s2;
    D6 t1 = d6c(eta_r * eta_r - eta_i * eta_i) - s2;
    D6 a2pb2 = d6_safe_sqrt(t1 * t1 + 4.f * eta_i * eta_i * eta_r * eta_r);
    D6 a = d6_safe_sqrt((a2pb2 + t1) * 0.5f);
    D6 term1 = a2pb2 + c2, term2 = c * a * 2.f;
    D6 rs = (term1 - term2) / (term1 + term2);
    D6 term3 = a2pb2 * c2 + s4, term4 = term2 * s2;
    D6 rp = rs * (term3 - term4) / (term3 + term4);
    return (rs + rp) * 0.5f;
}
// microfacet.h:365-392 (mf_g1), dual arguments
EPSM_HD D6 d6_g1(const DBsdf &b, const D6v3 &v, const D6v3 &m) {
    const bool ggx = (b.flags & EPSM_BSDF_GGX) != 0;
    D6 ax = v.x * b.alpha_u, ay = v.y * b.alpha_v;
    D6 xy = ax * ax + ay * ay, tan2 = xy / (v.z * v.z), r;
    if (!ggx) {
        D6 a = d6c(1.f) / d6_sqrt(tan2), a2 = a * a;
        r = a.v >= 1.6f ? d6c(1.f) : (a * 3.535f + a2 * 2.181f) / (d6c(1.f) + a * 2.276f + a2 * 2.577f);
    } else {
        r = d6c(2.f) / (d6c(1.f) + d6_sqrt(d6c(1.f) + tan2));
    }
    if (xy.v == 0.f) r = d6c(1.f);
    if (d6_dot(v, m).v * v.z.v <= 0.f) r = d6c(0.f);
    return r;
}
// BSDF::eval(si, wo) = value·cosθo as in bsdf_eval_pdf (epsm_device.cuh), with duals.  Branches follow the primal values.
EPSM_HDN void bsdf_eval_d6(const DBsdf &b, D6v3 wi, D6v3 wo, D6 out[3]) {
    out[0] = out[1] = out[2] = d6c(0.f);
    if (b.flags & EPSM_BSDF_TWOSIDED) { if (f2u(wi.z.v) & 0x80000000u) { wo.z = -wo.z; wi.z = -wi.z; } }          // twosided.cpp:226-232
    if (b.kind == EPSM_BSDF_DIFFUSE) {                                                                              // diffuse.cpp:158-175
        if (wi.z.v > 0.f && wo.z.v > 0.f) for (int c = 0; c < 3; ++c) out[c] = wo.z * (b.reflectance[c] * EPSM_INV_PI);
    } else if (b.kind == EPSM_BSDF_ROUGHCONDUCTOR) {                                                                // roughconductor.cpp:417-492
        D6v3 Hs{wo.x + wi.x, wo.y + wi.y, wo.z + wi.z};
        D6 il = d6c(1.f) / d6_sqrt(d6_dot(Hs, Hs));
        D6v3 H{Hs.x * il, Hs.y * il, Hs.z * il};
        D6 wiH = d6_dot(wi, H), woH = d6_dot(wo, H);
        if (!(wi.z.v > 0.f && wo.z.v > 0.f && wiH.v > 0.f && woH.v > 0.f)) return;
        const float au = b.alpha_u, av = b.alpha_v;
        D6 ct2 = H.z * H.z, hx = H.x * (1.f / au), hy = H.y * (1.f / av), D;
        if (!(b.flags & EPSM_BSDF_GGX)) D = d6_exp(-((hx * hx + hy * hy) / ct2)) / (ct2 * ct2 * (EPSM_PI * au * av));      // microfacet.h:183-210
        else { D6 q = hx * hx + hy * hy + ct2; D = d6c(1.f) / (q * q * (EPSM_PI * au * av)); }
        if (!(D.v * H.z.v > 1e-20f)) return;
        D6 val = D * d6_g1(b, wi, H) * d6_g1(b, wo, H) / (wi.z * 4.f);
        for (int c = 0; c < 3; ++c) out[c] = d6_fresnel_cond(wiH, b.eta[c], b.k[c]) * val * b.reflectance[c];
    }
}
// Σ_c a_c·f_c(wi, wo): value and gradient w.r.t. the two local directions.
EPSM_HDN void bsdf_grad(const DBsdf &b, f3 wi, f3 wo, f3 a, f3 &g_wi, f3 &g_wo, f3 &value) {
    D6v3 dwi{d6v(wi.x, 0), d6v(wi.y, 1), d6v(wi.z, 2)}, dwo{d6v(wo.x, 3), d6v(wo.y, 4), d6v(wo.z, 5)};
    D6 f[3]; bsdf_eval_d6(b, dwi, dwo, f);
    value = mk3(f[0].v, f[1].v, f[2].v);
    float g[6];
    for (int k = 0; k < 6; ++k) g[k] = a.x * f[0].d[k] + a.y * f[1].d[k] + a.z * f[2].d[k];
    g_wi = mk3(g[0], g[1], g[2]); g_wo = mk3(g[3], g[4], g[5]);
}

// ------------------------------------------------------------------------------------------------
//  Vector adjoints
// ------------------------------------------------------------------------------------------------
// y = x/|x|: g_x = (g − ŷ(ŷ·g))/|x|
EPSM_HD f3 normalize_bwd(f3 x, f3 g) { float il = rsqrt_e(sqn3(x)); f3 y = x * il; return (g - y * dot3(y, g)) * il; }
// coordinate_system(n).s (vector.h:116-136): g_n = (∂s/∂n)ᵀ g_s
EPSM_HD f3 onb_s_bwd(f3 n, f3 g) {
    const float sign = (f2u(n.z) & 0x80000000u) ? -1.0f : 1.0f, a = -rcp_e(sign + n.z), a2 = a * a;
    // s = (sign·nx²·a + 1, sign·nx·ny·a, −sign·nx);  ∂a/∂nz = a²
    const float gx = g.x * (sign * 2.f * n.x * a) + g.y * (sign * n.y * a) - g.z * sign;
    const float gy = g.y * (sign * n.x * a);
    const float gz = g.x * (sign * n.x * n.x * a2) + g.y * (sign * n.x * n.y * a2);
    return mk3(gx, gy, gz);
}
// Möller–Trumbore (mesh.h:347-364): adjoint of (u, v) w.r.t. the ray direction and the triangle.  o is detached.
struct MtGrad { f3 d, p0, p1, p2; };
EPSM_HD void mt_uv_bwd(f3 o, f3 d, f3 p0, f3 p1, f3 p2, float g_u, float g_v, MtGrad &G) {
    const f3 e1 = p1 - p0, e2 = p2 - p0, pvec = cross3(d, e2), tvec = o - p0, qvec = cross3(tvec, e1);
    const float det = dot3(e1, pvec), inv = 1.0f / det;
    const float g_inv = g_u * dot3(tvec, pvec) + g_v * dot3(d, qvec), g_det = -inv * inv * g_inv;
    f3 g_tvec = pvec * (g_u * inv), g_pvec = tvec * (g_u * inv) + e1 * g_det, g_e1 = pvec * g_det;
    const f3 g_qvec = d * (g_v * inv);
    f3 g_d = qvec * (g_v * inv);
    g_tvec = g_tvec + cross3(e1, g_qvec); g_e1 = g_e1 + cross3(g_qvec, tvec);          // qvec = tvec × e1
    g_d = g_d + cross3(e2, g_pvec); const f3 g_e2 = cross3(g_pvec, d);                  // pvec = d × e2
    G.d = g_d; G.p1 = g_e1; G.p2 = g_e2; G.p0 = -(g_tvec + g_e1 + g_e2);
}

// ------------------------------------------------------------------------------------------------
//  Reparameterisation (reparam.py)
// ------------------------------------------------------------------------------------------------
// e2k = expf(−2κ): the same for every auxiliary ray of a call, so the callers evaluate it once
EPSM_HD f3 square_to_von_mises_fisher(float sx, float sy_, float kappa, float e2k) {      // warp.h:546-583, approach 1.2
    const float sy = fmaxf(1.f - sy_, 1e-6f);
    const float cos_theta = 1.f + logf(fmaf(1.f - sy, e2k, sy)) / kappa;
    float s, c; sincos_dr(6.283185307179586f * sx, s, c);
    const float sin_theta = safe_sqrt(1.f - cos_theta * cos_theta);
    return mk3(c * sin_theta, s * sin_theta, cos_theta);
}
// si.boundary_test of the hit (RayFlags::BoundaryTest): mesh.cpp:823-886; rectangle.cpp:319-320 for tessellated rectangles
EPSM_HDN float boundary_test(const DScene &sc, const Hit &h, f3 ray_d) {
    const uint32_t mid = sc.face_mesh[h.prim]; const DMesh m = sc.meshes[mid];
    const uint32_t f0 = sc.faces[3 * h.prim], f1 = sc.faces[3 * h.prim + 1], f2 = sc.faces[3 * h.prim + 2];
    const f3 p0 = ld3(sc.pos, f0), p1 = ld3(sc.pos, f1), p2 = ld3(sc.pos, f2);
    const float b1 = h.u, b2 = h.v, b0 = 1.f - b1 - b2;
    const f3 p = fma3(p0, b0, fma3(p1, b1, p2 * b2));
    if (m.flags & EPSM_MESH_RECTANGLE) {
        const f3 c0 = ld3(sc.pos, m.v_off), ex = ld3(sc.pos, m.v_off + 1) - c0, ey = ld3(sc.pos, m.v_off + 3) - c0, r = p - c0;
        const float u = dot3(r, ex) / dot3(ex, ex), v = dot3(r, ey) / dot3(ey, ey);
        return fminf(0.5f - fabsf(u - 0.5f), 0.5f - fabsf(v - 0.5f));
    }
    const f3 dp0 = p1 - p0, dp1 = p2 - p0, rel = p - p0;
    const float bb1 = dot3(dp0, rel), bb2 = dot3(dp1, rel), a11 = dot3(dp0, dp0), a12 = dot3(dp0, dp1), a22 = dot3(dp1, dp1);
    const float inv_det = rcp_e(a11 * a22 - a12 * a12);
    const float u = fmaf(a22, bb1, -(a12 * bb2)) * inv_det, v = fmaf(-a12, bb1, a11 * bb2) * inv_det, w = 1.f - u - v;
    if (!m.has_normals) {                                                           // 2-D signed distance to an equilateral triangle
        const float hh = 0.5f * sqrtf(3.f);
        const float px = 0.f * w + 1.f * u + 0.5f * v, py = 0.f * w + 0.f * u + hh * v;
        const float tx[3] = {0.f, 1.f, 0.5f}, ty[3] = {0.f, 0.f, hh};
        const float ex[3] = {tx[1] - tx[0], tx[2] - tx[1], tx[0] - tx[2]}, ey[3] = {ty[1] - ty[0], ty[2] - ty[1], ty[0] - ty[2]};
        const float s = (ex[0] * ey[2] - ey[0] * ex[2]) < 0.f ? -1.f : 1.f;
        float dx = INFINITY, dy = INFINITY;
        for (int k = 0; k < 3; ++k) {
            const float vx = px - tx[k], vy = py - ty[k];
            const float tt = fminf(fmaxf((vx * ex[k] + vy * ey[k]) / (ex[k] * ex[k] + ey[k] * ey[k]), 0.f), 1.f);
            const float qx = vx - ex[k] * tt, qy = vy - ey[k] * tt;
            dx = fminf(dx, qx * qx + qy * qy); dy = fminf(dy, s * (vx * ey[k] - vy * ex[k]));
        }
        (void) dy;
        return sqrtf(dx) / (sqrtf(3.f) / 6.f);
    }
    const f3 n = fma3(ld3(sc.nrm, f0), w, fma3(ld3(sc.nrm, f1), u, ld3(sc.nrm, f2) * v));
    const float dp = dot3(n, -ray_d);
    return dp * dp;
}
struct AuxRec { uint32_t prim; float u, v, w; f3 dw; };
// One auxiliary ray of _sample_warp_field (reparam.py:83-118): its direction, and — from its closest hit — the detached
// quantities Z_i = w and dZ_i = d_w_omega.
// (fs, ft) = coordinate_system(d), the frame of the reparameterised ray: also the same for all its auxiliary rays
struct AuxFrame { f3 fs, ft; float e2k; };
EPSM_HD AuxFrame aux_frame(f3 d, float kappa) { AuxFrame a; coordinate_system(d, a.fs, a.ft); a.e2k = expf(-2.f * kappa); return a; }   // Frame3f(ray.d)
EPSM_HD f3 warp_field_direction(f3 d, const AuxFrame &fr, float sx, float sy, float kappa, f3 &tang) {
    const f3 om = square_to_von_mises_fisher(sx, sy, kappa, fr.e2k);
    const f3 fs = fr.fs, ft = fr.ft;
    tang = fma3(d, 0.f, fma3(ft, om.y, fs * om.x));                                // to_world((ω.x, ω.y, 0))
    return fma3(d, om.z, fma3(ft, om.y, fs * om.x));
}
EPSM_HDN void warp_field_weights(const DScene &sc, const Hit &h, f3 aux_d, f3 tang, float sy, float kappa, float e2k, float exponent, AuxRec &r) {
    const bool hit = h.prim != 0xffffffffu;
    const float B = hit ? boundary_test(sc, h, aux_d) : 1.0f;
    const float inv_vmf = rcp_e(fmaf(sy, e2k, 1.f - sy));
    const float w_denom = inv_vmf - 1.f + B;
    const float w_denom_rcp = w_denom > 1e-4f ? rcp_e(w_denom) : 0.f;
    // dr.power(x, exponent): the reference's default exponent 3 as two multiplications, anything else through powf
    const float w = (exponent == 3.f ? w_denom_rcp * w_denom_rcp * w_denom_rcp : powf(w_denom_rcp, exponent)) * inv_vmf;
    float tmp1 = inv_vmf * w * w_denom_rcp * kappa * exponent;
    tmp1 = fminf(fmaxf(tmp1, -1e10f), 1e10f);
    r.prim = h.prim; r.u = h.u; r.v = h.v; r.w = w; r.dw = tang * tmp1;
}

// ------------------------------------------------------------------------------------------------
//  Where the rays of a lane are traced.  MegaIo: inline, one thread walks the whole lane (k_prb_backward, host harness).
//  The replay-wavefront policy of epsm_kernels.cu (WfIo) has the same interface: in pass p it pushes the rays of round p
//  into the queues of the persistent traversal kernel and reads the results of earlier rounds from per-lane slots.
//  Rounds of a lane: 0 = primary ray + the camera ray's auxiliary rays; 1 + 2·it = NEE shadow ray and next path ray of
//  iteration it; 2 + 2·it = the run-ahead NEE shadow ray and the auxiliary rays of the emitter ray and of the main ray.
// ------------------------------------------------------------------------------------------------
template <bool COUNT>
struct MegaIo {
    static constexpr bool wavefront = false;
    Tracer &tr;
    EPSM_HD explicit MegaIo(Tracer &t) : tr(t) {}
    EPSM_HD Hit closest(int, uint32_t, const Ray &r) { return trace_closest<COUNT>(tr, r); }
    EPSM_HD bool occluded(int, uint32_t, const Ray &r) { return trace_any<COUNT>(tr, r); }
    EPSM_HD bool stop_after(int) const { return false; }           // pass == round: the round's rays were just pushed
    EPSM_HD bool adjoint_now(int) const { return true; }           // the adjoint that consumes round r runs in pass r + 1
    EPSM_HD bool aux_push_only(int) const { return false; }
    EPSM_HD bool aux_skip(int) const { return false; }
    EPSM_HD uint32_t aux_reserve(uint32_t) { return 0u; }
    EPSM_HD void aux_push(uint32_t, uint32_t, const Ray &) {}
    EPSM_HD Hit aux_hit(uint32_t, const Ray &r) { return trace_closest<COUNT>(tr, r); }
    EPSM_HD bool lane_radiance(f3 &) const { return false; }       // false: replay the primal path inline
    EPSM_HD void finished() {}
    template <typename C> EPSM_HD bool resume(C &) const { return false; }             // no checkpoints: one thread walks the whole lane
    template <typename C> EPSM_HD void checkpoint(uint32_t, const C &) const {}
};

// Where a ray origin is glued to (FollowShape: p = Σ bw_k·p_k with detached barycentrics); prim = 0xffffffff: detached origin.
struct Anchor { uint32_t prim; float bw[3]; };
template <typename Sink>
EPSM_HD void scatter_tri(const DScene &sc, Sink &sink, uint32_t prim, float b0, float b1, float b2, f3 g) {
    if (prim == 0xffffffffu || !(sc.meshes[sc.face_mesh[prim]].flags & EPSM_MESH_GRAD)) return;
    sink.pos(sc.faces[3 * prim], g * b0); sink.pos(sc.faces[3 * prim + 1], g * b1); sink.pos(sc.faces[3 * prim + 2], g * b2);
}
template <typename Sink>
EPSM_HD void scatter_tri3(const DScene &sc, Sink &sink, uint32_t prim, f3 g0, f3 g1, f3 g2) {      // per-vertex gradients of one triangle
    if (prim == 0xffffffffu || !(sc.meshes[sc.face_mesh[prim]].flags & EPSM_MESH_GRAD)) return;
    sink.pos(sc.faces[3 * prim], g0); sink.pos(sc.faces[3 * prim + 1], g1); sink.pos(sc.faces[3 * prim + 2], g2);
}
// Backward of one reparameterize_ray call (backward_symbolic, reparam.py:236-330).  The op's outputs have the values
// (ray.d, 1); (g_dir, g_det) are the gradients that arrived at them.  Traces the n auxiliary rays once (the reference
// walks them twice with the same random numbers: weight normalisation, then back-propagation), forms
//   g_V = ((I − d̂d̂ᵀ)g_dir/|d| − g_det·dZ/Z)/Z,  g_div = g_det/Z       [direction = normalize(d + V/Z), divergence = (div − V/Z·dZ)/Z at V = 0]
// and pushes w_i·g_V + g_div·dZ_i through V_direct_i = normalize(p_i − o) (hit; p_i follows its triangle) or = d (miss).
// The origin's gradient goes to `anchor`; with `d_from`: d = normalize(d_from − o) is attached too (the emitter ray).
// reparam_backward has three call sites in the lane function; kept out of line the reverse-pass kernels are smaller (218 instead of
// 230 registers) and 10 % faster (293.7 → 265.5 ms at 512² × 8 spp, profiles/r02_summary.md §N)
#ifndef EPSM_PRB_REPARAM_INLINE
#if defined(__CUDACC__)
#define EPSM_PRB_REPARAM_INLINE __host__ __device__ __noinline__
#else
#define EPSM_PRB_REPARAM_INLINE
#endif
#endif
template <typename Io, typename Sink>
EPSM_PRB_REPARAM_INLINE void reparam_backward(const DScene &sc, Io &io, Sink &sink, Pcg &aux_rng, const PrbConfig &cfg, int round, uint32_t aux_base, bool active,
                               f3 o, f3 d, f3 g_dir, float g_det, const Anchor &anchor, bool d_attached, f3 d_from) {
    if (!active) return;                                                           // masked lanes do not advance the generator
    const uint32_t n = cfg.n_rays;
    if (io.aux_skip(round)) { for (uint32_t i = 0; i < 2 * n; ++i) pcg_next(aux_rng); return; }        // replay of a round long consumed
    if (io.aux_push_only(round)) {                                                 // wavefront, pass == round: queue the auxiliary rays
        // the n rays of one reparameterised ray go into n CONSECUTIVE queue entries: they share an origin and lie in a cone of
        // κ = 1e5 around d, so a warp of the traversal kernel that picks them up together walks one path through the BVH
        // (one L1 wavefront per node instead of one per thread, no divergence)
        const uint32_t q0 = io.aux_reserve(n);
        const AuxFrame fr = aux_frame(d, cfg.kappa);
        for (uint32_t i = 0; i < n; ++i) {
            const float sx = pcg_f32(aux_rng), sy = pcg_f32(aux_rng);
            f3 tang; const f3 ad = warp_field_direction(d, fr, sx, sy, cfg.kappa, tang);
            (void) sy; io.aux_push(aux_base + i, q0 + i, Ray{o, ad, FLT_MAX});
        }
        return;
    }
    const bool live = g_det != 0.f || g_dir.x != 0.f || g_dir.y != 0.f || g_dir.z != 0.f;
    if (!live) { for (uint32_t i = 0; i < 2 * n; ++i) pcg_next(aux_rng); return; }
    AuxRec rec[EPSM_PRB_MAX_RAYS];
    float Z = 0.f; f3 dZ = mk3(0, 0, 0);
    const AuxFrame fr = aux_frame(d, cfg.kappa);
    for (uint32_t i = 0; i < n; ++i) {
        const float sx = pcg_f32(aux_rng), sy = pcg_f32(aux_rng);
        f3 tang; const f3 ad = warp_field_direction(d, fr, sx, sy, cfg.kappa, tang);
        const Hit h = io.aux_hit(aux_base + i, Ray{o, ad, FLT_MAX});
        warp_field_weights(sc, h, ad, tang, sy, cfg.kappa, fr.e2k, cfg.exponent, rec[i]);
        Z += rec[i].w; dZ = dZ + rec[i].dw;
    }
    Z = fmaxf(Z, 1e-8f);
    const float iz = 1.0f / Z;
    const f3 g_V = (normalize_bwd(d, g_dir) - dZ * (g_det * iz)) * iz;
    const float g_div = g_det * iz;
    f3 g_o = mk3(0, 0, 0), g_d = mk3(0, 0, 0);
    // The auxiliary rays form a tight cone (κ = 1e5 by default): consecutive ones mostly land on the same triangle, so their
    // vertex gradients are summed in registers and scattered once per run of equal triangles (9 atomics per run instead of
    // 9 per ray: the scatter was the bulk of the adjoint passes, profiles/r02_s_*).
    uint32_t run = 0xffffffffu; f3 a0 = mk3(0, 0, 0), a1 = mk3(0, 0, 0), a2 = mk3(0, 0, 0);
    for (uint32_t i = 0; i < n; ++i) {
        const f3 g_vd = g_V * rec[i].w + rec[i].dw * g_div;
        if (rec[i].prim != 0xffffffffu) {
            const uint32_t pr = rec[i].prim;
            const float b1 = rec[i].u, b2 = rec[i].v, b0 = 1.f - b1 - b2;
            const f3 p = fma3(ld3(sc.pos, sc.faces[3 * pr]), b0, fma3(ld3(sc.pos, sc.faces[3 * pr + 1]), b1, ld3(sc.pos, sc.faces[3 * pr + 2]) * b2));
            const f3 g_r = normalize_bwd(p - o, g_vd);
            if (pr != run) { scatter_tri3(sc, sink, run, a0, a1, a2); run = pr; a0 = a1 = a2 = mk3(0, 0, 0); }
            a0 = a0 + g_r * b0; a1 = a1 + g_r * b1; a2 = a2 + g_r * b2;
            g_o = g_o - g_r;
        } else {
            g_d = g_d + g_vd;
        }
    }
    scatter_tri3(sc, sink, run, a0, a1, a2);
    if (d_attached) g_o = g_o - normalize_bwd(d_from - o, g_d);                    // d = normalize(d_from − o), spawn_ray_to (interaction.h:141-147)
    scatter_tri(sc, sink, anchor.prim, anchor.bw[0], anchor.bw[1], anchor.bw[2], g_o);
}

// ------------------------------------------------------------------------------------------------
//  Adjoint of the shading frame of si_cur (mesh.cpp:676-822, interaction.h:262-275,499) for the hit (ray.o detached,
//  ray.d = reparameterised direction).  Inputs: gradients w.r.t. the local directions wi = to_local(−d), wo_a = to_local(a),
//  wo_b = to_local(b).  Scatters into the triangle's positions / normals, returns the gradients of d and a.
// ------------------------------------------------------------------------------------------------
template <typename Sink>
EPSM_HDN void frame_backward_si(const DScene &sc, Sink &sink, const SI &si, f3 ray_o, f3 ray_d, f3 a, f3 b, f3 g_wi, f3 g_wa, f3 g_wb, f3 &g_d, f3 &g_a) {
    const f3 s = si.sh_s, t = si.sh_t, n = si.sh_n;
    f3 g_s = a * g_wa.x + b * g_wb.x - ray_d * g_wi.x;
    f3 g_t = a * g_wa.y + b * g_wb.y - ray_d * g_wi.y;
    f3 g_n = a * g_wa.z + b * g_wb.z - ray_d * g_wi.z;
    g_d = g_d - (s * g_wi.x + t * g_wi.y + n * g_wi.z);
    g_a = g_a + (s * g_wa.x + t * g_wa.y + n * g_wa.z);
    const bool grad = (si.mesh_flags & EPSM_MESH_GRAD) != 0;
    const bool flip = (si.mesh_flags & EPSM_MESH_FLIP_NORMALS) != 0;
    // t = n × s
    g_n = g_n + cross3(s, g_t); g_s = g_s + cross3(g_t, n);
    // s = normalize(q), q = dp_du − n·(n·dp_du); dp_du = coordinate_system(n_g).s with n_g the UNFLIPPED face normal
    const f3 p0 = ld3(sc.pos, si.fi[0]), p1 = ld3(sc.pos, si.fi[1]), p2 = ld3(sc.pos, si.fi[2]);
    const f3 e1 = p1 - p0, e2 = p2 - p0, ng_raw = cross3(e1, e2), ng = norm3(ng_raw);
    f3 dp_du, dp_dv; coordinate_system(ng, dp_du, dp_dv);
    // meshes with texture coordinates: dp_du = (Δv₁·e1 − Δv₀·e2)/det (mesh.cpp:746-760), linear in the edges
    float ka = 0.f, kb = 0.f; bool uv_frame = false;
    if (sc.meshes[si.mesh].has_uv) {
        const float *t0 = sc.uv + 2 * (size_t) si.fi[0], *t1 = sc.uv + 2 * (size_t) si.fi[1], *t2 = sc.uv + 2 * (size_t) si.fi[2];
        const float du0 = t1[0] - t0[0], dv0 = t1[1] - t0[1], du1 = t2[0] - t0[0], dv1 = t2[1] - t0[1];
        const float det = fmaf(du0, dv1, -(dv0 * du1));
        if (det != 0.f) { const float inv_det = rcp_e(det); ka = dv1 * inv_det; kb = -dv0 * inv_det; dp_du = e1 * ka + e2 * kb; uv_frame = true; }
    }
    const float c = dot3(n, dp_du);
    const f3 q = fma3(n, -c, dp_du);
    const f3 g_q = normalize_bwd(q, g_s);
    const float ngq = dot3(n, g_q);
    const f3 g_dpdu = g_q - n * ngq;
    g_n = g_n - g_q * c - dp_du * ngq;
    f3 g_ng = uv_frame ? mk3(0, 0, 0) : onb_s_bwd(ng, g_dpdu);
    f3 g_shn = flip ? -g_n : g_n;                                                  // gradient of the unflipped shading normal
    float g_u = 0.f, g_v = 0.f;
    if (si.has_normals) {
        const f3 n0 = ld3(sc.nrm, si.fi[0]), n1 = ld3(sc.nrm, si.fi[1]), n2 = ld3(sc.nrm, si.fi[2]);
        const f3 ns = fma3(n2, si.bw[2], fma3(n1, si.bw[1], n0 * si.bw[0]));
        const f3 g_ns = normalize_bwd(ns, g_shn);
        if (grad) { sink.nrm(si.fi[0], g_ns * si.bw[0]); sink.nrm(si.fi[1], g_ns * si.bw[1]); sink.nrm(si.fi[2], g_ns * si.bw[2]); }
        g_u = dot3(n1 - n0, g_ns); g_v = dot3(n2 - n0, g_ns);                     // b0 = 1 − u − v
    } else {
        g_ng = g_ng + g_shn;
    }
    const f3 g_raw = normalize_bwd(ng_raw, g_ng);
    f3 g_p1 = cross3(e2, g_raw), g_p2 = cross3(g_raw, e1);                          // ng_raw = e1 × e2
    if (uv_frame) { g_p1 = g_p1 + g_dpdu * ka; g_p2 = g_p2 + g_dpdu * kb; }
    f3 g_p0 = -(g_p1 + g_p2);
    if (g_u != 0.f || g_v != 0.f) {                                                // (u, v) re-derived by Möller–Trumbore (mesh.cpp:676-686)
        MtGrad M; mt_uv_bwd(ray_o, ray_d, p0, p1, p2, g_u, g_v, M);
        g_d = g_d + M.d; g_p0 = g_p0 + M.p0; g_p1 = g_p1 + M.p1; g_p2 = g_p2 + M.p2;
    }
    if (grad) { sink.pos(si.fi[0], g_p0); sink.pos(si.fi[1], g_p1); sink.pos(si.fi[2], g_p2); }
}

// Inline primal replay for the policies that have no per-lane radiance buffer (MegaIo): EPSMIntegrator.sample's loop is path_primal.
struct LocalTrace {};
template <bool COUNT>
EPSM_HD f3 path_primal_io(const DScene &sc, MegaIo<COUNT> &io, LocalTrace &, Ray ray, Pcg &rng, uint32_t max_depth, uint32_t rr_depth, bool &valid) {
    return path_primal<COUNT>(sc, io.tr, ray, rng, max_depth, rr_depth, valid);
}

// GaussianFilter::eval and its derivative (gaussian.cpp:94-96): f = max(0, P(x²)), f' = 2x·P'(x²) where P > 0
EPSM_HD void filter_eval_d(const float *c, float x, float &f, float &df) {
    const float t = x * x;
    float p = c[9], dp = 0.f;
    for (int k = 8; k >= 0; --k) { dp = fmaf(dp, t, p); p = fmaf(p, t, c[k]); }
    f = filter_eval(c, x);
    df = f > 0.f ? 2.f * x * dp : 0.f;
}

// ------------------------------------------------------------------------------------------------
//  One lane of RBIntegrator.render_backward for prb_reparam.  `film` is the RGBW film of the primal pass over ALL lanes
//  (common.py:905-935), grad_in the (H,W,3) image gradient.  Gradients are pushed into `sink` (.pos / .nrm of a vertex).
// ------------------------------------------------------------------------------------------------
struct PrevVertex { bool valid; f3 p, sh_s, sh_t, sh_n, wi; int32_t bsdf; float uv[2]; Anchor anchor; };
// Everything iteration `it` of the loop below inherits from the iterations before it.  The replay wavefront stores it when a
// pass first reaches the top of an iteration (pass 2·it + 1) and the two later passes that need that iteration (2·it + 2: its
// second round of rays, 2·it + 3: its adjoint) start from there instead of replaying the lane from the camera — without it the
// replay work grows with the square of the path length (profiles/r02_h_*: k_prb_pass 333 of 500 ms at 512² × 8 spp).
struct PrbCheckpoint { Pcg rng, aux; Ray ray; Hit h; f3 L, beta, dL; float eta, mis_em; uint32_t depth, it; PrevVertex prev; };

template <typename Io, typename Sink>
EPSM_HDN void prb_backward_lane(const DScene &sc, Io &io, const DSensor &sensor, const PrbConfig &cfg, uint32_t px, uint32_t py,
                                uint32_t seed, uint32_t lane, const float *film, const float *grad_in, Sink &sink) {
    Pcg rng, aux; Hit h; f3 dL = mk3(0, 0, 0);
    const uint32_t n_aux = cfg.n_rays;
    const Anchor none{0xffffffffu, {0.f, 0.f, 0.f}};
    uint32_t depth = 0, it0 = 0; f3 L, beta = mk3(1, 1, 1); float eta = 1.f, mis_em = 1.f; bool active = true;
    Ray ray;
    PrevVertex prev; prev.valid = false; prev.anchor = none; prev.bsdf = -1;
    PrbCheckpoint ck;
    if (io.resume(ck)) {
        rng = ck.rng; aux = ck.aux; ray = ck.ray; h = ck.h; L = ck.L; beta = ck.beta; dL = ck.dL; eta = ck.eta; mis_em = ck.mis_em; depth = ck.depth; it0 = ck.it; prev = ck.prev;
    } else {
    rng = pcg_for_lane(seed, lane);
    aux = pcg_for_lane(0xffffffffu ^ seed, lane);                                   // _ReparamWrapper, common.py:1200-1203
    CamRay cr; camera_ray(sensor, px, py, rng, cr);
    const bool cam_active = 0 < cfg.reparam_max_depth;

    // ---- round 0: the primary ray and the auxiliary rays of the camera ray's reparameterisation (common.py:372-397)
    h = io.closest(0, 0, cr.ray);
    f3 Ltot = mk3(0, 0, 0);
    f3 g_d_cam = mk3(0, 0, 0); float g_det_cam = 0.f;
    if (!io.aux_push_only(0)) {
        // primal radiance of the lane: the primal pass with the cloned sampler (common.py:905-915)
        if (!io.lane_radiance(Ltot)) {
            LocalTrace lt; bool valid_path; Pcg rng_p = rng;
            Ltot = path_primal_io(sc, io, lt, cr.ray, rng_p, cfg.max_depth, cfg.rr_depth, valid_path);
        }
        // ---- adjoint of block.put(pos, L·det, weight = det) → develop (common.py:919-945): δL, ∂/∂pos, ∂/∂det
        float g_px = 0.f, g_py = 0.f;
        {
            const int FW = (int) sensor.fw, FH = (int) sensor.fh;
            const int pix = (int) floorf(cr.pos_x) - 2, piy = (int) floorf(cr.pos_y) - 2;
            float rel_x = (float) pix + .5f - cr.pos_x, rel_y = (float) piy + .5f - cr.pos_y;
            float wx[5], wy[5], dwx[5], dwy[5];
            for (int k = 0; k < 5; ++k) { filter_eval_d(sc.filter, rel_x, wx[k], dwx[k]); filter_eval_d(sc.filter, rel_y, wy[k], dwy[k]); rel_x += 1; rel_y += 1; }
            for (int ys = 0; ys < 5; ++ys) for (int xs = 0; xs < 5; ++xs) {
                const int x = pix + xs, y = piy + ys;
                if (x < 0 || y < 0 || x >= FW || y >= FH) continue;                // cropped border (hdrfilm.cpp:286-320)
                const float *fp = film + ((size_t) y * FW + x) * 4, *gp = grad_in + ((size_t) y * FW + x) * 3;
                const float W = fp[3], iw = W == 0.f ? 1.0f : 1.0f / W, nz = W == 0.f ? 0.f : 1.f;    // develop: rgb / (w == 0 ? 1 : w)
                const float w = wy[ys] * wx[xs];
                dL = dL + mk3(gp[0], gp[1], gp[2]) * (w * iw);
                const float A = (gp[0] * (Ltot.x - nz * fp[0] * iw) + gp[1] * (Ltot.y - nz * fp[1] * iw) + gp[2] * (Ltot.z - nz * fp[2] * iw)) * iw;
                g_det_cam += w * A; g_px -= wy[ys] * dwx[xs] * A; g_py -= dwy[ys] * wx[xs] * A;
            }
        }
        // pos = resolution · project(camera_to_sample · to_world⁻¹ · (o + d)) (common.py:390-396, perspective.cpp:282-306)
        {
            const f3 itp = cr.ray.o + cr.ray.d;
            const float *m = sensor.w2c, *c2s = sensor.c2s;
            const f3 q = mk3(m[0] * itp.x + m[1] * itp.y + m[2] * itp.z + m[3], m[4] * itp.x + m[5] * itp.y + m[6] * itp.z + m[7],
                             m[8] * itp.x + m[9] * itp.y + m[10] * itp.z + m[11]);
            const float hx = c2s[0] * q.x + c2s[1] * q.y + c2s[2] * q.z + c2s[3], hy = c2s[4] * q.x + c2s[5] * q.y + c2s[6] * q.z + c2s[7],
                        hw = c2s[12] * q.x + c2s[13] * q.y + c2s[14] * q.z + c2s[15];
            const float gux = g_px * (float) sensor.fw, guy = g_py * (float) sensor.fh, ihw = 1.0f / hw;
            const float ghx = gux * ihw, ghy = guy * ihw, ghw = -(gux * hx + guy * hy) * ihw * ihw;
            const f3 g_q = mk3(c2s[0] * ghx + c2s[4] * ghy + c2s[12] * ghw, c2s[1] * ghx + c2s[5] * ghy + c2s[13] * ghw, c2s[2] * ghx + c2s[6] * ghy + c2s[14] * ghw);
            g_d_cam = mk3(m[0] * g_q.x + m[4] * g_q.y + m[8] * g_q.z, m[1] * g_q.x + m[5] * g_q.y + m[9] * g_q.z, m[2] * g_q.x + m[6] * g_q.y + m[10] * g_q.z);
        }
    }
    reparam_backward(sc, io, sink, aux, cfg, 0, 0, cam_active, cr.ray.o, cr.ray.d, g_d_cam, g_det_cam, none, false, mk3(0, 0, 0));
    if (io.stop_after(0)) return;
    L = Ltot; ray = cr.ray;
    }

    // ---- sample(mode = Backward) (prb_reparam.py:329-602)
    SI si, si_next;
    for (uint32_t it = it0; it < cfg.max_depth && active; ++it) {
        const int r1 = 1 + 2 * (int) it, r2 = 2 + 2 * (int) it;                     // this iteration's two rounds of rays
        if (Io::wavefront) {
            ck.rng = rng; ck.aux = aux; ck.ray = ray; ck.h = h; ck.L = L; ck.beta = beta; ck.dL = dL; ck.eta = eta; ck.mis_em = mis_em; ck.depth = depth; ck.it = it; ck.prev = prev;
            io.checkpoint(it, ck);
        }
        compute_si(sc, ray, h, si);
        const f3 Le = beta * mis_em * emitter_eval(sc, si);
        bool active_next = (depth + 1 < cfg.max_depth) && si.valid;
        DBsdf b; uint32_t flags = 0;
        if (si.valid) { b = bsdf_at(sc, si.bsdf, si.uv); flags = b.bflags; }
        bool active_em = active_next && (flags & EPSM_BF_SMOOTH);
        const float e1 = pcg_f32(rng), e2 = pcg_f32(rng);
        DSamp ds; f3 em_weight;
        sample_emitter_direction(sc, si, e1, e2, active_em, ds, em_weight, [&](const Ray &r) { return io.occluded(r1, 2 * it, r); });
        active_em &= ds.pdf != 0.f;
        Ray em_ray{mk3(0, 0, 0), mk3(0, 0, 1), 0.f};
        f3 Lr_dir = mk3(0, 0, 0); float mis_direct = 0.f;
        if (active_em) {
            em_ray = spawn_ray_to(si, ds.p);
            f3 val; float bpdf;
            bsdf_eval_pdf(b, si.wi, to_local(si, ds.d), val, bpdf);
            mis_direct = mis_weight(ds.pdf, bpdf);
            Lr_dir = beta * mis_direct * val * em_weight;
        }
        const float s1 = pcg_f32(rng), s2x = pcg_f32(rng), s2y = pcg_f32(rng);
        BSample bs; bs.wo = bs.hf = bs.weight = mk3(0, 0, 0); bs.pdf = 0.f; bs.eta = 0.f; bs.type = 0;
        if (si.valid) bsdf_sample(b, si.wi, s1, s2x, s2y, active_next, bs);
        const f3 beta_in = beta;
        eta *= bs.eta; beta = beta * bs.weight;
        const f3 L_prev = L; L = L - Le - Lr_dir;
        const float bmax = max3(beta);
        active_next &= bmax != 0.f;
        const float rr_prob = fminf(bmax * eta * eta, .95f);
        const bool rr_active = depth >= cfg.rr_depth;
        if (rr_active) beta = beta * rcp_e(rr_prob);
        const bool rr_continue = pcg_f32(rng) < rr_prob;
        active_next &= !rr_active || rr_continue;
        Ray ray_next = ray;
        if (si.valid) ray_next = spawn_ray(si, to_world(si, bs.wo));
        Hit h_next{0.f, 0.f, 0.f, 0xffffffffu};
        if (active_next) h_next = io.closest(r1, 1 + it, ray_next);
        if (io.stop_after(r1)) return;                                             // round r1 queued: NEE visibility and the next hit are pending
        compute_si(sc, ray_next, h_next, si_next);
        const float pdf_em = pdf_emitter_direction(sc, si_next, si.p, !(bs.type & EPSM_BF_DELTA));
        const float mis_next = mis_weight(bs.pdf, pdf_em);
        // run-ahead with a clone of the sampler (prb_reparam.py:484-520)
        Pcg clone = rng;
        const bool active_nn = active_next && si_next.valid && (depth + 2 < cfg.max_depth);
        DBsdf bn; uint32_t nflags = 0;
        if (si_next.valid) { bn = bsdf_at(sc, si_next.bsdf, si_next.uv); nflags = bn.bflags; }
        bool active_em_n = active_nn && (nflags & EPSM_BF_SMOOTH);
        const float c1 = pcg_f32(clone), c2 = pcg_f32(clone);
        DSamp dsn; f3 emw_n;
        sample_emitter_direction(sc, si_next, c1, c2, active_em_n, dsn, emw_n, [&](const Ray &r) { return io.occluded(r2, 2 * it + 1, r); });
        active_em_n &= dsn.pdf != 0.f;
        f3 Lr_dir_next = mk3(0, 0, 0);
        if (active_em_n) {
            f3 valn; float bpdfn;
            bsdf_eval_pdf(bn, si_next.wi, to_local(si_next, dsn.d), valn, bpdfn);
            Lr_dir_next = beta * mis_weight(dsn.pdf, bpdfn) * valn * emw_n;
        }
        const float c3 = pcg_f32(clone), c4 = pcg_f32(clone), c5 = pcg_f32(clone);
        BSample bsn; bsn.wo = bsn.hf = bsn.weight = mk3(0, 0, 0); bsn.pdf = 0.f; bsn.eta = 0.f; bsn.type = 0;
        if (si_next.valid) bsdf_sample(bn, si_next.wi, c3, c4, c5, active_nn, bsn);
        const f3 Le_next = active_next ? beta * mis_next * emitter_eval(sc, si_next) : mk3(0, 0, 0);
        const f3 L_next = L - Le_next - Lr_dir_next;

        // ---- adjoint of Lo = (Le + Lr_dir + Lr_ind)·det + extra, seeded with δL (prb_reparam.py:576-588); in the wavefront it
        //      runs in the one pass in which the iteration's auxiliary hits are available
        f3 g_d = mk3(0, 0, 0), g_de = mk3(0, 0, 0); float g_det = 0.f, g_det_e = 0.f;
        const bool have_dl = dL.x != 0.f || dL.y != 0.f || dL.z != 0.f;
        if (io.adjoint_now(r2) && si.valid && have_dl) {
            // Lr_dir = β·mis·f(wi, to_local(d_e))·em_weight·det_e (:431-435);  Lr_ind = L·f(wi, to_local(ray_next.d))/detach(f) (:553-571)
            const f3 bv = bs.weight * bs.pdf;
            const f3 inv_bv = mk3(bv.x != 0.f ? 1.f / bv.x : 0.f, bv.y != 0.f ? 1.f / bv.y : 0.f, bv.z != 0.f ? 1.f / bv.z : 0.f);
            f3 g_wi = mk3(0, 0, 0), g_wa = mk3(0, 0, 0), g_wb = mk3(0, 0, 0), val;
            const f3 da = active_em ? ds.d : mk3(0, 0, 0), db = active_next ? ray_next.d : mk3(0, 0, 0);
            if (active_em) { f3 gi; bsdf_grad(b, si.wi, to_local(si, ds.d), dL * beta_in * mis_direct * em_weight, gi, g_wa, val); g_wi = g_wi + gi; }
            if (active_next) { f3 gi; bsdf_grad(b, si.wi, to_local(si, ray_next.d), dL * L * inv_bv, gi, g_wb, val); g_wi = g_wi + gi; }
            frame_backward_si(sc, sink, si, ray.o, ray.d, da, db, g_wi, g_wa, g_wb, g_d, g_de);
            // adjacent vertices through si_cur.p with the shape detached (:522-551)
            f3 g_pro = mk3(0, 0, 0);
            if (prev.valid && prev.bsdf >= 0) {
                const DBsdf bp = bsdf_at(sc, prev.bsdf, prev.uv);
                const f3 r = si.p - prev.p, w = norm3(r), wl = mk3(dot3(w, prev.sh_s), dot3(w, prev.sh_t), dot3(w, prev.sh_n));
                f3 f0, gi, go; bsdf_grad(bp, prev.wi, wl, mk3(0, 0, 0), gi, go, f0);
                const f3 aw = dL * L_prev * mk3(1.f / fmaxf(1e-8f, f0.x), 1.f / fmaxf(1e-8f, f0.y), 1.f / fmaxf(1e-8f, f0.z));
                bsdf_grad(bp, prev.wi, wl, aw, gi, go, f0);
                g_pro = g_pro + normalize_bwd(r, prev.sh_s * go.x + prev.sh_t * go.y + prev.sh_n * go.z);
            }
            if (si_next.valid) {
                const f3 r = si.p - si_next.p, w = norm3(r), wl = to_local(si_next, w);
                f3 f0, gi, go; bsdf_grad(bn, wl, bsn.wo, mk3(0, 0, 0), gi, go, f0);
                const f3 aw = dL * L_next * mk3(1.f / fmaxf(1e-8f, f0.x), 1.f / fmaxf(1e-8f, f0.y), 1.f / fmaxf(1e-8f, f0.z));
                bsdf_grad(bn, wl, bsn.wo, aw, gi, go, f0);
                g_pro = g_pro + normalize_bwd(r, si_next.sh_s * gi.x + si_next.sh_t * gi.y + si_next.sh_n * gi.z);
            }
            if (g_pro.x != 0.f || g_pro.y != 0.f || g_pro.z != 0.f) {
                const f3 p0 = ld3(sc.pos, si.fi[0]), p1 = ld3(sc.pos, si.fi[1]), p2 = ld3(sc.pos, si.fi[2]);
                MtGrad M; mt_uv_bwd(ray.o, ray.d, p0, p1, p2, dot3(p1 - p0, g_pro), dot3(p2 - p0, g_pro), M);
                g_d = g_d + M.d;
            }
            const f3 Lsum = Le + Lr_dir + L;
            g_det = it == 0 ? 0.f : dot3(dL, Lsum);                                // ray_reparam_det[first_vertex] = 1 (:364)
            g_det_e = dot3(dL, Lr_dir);
        } else if (io.adjoint_now(r2) && have_dl && it != 0) {
            g_det = dot3(dL, Le + L);                                              // escaped ray: Lo = (Le + L)·det with Le = 0 here
        }
        // the emitter ray's custom op is created later in the iteration ⇒ its backward runs first; both draw from one generator
        Anchor cur{si.valid ? si.prim : 0xffffffffu, {si.bw[0], si.bw[1], si.bw[2]}};
        reparam_backward(sc, io, sink, aux, cfg, r2, 0, active_em && (depth + 1 < cfg.reparam_max_depth), em_ray.o, em_ray.d, g_de, g_det_e, cur, true, ds.p);
        reparam_backward(sc, io, sink, aux, cfg, r2, n_aux, depth < cfg.reparam_max_depth, ray.o, ray.d, g_d, g_det, prev.anchor, false, mk3(0, 0, 0));
        if (io.stop_after(r2)) return;                                             // round r2 queued: the adjoint of this iteration runs next pass

        prev.valid = si.valid; prev.p = si.p; prev.sh_s = si.sh_s; prev.sh_t = si.sh_t; prev.sh_n = si.sh_n; prev.wi = si.wi; prev.bsdf = si.bsdf; prev.uv[0] = si.uv[0]; prev.uv[1] = si.uv[1]; prev.anchor = cur;
        mis_em = mis_next; h = h_next; ray = ray_next;
        if (si.valid) depth += 1;
        active = active_next;
    }
    io.finished();
}
