// hostsim.cu — TEST-ONLY harness.  Compiles the product's per-lane device code (epsm_device.cuh) for the
// HOST so that the exact lane logic (traversal, path loops, manifold solve, scatter seeds) can be checked
// against the oracle on a machine without a GPU.  It is not a fallback: nothing in epsm_mitsuba3_b200
// loads this library; the product path is libepsm_b200.so on a CUDA device only.
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../epsm_mitsuba3_b200/csrc/epsm_device.cuh"
#include "../../epsm_mitsuba3_b200/csrc/epsm_prb.cuh"
#include "../../epsm_mitsuba3_b200/csrc/epsm_host_scene.h"

struct sim_scene { HostScene hs; DScene sc; std::string err; };

namespace {
struct LocalStack { uint32_t ref[EPSM_SSTACK]; float t[EPSM_SSTACK]; };
inline void bind(TStack &st, LocalStack &ls) { st.s_ref = ls.ref; st.s_t = ls.t; st.stride = 1; st.sp = 0; }

DLaunch make_launch(const sim_scene *s, uint32_t sensor, uint32_t seed, uint32_t spp, uint32_t max_depth, uint32_t rr_depth, int variant) {
    DLaunch ln; memset(&ln, 0, sizeof(ln));
    ln.sensor = s->hs.sensors[sensor]; ln.seed = seed; ln.spp = spp; ln.max_depth = max_depth; ln.rr_depth = rr_depth; ln.variant = variant; ln.s_begin = 0; ln.s_end = spp;
    return ln;
}
void add3(float *buf, uint32_t idx, f3 v, bool pred) { if (!pred) return; buf[3 * idx] += v.x; buf[3 * idx + 1] += v.y; buf[3 * idx + 2] += v.z; }
}

extern "C" {

sim_scene *sim_scene_create(const epsm_mesh_desc *meshes, uint32_t n_meshes, const epsm_bsdf_desc *bsdfs, uint32_t n_bsdfs,
                            const epsm_emitter_desc *emitters, uint32_t n_emitters, const epsm_sensor_desc *sensors, uint32_t n_sensors) {
    sim_scene *s = new sim_scene();
    s->err = host_scene_build(s->hs, meshes, n_meshes, bsdfs, n_bsdfs, emitters, n_emitters, sensors, n_sensors);
    memset(&s->sc, 0, sizeof(s->sc));
    if (s->err.empty()) host_scene_view(s->hs, s->sc);
    return s;
}
const char *sim_scene_error(sim_scene *s) { return s->err.c_str(); }
void sim_scene_destroy(sim_scene *s) { delete s; }
void sim_scene_update_vertices(sim_scene *s, uint32_t mesh_id, const float *positions) {
    const DMesh &m = s->hs.meshes[mesh_id];
    memcpy(s->hs.pos.data() + 3 * (size_t) m.v_off, positions, 3 * (size_t) m.v_cnt * sizeof(float));
    host_scene_rebuild_accel(s->hs); host_scene_view(s->hs, s->sc);
}
// The device refit (k_refit_tris / k_refit_nodes) run serially: children are emitted after their parent by the builder,
// so descending node order is a valid bottom-up schedule.  Returns the refitted tree's cost (Σ node half-areas).
double sim_scene_refit(sim_scene *s, uint32_t mesh_id, const float *positions) {
    const DMesh &m = s->hs.meshes[mesh_id];
    memcpy(s->hs.pos.data() + 3 * (size_t) m.v_off, positions, 3 * (size_t) m.v_cnt * sizeof(float));
    f4 *nodes = s->hs.bvh.nodes.data(), *tris = s->hs.bvh.tris.data();
    const uint32_t n_nodes = (uint32_t) (s->hs.bvh.nodes.size() / EPSM_NODE_F4), n_slots = (uint32_t) (s->hs.bvh.tris.size() / EPSM_TRI_F4);
    float max_abs = 0.f;
    for (uint32_t i = 0; i < n_slots; ++i) max_abs = fmaxf(max_abs, refit_tri_slot(tris, s->hs.faces.data(), s->hs.pos.data(), i));
    const float pad = refit_pad(max_abs);
    double cost = 0;
    for (uint32_t i = n_nodes; i-- > 0;) {
        float lo[3], hi[3];
        cost += (double) refit_node(nodes, i, tris, s->hs.faces.data(), s->hs.pos.data(), pad, lo, hi);
        const uint32_t link = f2u(nodes[(size_t) i * EPSM_NODE_F4 + 7].x);
        if (link != 0xffffffffu) refit_store_parent(nodes, link, lo, hi);
    }
    host_scene_rebuild_area_tables(s->hs); host_scene_view(s->hs, s->sc);
    return cost;
}
double sim_scene_cost(sim_scene *s) { return epsm_bvh4_cost(s->hs.bvh.nodes.data(), s->hs.bvh.nodes.size() / EPSM_NODE_F4); }
void sim_scene_copy_accel(sim_scene *s, float *nodes, float *tris) {
    if (nodes) memcpy(nodes, s->hs.bvh.nodes.data(), s->hs.bvh.nodes.size() * sizeof(f4));
    if (tris) memcpy(tris, s->hs.bvh.tris.data(), s->hs.bvh.tris.size() * sizeof(f4));
}
// Texture-side probe of a hit with the product's compute_si / bsdf_at: out = [uv(2), sh_s(3), sh_t(3), reflectance(3)]
void sim_surface_uv(sim_scene *s, const float *r, float *out) {
    LocalStack ls; TStack st; bind(st, ls); Tracer tr(s->sc, st);
    Ray ray{mk3(r[0], r[1], r[2]), mk3(r[4], r[5], r[6]), r[3]};
    const Hit h = trace_closest<false>(tr, ray);
    SI si; compute_si(s->sc, ray, h, si);
    for (int i = 0; i < 11; ++i) out[i] = 0.f;
    if (!si.valid) return;
    const DBsdf b = bsdf_at(s->sc, si.bsdf, si.uv);
    out[0] = si.uv[0]; out[1] = si.uv[1]; out[2] = si.sh_s.x; out[3] = si.sh_s.y; out[4] = si.sh_s.z; out[5] = si.sh_t.x; out[6] = si.sh_t.y; out[7] = si.sh_t.z;
    out[8] = b.reflectance[0]; out[9] = b.reflectance[1]; out[10] = b.reflectance[2];
}
// The product's BSDF code on its own (for the known answers of the reference's BSDF tests).
void sim_bsdf_sample(sim_scene *s, uint32_t bsdf_id, const float *wi, float s1, float s2x, float s2y, float *out12) {
    BSample bs; bsdf_sample(s->hs.bsdfs[bsdf_id], mk3(wi[0], wi[1], wi[2]), s1, s2x, s2y, true, bs);
    const float v[12] = {bs.wo.x, bs.wo.y, bs.wo.z, bs.pdf, bs.eta, (float) bs.type, bs.weight.x, bs.weight.y, bs.weight.z, bs.hf.x, bs.hf.y, bs.hf.z};
    memcpy(out12, v, sizeof(v));
}
void sim_bsdf_eval_pdf(sim_scene *s, uint32_t bsdf_id, const float *wi, const float *wo, float *out4) {
    f3 value; float pdf; bsdf_eval_pdf(s->hs.bsdfs[bsdf_id], mk3(wi[0], wi[1], wi[2]), mk3(wo[0], wo[1], wo[2]), value, pdf);
    out4[0] = value.x; out4[1] = value.y; out4[2] = value.z; out4[3] = pdf;
}
// Σ_c a_c·f_c and its gradient w.r.t. (wi, wo) from the 6-wide forward duals of epsm_prb.cuh: out = value(3), g_wi(3), g_wo(3)
void sim_bsdf_grad(sim_scene *s, uint32_t bsdf_id, const float *wi, const float *wo, const float *a, float *out9) {
    f3 gi, go, v; bsdf_grad(s->hs.bsdfs[bsdf_id], mk3(wi[0], wi[1], wi[2]), mk3(wo[0], wo[1], wo[2]), mk3(a[0], a[1], a[2]), gi, go, v);
    const float o[9] = {v.x, v.y, v.z, gi.x, gi.y, gi.z, go.x, go.y, go.z}; memcpy(out9, o, sizeof(o));
}
void sim_scene_stats(sim_scene *s, uint64_t *out) {
    out[0] = s->hs.faces.size() / 3; out[1] = s->hs.bvh.nodes.size() / EPSM_NODE_F4; out[2] = s->hs.bvh.depth;
}

void sim_trace(sim_scene *s, const float *rays, uint64_t n, void *hits, int any_hit, uint64_t *counters) {
    uint64_t cn = 0, ct = 0;
    #pragma omp parallel for schedule(dynamic, 256) reduction(+ : cn, ct)
    for (int64_t i = 0; i < (int64_t) n; ++i) {
        LocalStack ls; TStack st; bind(st, ls); Tracer tr(s->sc, st);
        const float *r = rays + 8 * i;
        Ray ray{mk3(r[0], r[1], r[2]), mk3(r[4], r[5], r[6]), r[3]};
        if (any_hit) ((uint32_t *) hits)[i] = trace_any<true>(tr, ray) ? 1u : 0u;
        else {
            Hit h = trace_closest<true>(tr, ray);
            uint32_t *o = (uint32_t *) hits + 4 * i; o[0] = f2u(h.t); o[1] = f2u(h.u); o[2] = f2u(h.v); o[3] = h.prim;
        }
        cn += tr.n_nodes; ct += tr.n_tris;
    }
    if (counters) { counters[0] += cn; counters[1] += ct; counters[2] += n; counters[3] += n; }
}

void sim_sampler_stream(uint32_t seed, uint64_t lane_begin, uint64_t lane_end, uint32_t n, float *out) {
    for (uint64_t l = lane_begin; l < lane_end; ++l) { Pcg r = pcg_for_lane(seed, (uint32_t) l); for (uint32_t k = 0; k < n; ++k) out[(l - lane_begin) * n + k] = pcg_f32(r); }
}
void sim_sample_rays(sim_scene *s, uint32_t sensor, uint32_t seed, uint32_t spp, uint64_t lane_begin, uint64_t lane_end, float *out) {
    const DSensor &se = s->hs.sensors[sensor];
    for (uint64_t lane = lane_begin; lane < lane_end; ++lane) {
        uint32_t pixel = (uint32_t) (lane / spp); Pcg r = pcg_for_lane(seed, (uint32_t) lane);
        CamRay c; camera_ray(se, pixel % se.w, pixel / se.w, r, c);
        float *o = out + 14 * (lane - lane_begin);
        o[0] = c.ray.o.x; o[1] = c.ray.o.y; o[2] = c.ray.o.z; o[3] = c.ray.d.x; o[4] = c.ray.d.y; o[5] = c.ray.d.z;
        o[6] = c.d_x.x; o[7] = c.d_x.y; o[8] = c.d_x.z; o[9] = c.d_y.x; o[10] = c.d_y.y; o[11] = c.d_y.z; o[12] = c.pos_x; o[13] = c.pos_y;
    }
}

// k_render on the host: film is H·W·4, accumulated into.
// `staged` != 0: the lane runs the wavefront stages of k_pr_stage (pr_stage with the NEE term parked until its shadow
// ray has answered) instead of path_primal.
void sim_render_accumulate(sim_scene *s, uint32_t sensor, uint32_t seed, uint32_t spp, uint32_t max_depth, uint32_t rr_depth,
                           uint64_t lane_begin, uint64_t lane_end, float *film, int staged) {
    DLaunch ln = make_launch(s, sensor, seed, spp, max_depth, rr_depth, 0);
    const uint32_t W = ln.sensor.w;                                   // sampling grid width (lane → pixel)
    const uint32_t FW = ln.sensor.fw, FH = ln.sensor.fh;               // film
    const int64_t n = (int64_t) (lane_end - lane_begin);
    std::vector<float> tmp((size_t) n * 5);
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; ++i) {
        uint64_t lane = lane_begin + i; uint32_t pixel = (uint32_t) (lane / spp);
        LocalStack ls; TStack st; bind(st, ls); Tracer tr(s->sc, st);
        Pcg rng = pcg_for_lane(seed, (uint32_t) lane);
        CamRay c; camera_ray(ln.sensor, pixel % W, pixel / W, rng, c);
        f3 L;
        if (staged) {
            PrState S; S.rng = rng; S.ray = c.ray; pr_state_defaults(S);
            bool occluded = false;
            for (uint32_t it = 0; it < max_depth && (S.bits & PR_ACTIVE); ++it) {
                const Hit h = trace_closest<false>(tr, S.ray);
                pr_resolve_pending(S, occluded);
                Ray shadow;
                if (pr_stage(s->sc, S, h, max_depth, rr_depth, shadow)) occluded = trace_any<false>(tr, shadow);
            }
            pr_resolve_pending(S, occluded);
            L = S.L;
        } else { bool valid; L = path_primal<false>(s->sc, tr, c.ray, rng, max_depth, rr_depth, valid); }
        tmp[5 * i] = L.x; tmp[5 * i + 1] = L.y; tmp[5 * i + 2] = L.z; tmp[5 * i + 3] = c.pos_x; tmp[5 * i + 4] = c.pos_y;
    }
    for (int64_t i = 0; i < n; ++i) {
        const float pos_x = tmp[5 * i + 3], pos_y = tmp[5 * i + 4];
        const int pix = (int) floorf(pos_x) - 2, piy = (int) floorf(pos_y) - 2;
        float rel_x = (float) pix + .5f - pos_x, rel_y = (float) piy + .5f - pos_y, wx[5], wy[5];
        for (int k = 0; k < 5; ++k) { wx[k] = filter_eval(s->sc.filter, rel_x); wy[k] = filter_eval(s->sc.filter, rel_y); rel_x += 1; rel_y += 1; }
        for (int ys = 0; ys < 5; ++ys) for (int xs = 0; xs < 5; ++xs) {
            int x = pix + xs, y = piy + ys; if (x < 0 || y < 0 || x >= (int) FW || y >= (int) FH) continue;
            float w = wy[ys] * wx[xs]; float *d = film + ((size_t) y * FW + x) * 4;
            d[0] += tmp[5 * i] * w; d[1] += tmp[5 * i + 1] * w; d[2] += tmp[5 * i + 2] * w; d[3] += 1.0f * w;
        }
    }
}

// k_backward on the host.  records != NULL: per-lane records (EPSM_REC_FLOATS each); else scatter into grad_pos / grad_nrm.
void sim_backward(sim_scene *s, int variant, uint32_t sensor, uint32_t seed, uint32_t spp, uint32_t max_depth, uint32_t rr_depth,
                  const float *grad_in, uint64_t lane_begin, uint64_t lane_end, float *grad_pos, float *grad_nrm, float *records, int prune) {
    DLaunch ln = make_launch(s, sensor, seed, spp, max_depth, rr_depth, variant);
    const uint32_t W = ln.sensor.w;
    const uint32_t max_iter = max_depth < EPSM_MAX_ITER ? max_depth : EPSM_MAX_ITER;
    const int n_logged = (int) (max_iter < EPSM_MAXV ? max_iter : EPSM_MAXV);
    const int64_t n = (int64_t) (lane_end - lane_begin);
    std::vector<PathLog> logs(records ? 0 : n); std::vector<PathGrads> grads(records ? 0 : n);
    std::vector<uint8_t> mode(n, 0), fast_sel(n, 0); std::vector<int> navail(n, EPSM_MAXV);
    std::vector<f3> fast_dg(n); std::vector<uint32_t> fast_fi(3 * (size_t) n); std::vector<float> fast_bw(3 * (size_t) n);
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; ++i) {
        uint64_t lane = lane_begin + i; uint32_t pixel = (uint32_t) (lane / spp);
        LocalStack ls; TStack st; bind(st, ls); Tracer tr(s->sc, st);
        const int fx = (int) (pixel % W) - ln.sensor.border, fy = (int) (pixel / W) - ln.sensor.border;
        const bool in_film = fx >= 0 && fy >= 0 && fx < (int) ln.sensor.fw && fy < (int) ln.sensor.fh;
        const size_t fpix = in_film ? (size_t) fy * ln.sensor.fw + fx : 0;
        const float gx = in_film ? grad_in[fpix * 5 + 3] : 0.f, gy = in_film ? grad_in[fpix * 5 + 4] : 0.f;
        PathLog P;
        if (prune == 2 && !records) {
            // the production schedule of the gradient pass: first-vertex fast path (k_wf_stage_a) and a log that holds
            // only the gradient-carrying prefix (k_wf_finish)
            WfState S; wf_begin(ln.sensor, pixel % W, pixel / W, seed, (uint32_t) lane, gx, gy, S, P);
            const uint32_t iters = wf_replay_iterations(max_depth, true);
            int n_written = 0; bool scattered = false;
            for (uint32_t it = 0; it < iters && S.active; ++it) {
                Hit h = trace_closest<false>(tr, S.ray);
                SI si; bool occluded = false;
                const bool need_shadow = wf_stage_a(s->sc, S, P, h, it, max_depth, variant, true, si);
                if (wf_first_vertex_final(S, it, true)) {
                    f3 dg; fast_sel[i] = wf_first_vertex_seed(s->sc, S, P, si, dg) ? 1 : 0; fast_dg[i] = dg;
                    for (int c = 0; c < 3; ++c) { fast_fi[3 * i + c] = si.fi[c]; fast_bw[3 * i + c] = si.bw[c]; }
                    scattered = true; break;
                }
                if (need_shadow) occluded = trace_any<false>(tr, S.shadow_ray);
                Ray direct; bool dv; VLog scratch;
                VLog &v = it < EPSM_MAXV ? P.v[it] : scratch;
                wf_stage_b(s->sc, S, v, h, occluded, it, rr_depth, true, direct, dv);
                if (dv) wf_direct_hit(s->sc, v, trace_closest<false>(tr, direct));
                if (it < EPSM_MAXV) n_written = (int) it + 1;
            }
            mode[i] = scattered ? 1 : 0;
            if (scattered) continue;
            // k_wf_finish: keep the chain prefix, mark the vertex behind it inactive, poison the rest
            int n_avail = 0, n_diffuse = 0; bool chain = true;
            for (int k = 0; k < EPSM_MAXV; ++k) {
                if (k < n_logged && k < n_written && chain) {
                    const VLog &v = P.v[k]; const bool diffuse = (v.flags & EPSM_BF_DIFFUSE) != 0; n_diffuse += diffuse ? 1 : 0;
                    if (!v.active || !v.ismesh) chain = false;
                    if (variant == EPSM_VARIANT_MANIFOLD) { if (diffuse) chain = false; } else { if ((k == 0 && !diffuse) || n_diffuse >= 2) chain = false; }
                    n_avail = k + 1;
                } else if (k == n_avail) {
                    VLog &v = P.v[k]; memset(&v, 0xff, sizeof(VLog));
                    v.active = v.active_em = v.direct_valid = v.direct_any = v.direct_grad = v.direct_traced = false;
                    v.flags = v.ismesh = v.mesh_flags = v.has_normals = 0; v.prim = 0xffffffffu; v.p0 = v.p1 = v.p2 = mk3(0, 0, 0); v.b0 = v.b1 = 0.f;
                } else memset(&P.v[k], 0xff, sizeof(VLog));          // never read
            }
            navail[i] = n_avail;
            PathGrads G; memset(&G, 0xff, sizeof(G)); calc_grad(P, variant, n_logged, G, n_avail);
            logs[i] = P; grads[i] = G;
            continue;
        }
        path_log<false>(s->sc, tr, ln.sensor, pixel % W, pixel / W, seed, (uint32_t) lane, gx, gy, max_depth, rr_depth, variant, prune != 0 && !records, P);
        PathGrads G; calc_grad(P, variant, n_logged, G);
        if (records) {
            float *r = records + (size_t) i * EPSM_REC_FLOATS;
            for (int q = 0; q < EPSM_REC_FLOATS; ++q) r[q] = 0.f;
            for (int k = 0; k < n_logged; ++k) {
                const VLog &v = P.v[k]; float *o = r + k * EPSM_REC_VSTRIDE;
                o[0] = v.active ? 1.f : 0.f; o[1] = u2f(v.prim); o[2] = v.p.x; o[3] = v.p.y; o[4] = v.p.z; o[5] = v.b0; o[6] = v.b1;
                o[7] = v.active_em ? 1.f : 0.f; o[8] = v.light.x; o[9] = v.light.y; o[10] = v.light.z; o[11] = u2f(v.flags); o[12] = v.eta;
                o[13] = v.lr_sum; o[14] = v.dis; o[15] = u2f((v.direct_any ? 1u : 0u) | (v.direct_valid ? 2u : 0u) | (v.direct_grad ? 4u : 0u));
                float *g = r + EPSM_REC_GRAD_OFF + k * EPSM_REC_GSTRIDE;
                for (int cc = 0; cc < 5; ++cc) { g[3 * cc] = G.pg[k][cc].x; g[3 * cc + 1] = G.pg[k][cc].y; g[3 * cc + 2] = G.pg[k][cc].z; }
                g[15] = G.lg[k].x; g[16] = G.lg[k].y; g[17] = G.lg[k].z; g[18] = G.dg[k].x; g[19] = G.dg[k].y; g[20] = G.dg[k].z;
            }
            float *h = r + EPSM_REC_HEAD_OFF;
            h[0] = P.cam.x; h[1] = P.cam.y; h[2] = P.cam.z; h[3] = P.d.x; h[4] = P.d.y; h[5] = P.d.z;
            h[6] = P.dlduv0; h[7] = P.dlduv1; h[8] = P.dldp1.x; h[9] = P.dldp1.y; h[10] = P.dldp1.z; h[11] = P.L.x; h[12] = P.L.y; h[13] = P.L.z;
        } else { logs[i] = P; grads[i] = G; }
    }
    if (records) return;
    // scatter_path of epsm_kernels.cu with plain adds (same seeds, same masks)
    const int n_pg = variant == EPSM_VARIANT_MANIFOLD ? 5 * n_logged : 5 * (n_logged - 1) + 3;
    const bool shadow_term = max_iter <= 3;
    for (int64_t i = 0; i < n; ++i) {
        const PathLog &P = logs[i]; const PathGrads &G = grads[i];
        if (mode[i] == 1) {                                   // seed (ii) of the first vertex, scattered by k_wf_stage_a
            for (int c = 0; c < 3; ++c) add3(grad_pos, fast_fi[3 * i + c], fast_dg[i] * fast_bw[3 * i + c], fast_sel[i] != 0);
            continue;
        }
        for (int k = 0; k < n_logged && k < navail[i]; ++k) {
            const VLog &v = P.v[k];
            const bool sel = v.active && (v.mesh_flags & EPSM_MESH_GRAD);
            if (5 * k + 4 < n_pg) {
                for (int c = 0; c < 3; ++c) add3(grad_pos, v.fi[c], G.pg[k][c], sel);
                const f3 g3 = G.pg[k][3];
                const bool any_g3 = sel && (g3.x != 0.f || g3.y != 0.f || g3.z != 0.f);
                const float flip = (v.mesh_flags & EPSM_MESH_FLIP_NORMALS) ? -1.f : 1.f;
                if (any_g3 && v.has_normals) {
                    const f3 nt = v.n0 * v.bw[0] + v.n1 * v.bw[1] + v.n2 * v.bw[2];
                    const float il = 1.0f / len3(nt); const f3 nh = nt * il;
                    const f3 gt = (g3 - nh * dot3(nh, g3)) * (il * flip);
                    for (int c = 0; c < 3; ++c) add3(grad_nrm, v.fi[c], gt * v.bw[c], true);
                } else if (any_g3 && v.ismesh) {
                    const f3 e1 = v.p1 - v.p0, e2 = v.p2 - v.p0, cr = cross3(e1, e2);
                    const float il = 1.0f / len3(cr); const f3 nh = cr * il;
                    const f3 gc = (g3 - nh * dot3(nh, g3)) * (il * flip);
                    const f3 ge1 = cross3(e2, gc), ge2 = cross3(gc, e1);
                    add3(grad_pos, v.fi[1], ge1, true); add3(grad_pos, v.fi[2], ge2, true); add3(grad_pos, v.fi[0], -(ge1 + ge2), true);
                }
            }
            for (int c = 0; c < 3; ++c) add3(grad_pos, v.fi[c], G.dg[k] * v.bw[c], sel);
            if (k == 0 && shadow_term) { const f3 g = G.dg[0] * v.dis; for (int c = 0; c < 3; ++c) add3(grad_pos, v.dfi[c], g * v.dbw[c], v.direct_any && v.direct_grad); }
            const f3 gl = G.lg[k] * v.lr_sum;
            for (int c = 0; c < 3; ++c) add3(grad_pos, v.dfi[c], gl * v.dbw[c], v.direct_valid && v.direct_grad);
        }
    }
}

// k_prb_backward on the host: one lane at a time (serial, plain adds).  film = RGBW film of the primal pass over all lanes.
struct HostSink { float *gp, *gn; EPSM_HD void pos(uint32_t v, f3 g) { gp[3 * v] += g.x; gp[3 * v + 1] += g.y; gp[3 * v + 2] += g.z; }
                  EPSM_HD void nrm(uint32_t v, f3 g) { gn[3 * v] += g.x; gn[3 * v + 1] += g.y; gn[3 * v + 2] += g.z; } };
// The replay-wavefront policy of k_prb_pass (WfIo, epsm_kernels.cu) emulated per lane on the host: pass p queues the rays of round p,
// they are traced after the pass, and later passes read them back — the same slots, the same pass arithmetic.
struct HostWfIo {
    static constexpr bool wavefront = true;
    int pass; uint32_t K; std::vector<Hit> &hits; std::vector<uint32_t> &occl;
    std::vector<std::pair<Ray, uint32_t>> &qc, &qa; f3 L; bool done; PrbCheckpoint *saved;
    Hit closest(int round, uint32_t slot, const Ray &r) { if (round == pass) { qc.push_back({r, slot}); return Hit{0.f, 0.f, 0.f, 0xffffffffu}; } return hits[slot]; }
    bool occluded(int round, uint32_t slot, const Ray &r) { if (round == pass) { qa.push_back({r, slot}); return false; } return occl[slot] != 0u; }
    bool stop_after(int round) const { return pass == round; }
    bool adjoint_now(int round) const { return pass == round + 1; }
    bool aux_push_only(int round) const { return pass == round; }
    bool aux_skip(int round) const { return pass > round + 1; }
    uint32_t aux_reserve(uint32_t) { return 0u; }
    void aux_push(uint32_t j, uint32_t, const Ray &r) { qc.push_back({r, K + 1 + j}); }
    Hit aux_hit(uint32_t j, const Ray &) { return hits[K + 1 + j]; }
    bool lane_radiance(f3 &out) const { out = L; return true; }
    void finished() { done = true; }
    bool resume(PrbCheckpoint &c) const { if (pass < 2) return false; c = *saved; return true; }
    void checkpoint(uint32_t it, const PrbCheckpoint &c) const { if (pass == 2 * (int) it + 1) *saved = c; }
};
f3 path_primal_io(const DScene &, HostWfIo &, LocalTrace &, Ray, Pcg &, uint32_t, uint32_t, bool &) { return mk3(0, 0, 0); }

void sim_prb_backward(sim_scene *s, uint32_t sensor, uint32_t seed, uint32_t spp, const epsm_reparam_config *cfg, const float *film,
                      const float *grad_in, uint64_t lane_begin, uint64_t lane_end, float *grad_pos, float *grad_nrm, int wavefront) {
    const DSensor &se = s->hs.sensors[sensor];
    PrbConfig pc{cfg->max_depth, cfg->rr_depth, cfg->reparam_max_depth, cfg->reparam_rays, cfg->reparam_kappa, cfg->reparam_exp};
    HostSink sink{grad_pos, grad_nrm};
    for (uint64_t lane = lane_begin; lane < lane_end; ++lane) {
        LocalStack ls; TStack st; bind(st, ls); Tracer tr(s->sc, st);
        const uint32_t pixel = (uint32_t) (lane / spp);
        if (!wavefront) {
            MegaIo<false> io(tr);
            prb_backward_lane(s->sc, io, se, pc, pixel % se.w, pixel / se.w, seed, (uint32_t) lane, film, grad_in, sink);
            continue;
        }
        const uint32_t K = pc.max_depth;
        std::vector<Hit> hits(K + 1 + 2 * pc.n_rays, Hit{0.f, 0.f, 0.f, 0xffffffffu}); std::vector<uint32_t> occl(2 * K, 0u);
        std::vector<std::pair<Ray, uint32_t>> qc, qa;
        Pcg rng = pcg_for_lane(seed, (uint32_t) lane); CamRay cr; camera_ray(se, pixel % se.w, pixel / se.w, rng, cr);
        bool valid; const f3 L = path_primal<false>(s->sc, tr, cr.ray, rng, pc.max_depth, pc.rr_depth, valid);     // what k_pr_film writes to lane_L
        PrbCheckpoint saved;
        for (int pass = 0; pass <= 2 * (int) K + 1; ++pass) {
            HostWfIo io{pass, K, hits, occl, qc, qa, L, false, &saved};
            prb_backward_lane(s->sc, io, se, pc, pixel % se.w, pixel / se.w, seed, (uint32_t) lane, film, grad_in, sink);
            for (auto &q : qc) hits[q.second] = trace_closest<false>(tr, q.first);
            for (auto &q : qa) occl[q.second] = trace_any<false>(tr, q.first) ? 1u : 0u;
            qc.clear(); qa.clear();
            if (io.done) break;
        }
    }
}

} // extern "C"
