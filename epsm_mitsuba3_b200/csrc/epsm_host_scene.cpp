// epsm_host_scene.cpp — see epsm_host_scene.h.
#include "epsm_host_scene.h"
#include <math.h>
#include <string.h>

namespace {

// BSDF::flags() incl. the twosided wrapper (include/mitsuba/render/bsdf.h:37-110; src/bsdfs/twosided.cpp:80-100)
uint32_t bsdf_flags_word(const epsm_bsdf_desc &b) {
    const uint32_t Front = 0x8000u, Back = 0x10000u, NonSym = 0x4000u, Aniso = 0x1000u;
    uint32_t f = 0;
    switch (b.kind) {
        case EPSM_BSDF_DIFFUSE: f = EPSM_BF_DIFFUSE_REFLECTION | Front; break;
        case EPSM_BSDF_DIELECTRIC: f = EPSM_BF_DELTA_REFLECTION | EPSM_BF_DELTA_TRANSMISSION | Front | Back | NonSym; break;
        case EPSM_BSDF_CONDUCTOR: f = EPSM_BF_DELTA_REFLECTION | Front; break;
        case EPSM_BSDF_ROUGHCONDUCTOR: f = EPSM_BF_GLOSSY_REFLECTION | Front | (b.alpha_u != b.alpha_v ? Aniso : 0u); break;
    }
    if (b.flags & EPSM_BSDF_TWOSIDED) f |= Back;
    return f;
}

void mat4_mul(const double *a, const double *b, double *o) {
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += a[4 * i + k] * b[4 * k + j]; o[4 * i + j] = s; }
}
f3 project(const float *m, float x, float y, float z) {
    float X = m[0] * x + m[1] * y + m[2] * z + m[3], Y = m[4] * x + m[5] * y + m[6] * z + m[7],
          Z = m[8] * x + m[9] * y + m[10] * z + m[11], Wv = m[12] * x + m[13] * y + m[14] * z + m[15];
    float iw = 1.0f / Wv; return f3{X * iw, Y * iw, Z * iw};
}
// ProjectiveCamera transforms: perspective_projection (include/mitsuba/render/sensor.h:227-262) inverted analytically,
// position differentials as in PerspectiveCamera::update_camera_transforms (src/sensors/perspective.cpp:173-186)
void make_sensor(const epsm_sensor_desc &d, DSensor &s) {
    s.fw = d.width; s.fh = d.height; s.near_clip = d.near_clip; s.far_clip = d.far_clip;
    s.border = (d.flags & EPSM_SENSOR_SAMPLE_BORDER) ? 2 : 0;        // ReconstructionFilter::border_size = ceil(radius − ½) (rfilter.h), radius 2
    s.w = s.fw + 2 * (uint32_t) s.border; s.h = s.fh + 2 * (uint32_t) s.border;
    memcpy(s.to_world, d.to_world, sizeof(s.to_world));
    const double aspect = (double) d.width / (double) d.height, n = d.near_clip, f = d.far_clip;
    const double recip = 1.0 / (f - n), cot = 1.0 / tan(0.5 * (double) d.fov_x * M_PI / 180.0);
    const double A = f * recip, B = -n * f * recip;
    const double Pinv[16] = {1 / cot, 0, 0, 0, 0, 1 / cot, 0, 0, 0, 0, 0, 1, 0, 0, 1 / B, -A / B};
    const double Tm[16] = {1, 0, 0, 1, 0, 1, 0, 1 / aspect, 0, 0, 1, 0, 0, 0, 0, 1};
    const double Sm[16] = {-2, 0, 0, 0, 0, -2 / aspect, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    double t1[16], m[16]; mat4_mul(Tm, Sm, t1); mat4_mul(Pinv, t1, m);
    for (int i = 0; i < 16; ++i) s.s2c[i] = (float) m[i];
    {   // camera_to_sample = scale(−½, −½·aspect, 1) · translate(−1, −1/aspect, 0) · perspective (perspective.cpp:173-180)
        const double P[16] = {cot, 0, 0, 0, 0, cot, 0, 0, 0, 0, A, B, 0, 0, 1, 0};
        const double Tr[16] = {1, 0, 0, -1, 0, 1, 0, -1 / aspect, 0, 0, 1, 0, 0, 0, 0, 1};
        const double Sc[16] = {-0.5, 0, 0, 0, 0, -0.5 * aspect, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        double t2[16], c2s[16]; mat4_mul(Tr, P, t2); mat4_mul(Sc, t2, c2s);
        for (int i = 0; i < 16; ++i) s.c2s[i] = (float) c2s[i];
        // to_world⁻¹ for an affine matrix: [R t]⁻¹ = [R⁻¹ −R⁻¹t]
        double R[9]; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[3 * i + j] = d.to_world[4 * i + j];
        const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
        const double id = 1.0 / det;
        const double Ri[9] = {(R[4] * R[8] - R[5] * R[7]) * id, (R[2] * R[7] - R[1] * R[8]) * id, (R[1] * R[5] - R[2] * R[4]) * id,
                              (R[5] * R[6] - R[3] * R[8]) * id, (R[0] * R[8] - R[2] * R[6]) * id, (R[2] * R[3] - R[0] * R[5]) * id,
                              (R[3] * R[7] - R[4] * R[6]) * id, (R[1] * R[6] - R[0] * R[7]) * id, (R[0] * R[4] - R[1] * R[3]) * id};
        for (int i = 0; i < 3; ++i) {
            double tt = 0;
            for (int j = 0; j < 3; ++j) { s.w2c[4 * i + j] = (float) Ri[3 * i + j]; tt += Ri[3 * i + j] * d.to_world[4 * j + 3]; }
            s.w2c[4 * i + 3] = (float) -tt;
        }
        s.w2c[12] = s.w2c[13] = s.w2c[14] = 0.f; s.w2c[15] = 1.f;
    }
    f3 p0 = project(s.s2c, 0, 0, 0), px = project(s.s2c, 1.0f / d.width, 0, 0), py = project(s.s2c, 0, 1.0f / d.height, 0);
    s.dx = f3{px.x - p0.x, px.y - p0.y, px.z - p0.z}; s.dy = f3{py.x - p0.x, py.y - p0.y, py.z - p0.z};
}
// GaussianFilter (src/rfilters/gaussian.cpp:48-96), stddev 0.5, radius 2: scaled Remez coefficients
void make_filter(float *c) {
    const double coeff[10] = {9.992604880e-1, -4.977025247e-1, 1.222248550e-1, -1.932406282e-2, 2.136713061e-3,
                              -1.679873860e-4, 9.202145248e-6, -3.329417433e-7, 7.128382794e-9, -6.821193280e-11};
    const float stddev = 0.5f, radius = 2.0f;
    double scale = 1;
    for (int i = 0; i < 10; ++i) { c[i] = (float) (coeff[i] * scale); scale /= (double) stddev * (double) stddev; }
    // coeff_s[0] -= estrin(radius²)
    float x = radius * radius, x2 = x * x, x4 = x2 * x2, x8 = x4 * x4;
    float r0 = fmaf(x, c[1], c[0]), r1 = fmaf(x, c[3], c[2]), r2 = fmaf(x, c[5], c[4]), r3 = fmaf(x, c[7], c[6]), r4 = fmaf(x, c[9], c[8]);
    float q0 = fmaf(x2, r1, r0), q1 = fmaf(x2, r3, r2), p0 = fmaf(x4, q1, q0);
    c[0] -= fmaf(x8, r4, p0);
}

// Mesh::build_pmf (src/render/mesh.cpp:361-393) + DiscreteDistribution::compute_cdf (distr_1d.h:212-239)
void build_area_tables(HostScene *c) {
    c->cdf.clear(); c->pmf.clear();
    for (auto &m : c->meshes) {
        m.cdf_off = (uint32_t) c->cdf.size(); m.valid_lo = 0xffffffffu; m.valid_hi = 0;
        double sum = 0;
        const bool need_table = m.emitter >= 0 && !(m.flags & EPSM_MESH_RECTANGLE);
        for (uint32_t i = 0; i < m.f_cnt; ++i) {
            const uint32_t *f = &c->faces[3 * (size_t) (m.f_off + i)];
            const float *p0 = &c->pos[3 * (size_t) f[0]], *p1 = &c->pos[3 * (size_t) f[1]], *p2 = &c->pos[3 * (size_t) f[2]];
            const float ax = p1[0] - p0[0], ay = p1[1] - p0[1], az = p1[2] - p0[2], bx = p2[0] - p0[0], by = p2[1] - p0[1], bz = p2[2] - p0[2];
            const float cx = fmaf(ay, bz, -(az * by)), cy = fmaf(az, bx, -(ax * bz)), cz = fmaf(ax, by, -(ay * bx));
            const float area = 0.5f * sqrtf(fmaf(cz, cz, fmaf(cy, cy, cx * cx)));
            sum += (double) area;
            if (need_table) { c->pmf.push_back(area); c->cdf.push_back((float) sum); }
            if (area > 0) { if (m.valid_lo == 0xffffffffu) m.valid_lo = i; m.valid_hi = i; }
        }
        m.area_sum = (float) sum; m.area_norm = (float) (1.0 / sum);
        if ((m.flags & EPSM_MESH_RECTANGLE) && m.v_cnt >= 4) {                      // rectangle.cpp:107-108,145
            const float *c0 = &c->pos[3 * (size_t) m.v_off], *c1 = c0 + 3, *c3 = c0 + 9;
            const float ax = c1[0] - c0[0], ay = c1[1] - c0[1], az = c1[2] - c0[2], bx = c3[0] - c0[0], by = c3[1] - c0[1], bz = c3[2] - c0[2];
            const float cx = fmaf(ay, bz, -(az * by)), cy = fmaf(az, bx, -(ax * bz)), cz = fmaf(ax, by, -(ay * bx));
            m.area_norm = 1.0f / sqrtf(fmaf(cz, cz, fmaf(cy, cy, cx * cx)));
        }
    }
}


} // namespace

std::string host_scene_build(HostScene &hs, const epsm_mesh_desc *meshes, uint32_t n_meshes, const epsm_bsdf_desc *bsdfs, uint32_t n_bsdfs,
                             const epsm_emitter_desc *emitters, uint32_t n_emitters, const epsm_sensor_desc *sensors, uint32_t n_sensors) {
    if (!meshes || n_meshes == 0 || !bsdfs || n_bsdfs == 0 || !sensors || n_sensors == 0)
        return "scene needs at least one mesh, bsdf and sensor";
    hs.pos.clear(); hs.nrm.clear(); hs.faces.clear(); hs.face_mesh.clear(); hs.uv.clear(); hs.texels.clear(); hs.textures.clear();
    hs.meshes.clear(); hs.bsdfs.clear(); hs.emitters.clear(); hs.sensors.clear();
    for (uint32_t i = 0; i < n_meshes; ++i) {
        const epsm_mesh_desc &d = meshes[i];
        if (!d.positions || !d.faces || d.vertex_count == 0 || d.face_count == 0) return "empty mesh";
        if (d.bsdf_id < 0 || (uint32_t) d.bsdf_id >= n_bsdfs) return "mesh bsdf_id out of range";
        if (d.emitter_id >= (int32_t) n_emitters) return "mesh emitter_id out of range";
        DMesh m; memset(&m, 0, sizeof(m));
        m.v_off = (uint32_t) (hs.pos.size() / 3); m.f_off = (uint32_t) (hs.faces.size() / 3);
        m.v_cnt = d.vertex_count; m.f_cnt = d.face_count; m.bsdf = d.bsdf_id; m.emitter = d.emitter_id; m.flags = d.flags;
        m.has_normals = d.normals ? 1u : 0u; m.has_uv = d.texcoords ? 1u : 0u;
        if (d.texcoords) hs.uv.insert(hs.uv.end(), d.texcoords, d.texcoords + 2 * (size_t) d.vertex_count);
        else hs.uv.insert(hs.uv.end(), 2 * (size_t) d.vertex_count, 0.f);
        hs.pos.insert(hs.pos.end(), d.positions, d.positions + 3 * (size_t) d.vertex_count);
        if (d.normals) hs.nrm.insert(hs.nrm.end(), d.normals, d.normals + 3 * (size_t) d.vertex_count);
        else hs.nrm.insert(hs.nrm.end(), 3 * (size_t) d.vertex_count, 0.f);
        for (size_t f = 0; f < 3 * (size_t) d.face_count; ++f) {
            if (d.faces[f] >= d.vertex_count) return "face index out of range";
            hs.faces.push_back(d.faces[f] + m.v_off);
        }
        hs.face_mesh.insert(hs.face_mesh.end(), d.face_count, i);
        hs.meshes.push_back(m);
    }
    for (uint32_t i = 0; i < n_bsdfs; ++i) {
        const epsm_bsdf_desc &b = bsdfs[i];
        if (b.kind > EPSM_BSDF_ROUGHCONDUCTOR) return "unknown bsdf kind";
        DBsdf d; memset(&d, 0, sizeof(d));
        d.kind = b.kind; d.flags = b.flags; d.bflags = bsdf_flags_word(b);
        memcpy(d.reflectance, b.reflectance, 12); memcpy(d.transmittance, b.transmittance, 12); memcpy(d.eta, b.eta, 12); memcpy(d.k, b.k, 12);
        d.alpha_u = b.alpha_u; d.alpha_v = b.alpha_v;
        d.tex = -1;
        if (b.reflectance_texture) {
            const epsm_texture_desc &t = *b.reflectance_texture;
            if (b.kind != EPSM_BSDF_DIFFUSE) return "reflectance_texture is supported on diffuse BSDFs only";
            if (!t.texels || t.width == 0 || t.height == 0 || t.filter > EPSM_TEX_BILINEAR || t.wrap > EPSM_TEX_MIRROR) return "bad texture descriptor";
            DTexture dt; memset(&dt, 0, sizeof(dt));
            dt.off = (uint32_t) hs.texels.size(); dt.w = t.width; dt.h = t.height; dt.filter = t.filter; dt.wrap = t.wrap; memcpy(dt.to_uv, t.to_uv, sizeof(dt.to_uv));
            hs.texels.insert(hs.texels.end(), t.texels, t.texels + 3 * (size_t) t.width * t.height);
            d.tex = (int32_t) hs.textures.size(); hs.textures.push_back(dt);
        }
        hs.bsdfs.push_back(d);
    }
    for (uint32_t i = 0; i < n_emitters; ++i) {
        if (emitters[i].mesh_id < 0 || (uint32_t) emitters[i].mesh_id >= n_meshes) return "emitter mesh_id out of range";
        DEmitter e; memcpy(e.radiance, emitters[i].radiance, 12); e.mesh = emitters[i].mesh_id;
        hs.emitters.push_back(e);
    }
    for (uint32_t i = 0; i < n_sensors; ++i) {
        if (sensors[i].width == 0 || sensors[i].height == 0) return "empty film";
        DSensor s; make_sensor(sensors[i], s); hs.sensors.push_back(s);
    }
    make_filter(hs.filter);
    host_scene_rebuild_accel(hs);
    return std::string();
}

void host_scene_rebuild_accel(HostScene &hs) {
    epsm_build_bvh4(hs.pos.data(), hs.faces.data(), hs.faces.size() / 3, hs.bvh);
    build_area_tables(&hs);
}

void host_scene_rebuild_area_tables(HostScene &hs) { build_area_tables(&hs); }

std::string host_scene_set_sensor(HostScene &hs, uint32_t index, const epsm_sensor_desc &d) {
    if (index >= hs.sensors.size()) return "sensor index out of range";
    if (d.width != hs.sensors[index].fw || d.height != hs.sensors[index].fh) return "a sensor update cannot change the film size";
    if (!(d.fov_x > 0.f && d.fov_x < 180.f) || !(d.far_clip > d.near_clip)) return "bad sensor parameters";
    make_sensor(d, hs.sensors[index]);
    return "";
}

void host_scene_view(const HostScene &hs, DScene &sc) {
    sc.nodes = hs.bvh.nodes.data(); sc.tris = hs.bvh.tris.data(); sc.pos = hs.pos.data(); sc.nrm = hs.nrm.data();
    sc.faces = hs.faces.data(); sc.face_mesh = hs.face_mesh.data(); sc.meshes = hs.meshes.data(); sc.bsdfs = hs.bsdfs.data();
    sc.emitters = hs.emitters.data(); sc.cdf = hs.cdf.data(); sc.pmf = hs.pmf.data();
    sc.uv = hs.uv.data(); sc.textures = hs.textures.data(); sc.texels = hs.texels.data();
    sc.n_emitters = (uint32_t) hs.emitters.size(); sc.n_meshes = (uint32_t) hs.meshes.size();
    sc.n_tris = (uint32_t) (hs.faces.size() / 3); sc.n_nodes = (uint32_t) (hs.bvh.nodes.size() / EPSM_NODE_F4);
    memcpy(sc.filter, hs.filter, sizeof(sc.filter));
}
