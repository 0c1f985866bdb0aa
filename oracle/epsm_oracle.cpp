// epsm_oracle.cpp — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's algorithm for
// the EPSM manifold-integrator path (jkxing/EPSM_Mitsuba3), scalar C++, brute-force intersection.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// this library, and only as the checker or the reported CPU baseline.  The product
// (epsm_mitsuba3_b200/csrc) never links, imports or calls it.
//
// Parity pinning: TEA and PCG32 are pinned by the reference's own known-answer tests
// (src/core/tests/test_random.py:8-16, src/samplers/tests/test_independent.py:17-29) and the
// triangle/surface-interaction/BSDF/emitter/sensor functions by values read from the reference's
// scalar_rgb module, built unmodified by baseline/build_ref.sh (tests/golden/reference_kats.json and ≥ 1 k seeded vectors
// per function in tests/golden/reference_vectors.npz).  The manifold solve (oracle/calc_grad_ref.py) is pinned to the
// outputs of the reference's own calc_grad source executed on CPU torch (tests/golden/calc_grad_reference.npz).
// "Parity unpinned" remains for prb_reparam's reverse pass (orc_prb_log + oracle/prb_reparam_ref.py: no AD variant of the
// reference runs in this image; checked against finite differences like the reference's own tests) and for the matcher.
//
// Arithmetic conventions follow Dr.Jit's JIT array code so that the CUDA path can be bit-compared:
//   dot(a,b)   = fma(a.z,b.z, fma(a.y,b.y, a.x*b.x))          ext/drjit/include/drjit/array_base.h:603-628
//   cross(a,b) = fmsub(a.yzx, b.zxy, a.zxy*b.yzx)             ext/drjit/include/drjit/array_router.h:655-666
//   normalize  = v * rsqrt(squared_norm(v))                   array_router.h:635-652
// Compile with -ffp-contract=off (see oracle/Makefile): the only FMAs are the explicit fmaf() calls.
//
// All file:line citations are relative to the reference root.

#include <cmath>
#include <cstdint>
#include <cstring>
#include <cfloat>
#include <vector>
#include <algorithm>
#include "../include/epsm_b200.h"

namespace {

// ------------------------------------------------------------------------------------------------
//  Small vector algebra
// ------------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
inline V3 mk(float x, float y, float z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
inline float dot(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
inline V3 cross(V3 a, V3 b) {
    return mk(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
inline float sqnorm(V3 a) { return fmaf(a.z, a.z, fmaf(a.y, a.y, a.x * a.x)); }
inline float rsqrt_(float x) { return 1.0f / sqrtf(x); }
inline float rcp_(float x) { return 1.0f / x; }
inline V3 normalize(V3 a) { return a * rsqrt_(sqnorm(a)); }
inline float norm(V3 a) { return sqrtf(sqnorm(a)); }
inline V3 fmadd(V3 a, float b, V3 c) { return mk(fmaf(a.x, b, c.x), fmaf(a.y, b, c.y), fmaf(a.z, b, c.z)); }
inline float mulsign(float a, float b) { return std::signbit(b) ? -a : a; }
inline float mulsign_neg(float a, float b) { return std::signbit(b) ? a : -a; }
inline float safe_sqrt(float x) { return sqrtf(fmaxf(x, 0.0f)); }
inline float sqr(float x) { return x * x; }
inline float max3(V3 a) { return fmaxf(fmaxf(a.x, a.y), a.z); }

const float kPi = 3.14159265358979323846f, kInvPi = 0.31830988618379067154f;
const float kRayEps = 1500.0f * 5.9604644775390625e-8f;   // include/mitsuba/core/math.h:18-22 (Epsilon = 2^-24)
const float kShadowEps = kRayEps * 10.0f;

// ------------------------------------------------------------------------------------------------
//  Random numbers: TEA + PCG32 (integer, bit-exact)
// ------------------------------------------------------------------------------------------------
// include/mitsuba/core/random.h:76-90
void tea32(uint32_t v0, uint32_t v1, int rounds, uint32_t *o0, uint32_t *o1) {
    uint32_t sum = 0;
    for (int i = 0; i < rounds; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    *o0 = v0; *o1 = v1;
}

// ext/drjit/include/drjit/random.h:25-27,55-75,130-132
struct PCG32 {
    uint64_t state, inc;
    uint32_t next_u32() {
        uint64_t old = state;
        state = old * 0x5851f42d4c957f2dULL + inc;
        uint32_t xs = (uint32_t) (((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t) (old >> 59u);
        return (xs >> rot) | (xs << ((~rot + 1u) & 31u));
    }
    void seed(uint64_t initstate, uint64_t initseq) {
        state = 0; inc = (initseq << 1u) | 1u;
        next_u32(); state += initstate; next_u32();
    }
    float next_f32() {
        uint32_t u = (next_u32() >> 9) | 0x3f800000u;
        float f; memcpy(&f, &u, 4); return f - 1.0f;
    }
};

// src/render/sampler.cpp:115-130 (JIT branch): per-lane stream
PCG32 sampler_for_lane(uint32_t seed_value, uint32_t lane) {
    uint32_t v0, v1; tea32(seed_value, lane, 4, &v0, &v1);
    PCG32 r; r.seed(v0, v1); return r;
}

// ------------------------------------------------------------------------------------------------
//  Dr.Jit single-precision sincos (ext/drjit/include/drjit/math.h:76-186), restated
// ------------------------------------------------------------------------------------------------
void sincos_dr(float x, float *s_out, float *c_out) {
    float xa = fabsf(x);
    int32_t j = (int32_t) (xa * 1.2732395447351626862f);
    j = (j + 1) & ~1;
    float y = (float) j;
    uint32_t xb; memcpy(&xb, &x, 4);
    uint32_t sign_sin = ((uint32_t) j << 29) ^ xb;
    uint32_t sign_cos = (uint32_t) (~(j - 2)) << 29;
    y = xa - y * 0.78515625f - y * 2.4187564849853515625e-4f - y * 3.77489497744594108e-8f;
    float z = y * y;
    // estrin(z, c0, c1, c2) = fma(z*z, c2, fma(z, c1, c0))   (math.h:26-42)
    float s = fmaf(z * z, -1.9515295891e-4f, fmaf(z, 8.3321608736e-3f, -1.6666654611e-1f)) * z;
    float c = fmaf(z * z, 2.443315711809948e-5f, fmaf(z, -1.388731625493765e-3f, 4.166664568298827e-2f)) * z;
    s = fmaf(s, y, y);
    c = fmaf(c, z, fmaf(z, -0.5f, 1.0f));
    bool poly = (j & 2) == 0;
    float rs = poly ? s : c, rc = poly ? c : s;
    uint32_t b;
    memcpy(&b, &rs, 4); b ^= (sign_sin & 0x80000000u); memcpy(s_out, &b, 4);
    memcpy(&b, &rc, 4); b ^= (sign_cos & 0x80000000u); memcpy(c_out, &b, 4);
}

// ------------------------------------------------------------------------------------------------
//  Warps (include/mitsuba/core/warp.h)
// ------------------------------------------------------------------------------------------------
// warp.h:54-95
void square_to_uniform_disk_concentric(float sx, float sy, float *ox, float *oy) {
    float x = fmaf(2.0f, sx, -1.0f), y = fmaf(2.0f, sy, -1.0f);
    bool is_zero = (x == 0.0f) && (y == 0.0f), q13 = fabsf(x) < fabsf(y);
    float r = q13 ? y : x, rp = q13 ? x : y;
    float phi = 0.25f * kPi * rp / r;
    if (q13) phi = 0.5f * kPi - phi;
    if (is_zero) phi = 0.0f;
    float s, c; sincos_dr(phi, &s, &c);
    *ox = r * c; *oy = r * s;
}
// warp.h:320-328
V3 square_to_cosine_hemisphere(float sx, float sy) {
    float px, py; square_to_uniform_disk_concentric(sx, sy, &px, &py);
    float z = safe_sqrt(1.0f - fmaf(py, py, px * px));
    return mk(px, py, z);
}
// include/mitsuba/core/vector.h:116-136
void coordinate_system(V3 n, V3 *s, V3 *t) {
    float sign = std::signbit(n.z) ? -1.0f : 1.0f;
    float a = -rcp_(sign + n.z), b = n.x * n.y * a;
    *s = mk(mulsign(sqr(n.x) * a, n.z) + 1.0f, mulsign(b, n.z), mulsign_neg(n.x, n.z));
    *t = mk(b, fmaf(n.y, n.y * a, sign), -n.y);
}

// ------------------------------------------------------------------------------------------------
//  Scene
// ------------------------------------------------------------------------------------------------
struct Mesh {
    uint32_t v_off, f_off, v_cnt, f_cnt;
    int32_t bsdf, emitter; uint32_t flags; bool has_normals, has_uv;
    std::vector<float> cdf; float area_sum, area_norm; uint32_t valid_lo, valid_hi;   // distr_1d.h:212-239
    std::vector<float> pmf;
};
struct Sensor {
    float s2c[16], to_world[16]; V3 dx, dy; float near_clip, far_clip;
    uint32_t w, h;          // sampling grid: film + 2·border when the film has sample_border (common.py:318-324)
    uint32_t fw, fh;        // film / crop size
    int border;             // rfilter.border_size() = 2 for the Gaussian, 0 without sample_border
};
struct Texture { epsm_texture_desc d; std::vector<float> texels; };
struct Scene {
    std::vector<V3> pos, nrm;
    std::vector<float> uv;              // 2 per vertex (zeros for meshes without texture coordinates)
    std::vector<Texture> textures;      // owned copies; bsdfs[i].reflectance_texture points into them
    std::vector<uint32_t> faces;        // global vertex ids
    std::vector<uint32_t> face_mesh;
    std::vector<Mesh> meshes;
    std::vector<epsm_bsdf_desc> bsdfs;
    std::vector<epsm_emitter_desc> emitters;
    std::vector<Sensor> sensors;
    float filter_coeff[10];
};

// Reconstruction filter: src/rfilters/gaussian.cpp:48-96 (non-CUDA variants: Remez polynomial)
float estrin10(float x, const float *c) {               // math.h:26-42 unrolled for n = 10
    float x2 = x * x, x4 = x2 * x2, x8 = x4 * x4;
    float r0 = fmaf(x, c[1], c[0]), r1 = fmaf(x, c[3], c[2]), r2 = fmaf(x, c[5], c[4]),
          r3 = fmaf(x, c[7], c[6]), r4 = fmaf(x, c[9], c[8]);
    float q0 = fmaf(x2, r1, r0), q1 = fmaf(x2, r3, r2), q2 = r4;
    float p0 = fmaf(x4, q1, q0), p1 = q2;
    return fmaf(x8, p1, p0);
}
void init_filter(Scene &sc) {
    const double coeff[10] = { 9.992604880e-1, -4.977025247e-1, 1.222248550e-1, -1.932406282e-2, 2.136713061e-3,
                               -1.679873860e-4, 9.202145248e-6, -3.329417433e-7, 7.128382794e-9, -6.821193280e-11 };
    const float stddev = 0.5f, radius = 4 * stddev;
    double scale = 1;
    for (int i = 0; i < 10; ++i) { sc.filter_coeff[i] = (float) (coeff[i] * scale); scale /= (double) stddev * (double) stddev; }
    sc.filter_coeff[0] -= estrin10(radius * radius, sc.filter_coeff);
}
inline float filter_eval(const Scene &sc, float x) { return fmaxf(estrin10(x * x, sc.filter_coeff), 0.0f); }

// include/mitsuba/render/sensor.h:227-262 and src/sensors/perspective.cpp:173-199, built in double, stored as float
void mat4_mul(const double *a, const double *b, double *o) {
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
        double s = 0; for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j]; o[i * 4 + j] = s;
    }
}
V3 xform_point(const float *m, V3 p) {                  // projective: divide by w
    float x = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
    float y = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    float z = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
    float w = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    float iw = 1.0f / w; return mk(x * iw, y * iw, z * iw);
}
V3 xform_vec(const float *m, V3 v) {
    return mk(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z);
}
void init_sensor(Sensor &s, const epsm_sensor_desc &d) {
    s.fw = d.width; s.fh = d.height; s.near_clip = d.near_clip; s.far_clip = d.far_clip;
    s.border = (d.flags & EPSM_SENSOR_SAMPLE_BORDER) ? 2 : 0;
    s.w = s.fw + 2 * s.border; s.h = s.fh + 2 * s.border;
    memcpy(s.to_world, d.to_world, sizeof(s.to_world));
    double aspect = (double) d.width / (double) d.height;
    double recip = 1.0 / ((double) d.far_clip - (double) d.near_clip);
    double cot = 1.0 / std::tan((double) d.fov_x * 0.5 * M_PI / 180.0);
    double A = d.far_clip * recip, B = -(double) d.near_clip * d.far_clip * recip;
    // sample_to_camera = P^-1 · T(1, 1/aspect, 0) · S(-2, -2/aspect, 1)
    double Pinv[16] = { 1 / cot, 0, 0, 0,  0, 1 / cot, 0, 0,  0, 0, 0, 1,  0, 0, 1 / B, -A / B };
    double T[16] = { 1, 0, 0, 1,  0, 1, 0, 1 / aspect,  0, 0, 1, 0,  0, 0, 0, 1 };
    double S[16] = { -2, 0, 0, 0,  0, -2 / aspect, 0, 0,  0, 0, 1, 0,  0, 0, 0, 1 };
    double tmp[16], m[16]; mat4_mul(T, S, tmp); mat4_mul(Pinv, tmp, m);
    for (int i = 0; i < 16; ++i) s.s2c[i] = (float) m[i];
    V3 p0 = xform_point(s.s2c, mk(0, 0, 0));
    s.dx = xform_point(s.s2c, mk(1.0f / d.width, 0, 0)) - p0;
    s.dy = xform_point(s.s2c, mk(0, 1.0f / d.height, 0)) - p0;
}

// src/render/mesh.cpp:361-393 + include/mitsuba/core/distr_1d.h:212-239
void build_pmf(const Scene &sc, Mesh &m) {
    m.pmf.resize(m.f_cnt); m.cdf.resize(m.f_cnt);
    double sum = 0; m.valid_lo = (uint32_t) -1; m.valid_hi = 0;
    for (uint32_t i = 0; i < m.f_cnt; ++i) {
        const uint32_t *f = &sc.faces[3 * (m.f_off + i)];
        V3 p0 = sc.pos[f[0]], p1 = sc.pos[f[1]], p2 = sc.pos[f[2]];
        float a = 0.5f * norm(cross(p1 - p0, p2 - p0));
        m.pmf[i] = a; sum += (double) a; m.cdf[i] = (float) sum;
        if (a > 0) { if (m.valid_lo == (uint32_t) -1) m.valid_lo = i; m.valid_hi = i; }
    }
    m.area_sum = (float) sum; m.area_norm = (float) (1.0 / sum);
    if (m.flags & EPSM_MESH_RECTANGLE) {                    // rectangle.cpp:107-108,145: 1 / |dp_du × dp_dv|
        V3 c0 = sc.pos[m.v_off], ex = sc.pos[m.v_off + 1] - c0, ey = sc.pos[m.v_off + 3] - c0;
        m.area_norm = rcp_(norm(cross(ex, ey)));
    }
}

// ------------------------------------------------------------------------------------------------
//  Ray / triangle (include/mitsuba/render/mesh.h:342-367) and brute-force scene queries
// ------------------------------------------------------------------------------------------------
struct Ray { V3 o, d; float maxt; };
inline bool moeller_trumbore(const Ray &ray, V3 p0, V3 p1, V3 p2, float *t_, float *u_, float *v_) {
    V3 e1 = p1 - p0, e2 = p2 - p0;
    V3 pvec = cross(ray.d, e2);
    float inv_det = rcp_(dot(e1, pvec));
    V3 tvec = ray.o - p0;
    float u = dot(tvec, pvec) * inv_det;
    bool active = u >= 0.f && u <= 1.f;
    V3 qvec = cross(tvec, e1);
    float v = dot(ray.d, qvec) * inv_det;
    active &= v >= 0.f && u + v <= 1.f;
    float t = dot(e2, qvec) * inv_det;
    active &= t >= 0.f && t <= ray.maxt;
    *t_ = t; *u_ = u; *v_ = v;
    return active;
}
struct PI { bool valid; float t, u, v; uint32_t prim; };
// Closest hit, scanning primitives in index order with `t <= maxt` (mesh.h:363) and maxt shrinking on
// every accepted hit, as the scalar kd-tree leaf loop does (kdtree.h:2188-2290): later primitive wins ties.
PI intersect_closest(const Scene &sc, Ray ray) {
    PI pi{false, INFINITY, 0, 0, 0xffffffffu};
    size_t nf = sc.faces.size() / 3;
    for (size_t i = 0; i < nf; ++i) {
        const uint32_t *f = &sc.faces[3 * i];
        float t, u, v;
        if (moeller_trumbore(ray, sc.pos[f[0]], sc.pos[f[1]], sc.pos[f[2]], &t, &u, &v)) {
            pi.valid = true; pi.t = t; pi.u = u; pi.v = v; pi.prim = (uint32_t) i; ray.maxt = t;
        }
    }
    return pi;
}
bool intersect_any(const Scene &sc, const Ray &ray) {
    size_t nf = sc.faces.size() / 3;
    for (size_t i = 0; i < nf; ++i) {
        const uint32_t *f = &sc.faces[3 * i];
        float t, u, v;
        if (moeller_trumbore(ray, sc.pos[f[0]], sc.pos[f[1]], sc.pos[f[2]], &t, &u, &v)) return true;
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
//  Surface interaction: src/render/mesh.cpp:632-892 + include/mitsuba/render/interaction.h:262-275,482-505
// ------------------------------------------------------------------------------------------------
struct SI {
    bool valid; float t; V3 p, n, sh_n, sh_s, sh_t, wi;
    uint32_t prim, mesh, fi[3]; float bw[3];           // bw = (b0,b1,b2) = (1-u-v, u, v)
    // EPSM additions (interaction.h:221-224)
    float b0, b1; V3 p0, p1, p2, n0, n1, n2; uint32_t ismesh;
    float uv[2];                                       // si.uv (mesh.cpp:731-744)
    V3 dp_du, dp_dv;                                   // mesh.cpp:734,746-763
};
inline V3 to_local(const SI &si, V3 v) { return mk(dot(v, si.sh_s), dot(v, si.sh_t), dot(v, si.sh_n)); }
inline V3 to_world(const SI &si, V3 v) { return fmadd(si.sh_n, v.z, fmadd(si.sh_t, v.y, si.sh_s * v.x)); }   // frame.h:41

SI compute_si(const Scene &sc, const Ray &ray, const PI &pi) {
    SI si; memset(&si, 0, sizeof(si));
    si.valid = pi.valid; si.t = INFINITY; si.prim = 0xffffffffu;
    if (!pi.valid) { si.wi = -ray.d; return si; }
    uint32_t mid = sc.face_mesh[pi.prim]; const Mesh &m = sc.meshes[mid];
    const uint32_t *f = &sc.faces[3 * pi.prim];
    V3 p0 = sc.pos[f[0]], p1 = sc.pos[f[1]], p2 = sc.pos[f[2]];
    float b1 = pi.u, b2 = pi.v, b0 = 1.f - b1 - b2;
    V3 dp0 = p1 - p0, dp1 = p2 - p0;
    si.t = pi.t; si.prim = pi.prim; si.mesh = mid;
    si.fi[0] = f[0]; si.fi[1] = f[1]; si.fi[2] = f[2]; si.bw[0] = b0; si.bw[1] = b1; si.bw[2] = b2;
    si.p = fmadd(p0, b0, fmadd(p1, b1, p2 * b2));                                   // mesh.cpp:707
    si.n = normalize(cross(dp0, dp1));                                              // mesh.cpp:729
    V3 dp_du, dp_dv; coordinate_system(si.n, &dp_du, &dp_dv);                       // mesh.cpp:734
    si.uv[0] = b1; si.uv[1] = b2;                                                   // mesh.cpp:731
    if (m.has_uv) {                                                                 // mesh.cpp:736-763 (RayFlags::All)
        const float *t0 = &sc.uv[2 * f[0]], *t1 = &sc.uv[2 * f[1]], *t2 = &sc.uv[2 * f[2]];
        si.uv[0] = fmaf(t2[0], b2, fmaf(t1[0], b1, t0[0] * b0)); si.uv[1] = fmaf(t2[1], b2, fmaf(t1[1], b1, t0[1] * b0));
        const float du0 = t1[0] - t0[0], dv0 = t1[1] - t0[1], du1 = t2[0] - t0[0], dv1 = t2[1] - t0[1];
        const float det = fmaf(du0, dv1, -(dv0 * du1));                             // fmsub(duv0.x, duv1.y, duv0.y·duv1.x)
        if (det != 0.f) {
            const float inv_det = rcp_(det);
            // fmsub(duv1.y, dp0, duv0.y·dp1)·inv_det ; fnmadd(duv1.x, dp0, duv0.x·dp1)·inv_det
            dp_du = mk(fmaf(dv1, dp0.x, -(dv0 * dp1.x)), fmaf(dv1, dp0.y, -(dv0 * dp1.y)), fmaf(dv1, dp0.z, -(dv0 * dp1.z))) * inv_det;
            dp_dv = mk(fmaf(-du1, dp0.x, du0 * dp1.x), fmaf(-du1, dp0.y, du0 * dp1.y), fmaf(-du1, dp0.z, du0 * dp1.z)) * inv_det;
        }
    }
    V3 n0, n1, n2;
    if (m.has_normals) {
        n0 = sc.nrm[f[0]]; n1 = sc.nrm[f[1]]; n2 = sc.nrm[f[2]];
        V3 n = fmadd(n2, b2, fmadd(n1, b1, n0 * b0));                               // mesh.cpp:788
        float il = rsqrt_(sqnorm(n));
        si.sh_n = n * il;
    } else {
        si.sh_n = si.n; n0 = n1 = n2 = si.n;                                        // mesh.cpp:810-815
    }
    si.dp_du = dp_du; si.dp_dv = dp_dv;
    if (m.flags & EPSM_MESH_FLIP_NORMALS) { si.n = -si.n; si.sh_n = -si.sh_n; n0 = -n0; n1 = -n1; n2 = -n2; }
    // initialize_sh_frame, interaction.h:262-275
    si.sh_s = normalize(fmadd(si.sh_n, -dot(si.sh_n, dp_du), dp_du));
    si.sh_t = cross(si.sh_n, si.sh_s);
    si.wi = to_local(si, -ray.d);                                                   // interaction.h:499
    if (m.flags & EPSM_MESH_ISMESH) {                                               // mesh.cpp:712-720
        si.b0 = b0; si.b1 = b1; si.p0 = p0; si.p1 = p1; si.p2 = p2; si.n0 = n0; si.n1 = n1; si.n2 = n2; si.ismesh = 1;
    }
    return si;
}

// Dr.Jit Texture<Float, 2> without hardware acceleration (ext/drjit/include/drjit/texture.h: wrap :1143-1168, nearest
// :346-360, bilinear :362-405) under src/textures/bitmap.cpp:582-625: uv' = to_uv·uv, RGB texels are linear floats.
inline int tex_wrap(int p, int n, uint32_t mode) {
    if (mode == EPSM_TEX_CLAMP) return p < 0 ? 0 : (p > n - 1 ? n - 1 : p);
    // floor division (divisor<int32_t> with the sign fix-up of texture.h:1150-1153)
    int div = p / n; if (p < 0 && div * n != p) --div;
    int mod = p - div * n;
    if (mode == EPSM_TEX_MIRROR && (div & 1)) mod = n - 1 - mod;
    return mod;
}
inline V3 tex_fetch(const epsm_texture_desc &t, int x, int y) {
    const float *c = t.texels + 3 * ((size_t) tex_wrap(y, (int) t.height, t.wrap) * t.width + (size_t) tex_wrap(x, (int) t.width, t.wrap));
    return mk(c[0], c[1], c[2]);
}
inline V3 tex_eval(const epsm_texture_desc &t, const float *uv_in) {
    const float u = fmaf(t.to_uv[0], uv_in[0], fmaf(t.to_uv[1], uv_in[1], t.to_uv[2])), v = fmaf(t.to_uv[3], uv_in[0], fmaf(t.to_uv[4], uv_in[1], t.to_uv[5]));
    if (t.filter == EPSM_TEX_NEAREST) {
        const float fx = u * (float) t.width, fy = v * (float) t.height;
        return tex_fetch(t, (int) floorf(fx), (int) floorf(fy));
    }
    const float fx = fmaf(u, (float) t.width, -.5f), fy = fmaf(v, (float) t.height, -.5f);
    const int ix = (int) floorf(fx), iy = (int) floorf(fy);
    const float w1x = fx - (float) ix, w1y = fy - (float) iy, w0x = 1.f - w1x, w0y = 1.f - w1y;
    const V3 v00 = tex_fetch(t, ix, iy), v10 = tex_fetch(t, ix + 1, iy), v01 = tex_fetch(t, ix, iy + 1), v11 = tex_fetch(t, ix + 1, iy + 1);
    // fmadd(w0.y, fmadd(w0.x, v00, w1.x·v10), w1.y·fmadd(w0.x, v01, w1.x·v11))   (texture.h:398-403)
    const V3 r0 = fmadd(v00, w0x, v10 * w1x), r1 = fmadd(v01, w0x, v11 * w1x);
    return fmadd(r0, w0y, r1 * w1y);
}
// The BSDF of a hit with its textured parameters evaluated at si.uv (diffuse.cpp:107-140: m_reflectance->eval(si)).
inline epsm_bsdf_desc bsdf_at(const Scene &sc, const SI &si) {
    epsm_bsdf_desc b = sc.bsdfs[sc.meshes[si.mesh].bsdf];
    if (b.reflectance_texture) { const V3 c = tex_eval(*b.reflectance_texture, si.uv); b.reflectance[0] = c.x; b.reflectance[1] = c.y; b.reflectance[2] = c.z; }
    return b;
}

// interaction.h:136-165
inline V3 offset_p(const SI &si, V3 d) {
    float mag = (1.f + max3(mk(fabsf(si.p.x), fabsf(si.p.y), fabsf(si.p.z)))) * kRayEps;
    mag = mulsign(mag, dot(si.n, d));
    return fmadd(si.n, mag, si.p);
}
inline Ray spawn_ray(const SI &si, V3 d) { return Ray{offset_p(si, d), d, FLT_MAX}; }
inline Ray spawn_ray_to(const SI &si, V3 t) {
    V3 o = offset_p(si, t - si.p); V3 d = t - o; float dist = norm(d);
    d = d * (1.0f / dist);
    // `d /= dist` in the reference: Dr.Jit divides arrays by a scalar via multiplication with rcp
    return Ray{o, d, dist * (1.f - kShadowEps)};
}

// ------------------------------------------------------------------------------------------------
//  BSDFs
// ------------------------------------------------------------------------------------------------
struct BSample { V3 wo, hf; float pdf, eta; uint32_t type; V3 weight; };

uint32_t bsdf_flags(const epsm_bsdf_desc &b) {      // BSDF::flags() incl. twosided's (bsdf.h; twosided.cpp:80-100)
    const uint32_t Front = 0x8000u, Back = 0x10000u, NonSym = 0x4000u;
    uint32_t f = 0;
    switch (b.kind) {
        case EPSM_BSDF_DIFFUSE: f = EPSM_BF_DIFFUSE_REFLECTION | Front; break;
        case EPSM_BSDF_DIELECTRIC: f = EPSM_BF_DELTA_REFLECTION | EPSM_BF_DELTA_TRANSMISSION | Front | Back | NonSym; break;
        case EPSM_BSDF_CONDUCTOR: f = EPSM_BF_DELTA_REFLECTION | Front; break;
        case EPSM_BSDF_ROUGHCONDUCTOR: f = EPSM_BF_GLOSSY_REFLECTION | Front | (b.alpha_u != b.alpha_v ? 0x1000u : 0u); break;
    }
    if (b.flags & EPSM_BSDF_TWOSIDED) f |= Back;
    return f;
}

// include/mitsuba/render/fresnel.h:35-72
void fresnel(float cos_theta_i, float eta, float *r_, float *cos_theta_t_, float *eta_it_, float *eta_ti_) {
    bool outside = cos_theta_i >= 0.f;
    float rcp_eta = rcp_(eta), eta_it = outside ? eta : rcp_eta, eta_ti = outside ? rcp_eta : eta;
    float cos_theta_t_sqr = fmaf(-fmaf(-cos_theta_i, cos_theta_i, 1.f), eta_ti * eta_ti, 1.f);
    float ci = fabsf(cos_theta_i), ct = safe_sqrt(cos_theta_t_sqr);
    bool index_matched = eta == 1.f, special = index_matched || ci == 0.f;
    float r_sc = index_matched ? 0.f : 1.f;
    float a_s = fmaf(-eta_it, ct, ci) / fmaf(eta_it, ct, ci);
    float a_p = fmaf(-eta_it, ci, ct) / fmaf(eta_it, ci, ct);
    float r = 0.5f * (sqr(a_s) + sqr(a_p));
    if (special) r = r_sc;
    *r_ = r; *cos_theta_t_ = mulsign_neg(ct, cos_theta_i); *eta_it_ = eta_it; *eta_ti_ = eta_ti;
}
// fresnel.h:93-117
float fresnel_conductor(float cos_theta_i, float eta_r, float eta_i) {
    float c2 = cos_theta_i * cos_theta_i, s2 = 1.f - c2, s4 = s2 * s2;
    float temp_1 = eta_r * eta_r - eta_i * eta_i - s2;
    float a_2_pb_2 = safe_sqrt(temp_1 * temp_1 + 4.f * eta_i * eta_i * eta_r * eta_r);
    float a = safe_sqrt(.5f * (a_2_pb_2 + temp_1));
    float term_1 = a_2_pb_2 + c2, term_2 = 2.f * cos_theta_i * a;
    float r_s = (term_1 - term_2) / (term_1 + term_2);
    float term_3 = a_2_pb_2 * c2 + s4, term_4 = term_2 * s2;
    float r_p = r_s * (term_3 - term_4) / (term_3 + term_4);
    return 0.5f * (r_s + r_p);
}
V3 fresnel_conductor3(float c, const float *eta, const float *k) {
    return mk(fresnel_conductor(c, eta[0], k[0]), fresnel_conductor(c, eta[1], k[1]), fresnel_conductor(c, eta[2], k[2]));
}

// include/mitsuba/render/microfacet.h:183-210, 244-302 (the EPSM fork's `if (true)` branch: plain D sampling), 365-392
struct Microfacet {
    bool ggx; float au, av; bool sample_visible;
    float eval(V3 m) const {
        float alpha_uv = au * av, ct = m.z, ct2 = sqr(ct), result;
        if (!ggx)
            result = (float) std::exp((double) (-(sqr(m.x / au) + sqr(m.y / av)) / ct2)) / (kPi * alpha_uv * sqr(ct2));
        else
            result = rcp_(kPi * alpha_uv * sqr(sqr(m.x / au) + sqr(m.y / av) + sqr(m.z)));
        return result * ct > 1e-20f ? result : 0.f;
    }
    float smith_g1(V3 v, V3 m) const {
        float xy_alpha_2 = sqr(au * v.x) + sqr(av * v.y), tan_theta_alpha_2 = xy_alpha_2 / sqr(v.z), result;
        if (!ggx) {
            float a = rsqrt_(tan_theta_alpha_2), a_sqr = sqr(a);
            result = a >= 1.6f ? 1.f : (3.535f * a + 2.181f * a_sqr) / (1.f + 2.276f * a + 2.577f * a_sqr);
        } else {
            result = 2.f / (1.f + sqrtf(1.f + tan_theta_alpha_2));
        }
        if (xy_alpha_2 == 0.f) result = 1.f;
        if (dot(v, m) * v.z <= 0.f) result = 0.f;
        return result;
    }
    void sample(float sx, float sy, V3 *m, float *pdf) const {
        float sin_phi, cos_phi, cos_theta, cos_theta_2, alpha_2;
        if (au == av) {
            sincos_dr((2.f * kPi) * sy, &sin_phi, &cos_phi);
            alpha_2 = au * au;
        } else {
            float ratio = av / au, tmp = ratio * (float) std::tan((double) ((2.f * kPi) * sy));
            cos_phi = rsqrt_(fmaf(tmp, tmp, 1.f));
            cos_phi = mulsign(cos_phi, fabsf(sy - .5f) - .25f);
            sin_phi = cos_phi * tmp;
            alpha_2 = rcp_(sqr(cos_phi / au) + sqr(sin_phi / av));
        }
        if (!ggx) {
            cos_theta = rsqrt_(fmaf(-alpha_2, (float) std::log((double) (1.f - sx)), 1.f));
            cos_theta_2 = sqr(cos_theta);
            float cos_theta_3 = fmaxf(cos_theta_2 * cos_theta, 1e-20f);
            *pdf = (1.f - sx) / (kPi * au * av * cos_theta_3);
        } else {
            float tan_theta_m_2 = alpha_2 * sx / (1.f - sx);
            cos_theta = rsqrt_(1.f + tan_theta_m_2);
            cos_theta_2 = sqr(cos_theta);
            float temp = 1.f + tan_theta_m_2 / alpha_2, cos_theta_3 = fmaxf(cos_theta_2 * cos_theta, 1e-20f);
            *pdf = rcp_(kPi * au * av * cos_theta_3 * sqr(temp));
        }
        float sin_theta = sqrtf(1.f - cos_theta_2);
        *m = mk(cos_phi * sin_theta, sin_phi * sin_theta, cos_theta);
    }
};
inline V3 reflect(V3 wi, V3 m) { float s = 2.f * dot(wi, m); return mk(fmaf(m.x, s, -wi.x), fmaf(m.y, s, -wi.y), fmaf(m.z, s, -wi.z)); }  // fresnel.h:282-284

// The nested BSDF's sample(); `wi` already mapped by twosided.
BSample bsdf_sample_inner(const epsm_bsdf_desc &b, V3 wi, float s1, float s2x, float s2y, bool active) {
    BSample bs; memset(&bs, 0, sizeof(bs));
    V3 refl = mk(b.reflectance[0], b.reflectance[1], b.reflectance[2]);
    switch (b.kind) {
    case EPSM_BSDF_DIFFUSE: {                                     // src/bsdfs/diffuse.cpp:101-125
        active &= wi.z > 0.f;
        if (!active) return bs;
        bs.wo = square_to_cosine_hemisphere(s2x, s2y);
        bs.pdf = kInvPi * bs.wo.z; bs.eta = 1.f; bs.type = EPSM_BF_DIFFUSE_REFLECTION;
        if (bs.pdf > 0.f) bs.weight = refl;
        return bs;
    }
    case EPSM_BSDF_DIELECTRIC: {                                  // src/bsdfs/dielectric.cpp:250-366
        float r_i, cos_theta_t, eta_it, eta_ti;
        fresnel(wi.z, b.eta[0], &r_i, &cos_theta_t, &eta_it, &eta_ti);
        float t_i = 1.f - r_i;
        bool sel_r = s1 <= r_i && active, sel_t = !sel_r && active;
        bs.pdf = sel_r ? r_i : t_i;
        bs.type = sel_r ? EPSM_BF_DELTA_REFLECTION : EPSM_BF_DELTA_TRANSMISSION;
        bs.wo = sel_r ? mk(-wi.x, -wi.y, wi.z) : mk(-eta_ti * wi.x, -eta_ti * wi.y, cos_theta_t);
        bs.eta = sel_r ? 1.f : eta_it;
        V3 w = mk(1, 1, 1);
        if (sel_r) w = w * refl;
        if (sel_t) { w = w * mk(b.transmittance[0], b.transmittance[1], b.transmittance[2]); w = w * sqr(eta_ti); }
        if (active) bs.weight = w;
        return bs;
    }
    case EPSM_BSDF_CONDUCTOR: {                                   // src/bsdfs/conductor.cpp:248-300
        active &= wi.z > 0.f;
        if (!active) return bs;
        bs.type = EPSM_BF_DELTA_REFLECTION; bs.wo = mk(-wi.x, -wi.y, wi.z); bs.eta = 1.f; bs.pdf = 1.f;
        bs.weight = refl * fresnel_conductor3(wi.z, b.eta, b.k);
        return bs;
    }
    case EPSM_BSDF_ROUGHCONDUCTOR: {                              // src/bsdfs/roughconductor.cpp:226-305
        active &= wi.z > 0.f;
        if (!active) return bs;
        Microfacet d{(b.flags & EPSM_BSDF_GGX) != 0, b.alpha_u, b.alpha_v, (b.flags & EPSM_BSDF_SAMPLE_VISIBLE) != 0};
        V3 m; d.sample(s2x, s2y, &m, &bs.pdf);
        bs.wo = reflect(wi, m); bs.eta = 1.f; bs.type = EPSM_BF_GLOSSY_REFLECTION; bs.hf = m;
        active &= bs.pdf != 0.f && bs.wo.z > 0.f;
        float weight = d.sample_visible ? d.smith_g1(bs.wo, m)
                                        : d.smith_g1(wi, m) * d.smith_g1(bs.wo, m) * dot(wi, m) / (wi.z * m.z);
        bs.pdf /= 4.f * dot(bs.wo, m);
        V3 F = fresnel_conductor3(dot(wi, m), b.eta, b.k);
        if (active) bs.weight = F * weight * refl;
        return bs;
    }
    }
    return bs;
}
// twosided.cpp:111-150 (both sides share one BRDF)
BSample bsdf_sample(const epsm_bsdf_desc &b, V3 wi, float s1, float s2x, float s2y, bool active) {
    if (b.flags & EPSM_BSDF_TWOSIDED) {
        V3 w = wi; w.z = fabsf(w.z);
        BSample bs = bsdf_sample_inner(b, w, s1, s2x, s2y, active);
        bs.wo.z = mulsign(bs.wo.z, wi.z);
        return bs;
    }
    return bsdf_sample_inner(b, wi, s1, s2x, s2y, active);
}
void bsdf_eval_pdf_inner(const epsm_bsdf_desc &b, V3 wi, V3 wo, bool active, V3 *value, float *pdf) {
    *value = mk(0, 0, 0); *pdf = 0.f;
    V3 refl = mk(b.reflectance[0], b.reflectance[1], b.reflectance[2]);
    switch (b.kind) {
    case EPSM_BSDF_DIFFUSE: {                                     // diffuse.cpp:158-175
        active &= wi.z > 0.f && wo.z > 0.f;
        if (active) { *value = refl * kInvPi * wo.z; *pdf = kInvPi * wo.z; }
        return;
    }
    case EPSM_BSDF_ROUGHCONDUCTOR: {                              // roughconductor.cpp:417-492
        V3 H = normalize(wo + wi);
        active &= wi.z > 0.f && wo.z > 0.f && dot(wi, H) > 0.f && dot(wo, H) > 0.f;
        if (!active) return;
        Microfacet d{(b.flags & EPSM_BSDF_GGX) != 0, b.alpha_u, b.alpha_v, (b.flags & EPSM_BSDF_SAMPLE_VISIBLE) != 0};
        float D = d.eval(H);
        active &= D != 0.f;
        float g1 = d.smith_g1(wi, H), G = g1 * d.smith_g1(wo, H);
        float val = D * G / (4.f * wi.z);
        V3 F = fresnel_conductor3(dot(wi, H), b.eta, b.k);
        float p = d.sample_visible ? D * g1 / (4.f * wi.z) : D * H.z / (4.f * dot(wo, H));
        if (active) { *value = F * (refl * val); *pdf = p; }
        return;
    }
    default: return;                                              // delta BSDFs: eval = pdf = 0
    }
}
void bsdf_eval_pdf(const epsm_bsdf_desc &b, V3 wi, V3 wo, bool active, V3 *value, float *pdf) {
    if (b.flags & EPSM_BSDF_TWOSIDED) { wo.z = mulsign(wo.z, wi.z); wi.z = fabsf(wi.z); }   // twosided.cpp:226-232
    bsdf_eval_pdf_inner(b, wi, wo, active, value, pdf);
}

// ------------------------------------------------------------------------------------------------
//  Emitters: src/render/scene.cpp:161-290, src/emitters/area.cpp:82-160, src/render/shape.cpp:352-383,
//            src/render/mesh.cpp:516-571, include/mitsuba/core/distr_1d.h:117-176
// ------------------------------------------------------------------------------------------------
struct DS { V3 p, n, d; float pdf, dist; int32_t emitter; };

uint32_t pmf_sample(const Mesh &m, float value) {
    value *= m.area_sum;
    uint32_t lo = m.valid_lo, hi = m.valid_hi;           // dr::binary_search: first index with !(cdf[i] < value)
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (m.cdf[mid] < value) lo = std::min(mid + 1, hi); else hi = mid; }
    return lo;
}
V3 emitter_eval(const Scene &sc, const SI &si) {                                    // area.cpp:82-89
    if (!si.valid) return mk(0, 0, 0);
    int32_t e = sc.meshes[si.mesh].emitter;
    if (e < 0 || !(si.wi.z > 0.f)) return mk(0, 0, 0);
    return mk(sc.emitters[e].radiance[0], sc.emitters[e].radiance[1], sc.emitters[e].radiance[2]);
}
// Scene::sample_emitter_direction with test_visibility = true.  `spec` is radiance/pdf (0 if invalid/occluded).
bool ray_test(const Scene &sc, const Ray &r);

void sample_emitter_direction(const Scene &sc, const SI &ref, float sx, float sy, bool active, DS *ds, V3 *spec) {
    memset(ds, 0, sizeof(*ds)); ds->emitter = -1; *spec = mk(0, 0, 0);
    size_t ne = sc.emitters.size();
    if (ne == 0) return;
    uint32_t index = 0; float emitter_weight = 1.f, pmf = 1.f;
    if (ne > 1) {                                                                   // scene.cpp:161-178
        float ef = (float) ne, scaled = sx * ef;
        index = std::min((uint32_t) scaled, (uint32_t) ne - 1u);
        sx = scaled - (float) index; emitter_weight = ef; pmf = 1.f / ef;          // m_emitter_pmf, scene.cpp:87
    }
    const epsm_emitter_desc &em = sc.emitters[index]; const Mesh &m = sc.meshes[em.mesh_id];
    V3 n;
    if (m.flags & EPSM_MESH_RECTANGLE) {
        // Rectangle::sample_position, src/shapes/rectangle.cpp:152-166
        V3 c0 = sc.pos[m.v_off], ex = sc.pos[m.v_off + 1] - c0, ey = sc.pos[m.v_off + 3] - c0;
        ds->p = fmadd(ex, sx, fmadd(ey, sy, c0));
        n = normalize(cross(ex, ey));
    } else {
        // Mesh::sample_position
        uint32_t face = pmf_sample(m, sy);
        {   // sample_reuse, distr_1d.h:167-176
            float pm = m.pmf[face] * m.area_norm, cdf = face > 0 ? m.cdf[face - 1] * m.area_norm : 0.f;
            sy = (sy - cdf) / pm;
        }
        const uint32_t *f = &sc.faces[3 * (m.f_off + face)];
        V3 p0 = sc.pos[f[0]], p1 = sc.pos[f[1]], p2 = sc.pos[f[2]], e0 = p1 - p0, e1 = p2 - p0;
        float t = safe_sqrt(1.f - sx), bx = 1.f - t, by = t * sy;                   // warp.h:153-156
        ds->p = fmadd(e0, bx, fmadd(e1, by, p0));
        if (m.has_normals) n = fmadd(sc.nrm[f[0]], (1.f - bx - by), fmadd(sc.nrm[f[1]], bx, sc.nrm[f[2]] * by));
        else n = cross(e0, e1);
        n = normalize(n);
    }
    if (m.flags & EPSM_MESH_FLIP_NORMALS) n = -n;
    ds->n = n; ds->pdf = m.area_norm; ds->emitter = (int32_t) index;
    // Shape::sample_direction
    ds->d = ds->p - ref.p;
    float dist2 = sqnorm(ds->d); ds->dist = sqrtf(dist2);
    ds->d = ds->d * (1.0f / ds->dist);
    float dp = fabsf(dot(ds->d, ds->n)), x = dist2 / dp;
    ds->pdf *= std::isfinite(x) ? x : 0.f;
    // area.cpp:127
    bool act = active && dot(ds->d, ds->n) < 0.f && ds->pdf != 0.f;
    V3 rad = mk(em.radiance[0], em.radiance[1], em.radiance[2]);
    V3 s = act ? rad * (1.0f / ds->pdf) : mk(0, 0, 0);
    ds->pdf *= pmf; s = s * emitter_weight;
    bool act2 = active && ds->pdf != 0.f;                                           // scene.cpp:253 (independent of back-facing test)
    if (act2) {
        if (ray_test(sc, spawn_ray_to(ref, ds->p))) { s = mk(0, 0, 0); ds->pdf = 0.f; }
    }
    if (!active) { ds->pdf = 0.f; s = mk(0, 0, 0); }
    *spec = s;
}
// Scene::pdf_emitter_direction for a DirectionSample built from a surface hit (records.h:173-180)
float pdf_emitter_direction(const Scene &sc, const SI &si, V3 ref_p, bool active) {
    if (!si.valid || !active) return 0.f;
    int32_t e = sc.meshes[si.mesh].emitter;
    if (e < 0) return 0.f;
    V3 rel = si.p - ref_p; float dist = norm(rel); V3 d = rel * (1.0f / dist);
    float dp = dot(d, si.sh_n);
    if (!(dp < 0.f)) return 0.f;                                                    // area.cpp:140-142
    const Mesh &m = sc.meshes[sc.emitters[e].mesh_id];
    float adp = fabsf(dp);
    float pdf = m.area_norm * (adp != 0.f ? (dist * dist) / adp : 0.f);            // shape.cpp:372-383
    return pdf * (1.0f / (float) sc.emitters.size());
}
inline float mis_weight(float a, float b) { float a2 = a * a; return a > 0.f ? a2 / fmaf(b, b, a2) : 0.f; }   // common.py:1224-1230

// Optional median-split BVH used only to make the CPU baseline tractable on large scenes; validated against
// the brute-force loops above in tests/test_oracle.py.  Same Möller–Trumbore, same tie rule.
struct BNode { float lo[3], hi[3]; uint32_t left, right, first, count; };
struct OBVH { std::vector<BNode> nodes; std::vector<uint32_t> prims; bool built = false; };

} // namespace

struct orc_scene { Scene sc; OBVH bvh; bool use_bvh = false; };

namespace {
int g_prune_variant = -1;                  // orc_set_prune: −1 full replay (the reference), 0 / 1 = prune like the product for manifold / manifold_caustic
const orc_scene *g_accel_owner = nullptr;   // set per call (single-threaded entry; OpenMP threads only read)

void bvh_build_rec(const Scene &sc, OBVH &b, uint32_t node, uint32_t first, uint32_t count, std::vector<V3> &cent) {
    BNode &n = b.nodes[node];
    for (int a = 0; a < 3; ++a) { n.lo[a] = FLT_MAX; n.hi[a] = -FLT_MAX; }
    float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (uint32_t i = first; i < first + count; ++i) {
        uint32_t p = b.prims[i]; const uint32_t *f = &sc.faces[3 * p];
        for (int k = 0; k < 3; ++k) {
            const float *v = &sc.pos[f[k]].x;
            for (int a = 0; a < 3; ++a) { n.lo[a] = std::min(n.lo[a], v[a]); n.hi[a] = std::max(n.hi[a], v[a]); }
        }
        const float *c = &cent[p].x;
        for (int a = 0; a < 3; ++a) { clo[a] = std::min(clo[a], c[a]); chi[a] = std::max(chi[a], c[a]); }
    }
    for (int a = 0; a < 3; ++a) {           // pad: Möller–Trumbore accepts hits a few ulp outside the exact box
        float pad = 1e-4f * (1.f + std::max(fabsf(n.lo[a]), fabsf(n.hi[a])));
        n.lo[a] -= pad; n.hi[a] += pad;
    }
    n.first = first; n.count = count; n.left = n.right = 0;
    if (count <= 4) return;
    int axis = 0; float ext = chi[0] - clo[0];
    for (int a = 1; a < 3; ++a) if (chi[a] - clo[a] > ext) { ext = chi[a] - clo[a]; axis = a; }
    if (!(ext > 0)) return;
    uint32_t mid = first + count / 2;
    std::nth_element(b.prims.begin() + first, b.prims.begin() + mid, b.prims.begin() + first + count,
                     [&](uint32_t x, uint32_t y) { return (&cent[x].x)[axis] < (&cent[y].x)[axis]; });
    uint32_t l = (uint32_t) b.nodes.size(); b.nodes.push_back(BNode()); b.nodes.push_back(BNode());
    b.nodes[node].left = l; b.nodes[node].right = l + 1; b.nodes[node].count = 0;
    bvh_build_rec(sc, b, l, first, mid - first, cent);
    bvh_build_rec(sc, b, l + 1, mid, first + count - mid, cent);
}
void bvh_build(const Scene &sc, OBVH &b) {
    size_t nf = sc.faces.size() / 3;
    b.prims.resize(nf); std::vector<V3> cent(nf);
    for (size_t i = 0; i < nf; ++i) {
        b.prims[i] = (uint32_t) i; const uint32_t *f = &sc.faces[3 * i];
        cent[i] = (sc.pos[f[0]] + sc.pos[f[1]] + sc.pos[f[2]]) * (1.f / 3.f);
    }
    b.nodes.clear(); b.nodes.reserve(nf); b.nodes.push_back(BNode());
    bvh_build_rec(sc, b, 0, 0, (uint32_t) nf, cent);
    b.built = true;
}
inline bool box_hit(const BNode &n, const Ray &r, const float *inv, float tmax) {
    float t0 = 0.f, t1 = tmax;
    const float *o = &r.o.x;
    for (int a = 0; a < 3; ++a) {
        float ta = (n.lo[a] - o[a]) * inv[a], tb = (n.hi[a] - o[a]) * inv[a];
        if (ta > tb) std::swap(ta, tb);
        if (!(ta != ta)) t0 = std::max(t0, ta);
        if (!(tb != tb)) t1 = std::min(t1, tb);
    }
    return t0 <= t1 * 1.00001f;
}
PI bvh_closest(const orc_scene &os, Ray ray) {
    const Scene &sc = os.sc; const OBVH &b = os.bvh;
    PI pi{false, INFINITY, 0, 0, 0xffffffffu};
    float inv[3] = {1.f / ray.d.x, 1.f / ray.d.y, 1.f / ray.d.z};
    uint32_t stack[128]; int sp = 0; stack[sp++] = 0;
    while (sp) {
        const BNode &n = b.nodes[stack[--sp]];
        if (!box_hit(n, ray, inv, ray.maxt)) continue;
        if (n.count) {
            for (uint32_t i = n.first; i < n.first + n.count; ++i) {
                uint32_t p = b.prims[i]; const uint32_t *f = &sc.faces[3 * p]; float t, u, v;
                if (moeller_trumbore(ray, sc.pos[f[0]], sc.pos[f[1]], sc.pos[f[2]], &t, &u, &v)) {
                    // order-independent form of "later primitive wins ties"
                    if (t < pi.t || !pi.valid || (t == pi.t && p > pi.prim)) { pi.valid = true; pi.t = t; pi.u = u; pi.v = v; pi.prim = p; ray.maxt = t; }
                }
            }
        } else { stack[sp++] = n.left; stack[sp++] = n.right; }
    }
    return pi;
}
bool bvh_any(const orc_scene &os, const Ray &ray) {
    const Scene &sc = os.sc; const OBVH &b = os.bvh;
    float inv[3] = {1.f / ray.d.x, 1.f / ray.d.y, 1.f / ray.d.z};
    uint32_t stack[128]; int sp = 0; stack[sp++] = 0;
    while (sp) {
        const BNode &n = b.nodes[stack[--sp]];
        if (!box_hit(n, ray, inv, ray.maxt)) continue;
        if (n.count) {
            for (uint32_t i = n.first; i < n.first + n.count; ++i) {
                const uint32_t *f = &sc.faces[3 * b.prims[i]]; float t, u, v;
                if (moeller_trumbore(ray, sc.pos[f[0]], sc.pos[f[1]], sc.pos[f[2]], &t, &u, &v)) return true;
            }
        } else { stack[sp++] = n.left; stack[sp++] = n.right; }
    }
    return false;
}
inline PI scene_closest(const Scene &sc, const Ray &r) {
    const orc_scene *os = g_accel_owner;
    return (os && os->use_bvh) ? bvh_closest(*os, r) : intersect_closest(sc, r);
}
bool ray_test(const Scene &sc, const Ray &r) {
    const orc_scene *os = g_accel_owner;
    return (os && os->use_bvh) ? bvh_any(*os, r) : intersect_any(sc, r);
}

// ------------------------------------------------------------------------------------------------
//  Ray generation: common.py:291-422 + perspective.cpp:238-279
// ------------------------------------------------------------------------------------------------
struct CamRay { Ray ray; V3 d_x, d_y; float pos_x, pos_y; uint32_t px, py; };
CamRay sample_camera_ray(const Sensor &s, uint32_t lane, uint32_t spp, PCG32 &rng) {
    uint32_t idx = lane / spp;
    uint32_t py = idx / s.w, px = idx - py * s.w;
    float jx = rng.next_f32(), jy = rng.next_f32();
    float pos_x = (float) ((int) px - s.border) + jx, pos_y = (float) ((int) py - s.border) + jy;       // common.py:337-338
    float sx = 1.0f / (float) s.fw, sy = 1.0f / (float) s.fh;
    float ax = fmaf(pos_x, sx, -0.f * sx), ay = fmaf(pos_y, sy, -0.f * sy);      // offset = -crop_offset*scale = -0
    V3 near_p = xform_point(s.s2c, mk(ax, ay, 0.f));
    V3 d = normalize(near_p);
    CamRay c;
    V3 o = mk(s.to_world[3], s.to_world[7], s.to_world[11]);
    V3 dw = xform_vec(s.to_world, d);
    float inv_z = rcp_(d.z), near_t = s.near_clip * inv_z, far_t = s.far_clip * inv_z;
    c.ray.o = o + dw * near_t; c.ray.d = dw; c.ray.maxt = far_t - near_t;
    c.d_x = xform_vec(s.to_world, normalize(near_p + s.dx));
    c.d_y = xform_vec(s.to_world, normalize(near_p + s.dy));
    c.pos_x = pos_x; c.pos_y = pos_y; c.px = px; c.py = py;
    return c;
}

// ------------------------------------------------------------------------------------------------
//  EPSMIntegrator.sample — primal path loop (epsm.py:308-501), 6 random numbers per bounce
// ------------------------------------------------------------------------------------------------
V3 sample_primal(const Scene &sc, Ray ray, PCG32 &rng, uint32_t max_depth, uint32_t rr_depth, bool *valid_out) {
    uint32_t depth = 0; V3 L = mk(0, 0, 0), beta = mk(1, 1, 1); float eta = 1.f; bool active = true;
    V3 prev_p = mk(0, 0, 0); float prev_pdf = 1.f; bool prev_delta = true;
    for (uint32_t it = 0; it < max_depth && active; ++it) {
        PI pi = scene_closest(sc, ray);
        SI si = compute_si(sc, ray, pi);
        float em_pdf = pdf_emitter_direction(sc, si, prev_p, !prev_delta);
        float mis = mis_weight(prev_pdf, em_pdf);
        V3 Le = beta * mis * emitter_eval(sc, si);
        bool active_next = (depth + 1 < max_depth) && si.valid;
        epsm_bsdf_desc b_tex; const epsm_bsdf_desc *b = nullptr; if (si.valid) { b_tex = bsdf_at(sc, si); b = &b_tex; }
        uint32_t flags = b ? bsdf_flags(*b) : 0u;
        bool active_em = active_next && (flags & EPSM_BF_SMOOTH);
        float e1 = rng.next_f32(), e2 = rng.next_f32();
        DS ds; V3 em_weight;
        sample_emitter_direction(sc, si, e1, e2, active_em, &ds, &em_weight);
        active_em &= ds.pdf != 0.f;
        V3 Lr_dir = mk(0, 0, 0);
        if (active_em) {
            V3 wo = to_local(si, ds.d); V3 val; float bpdf;
            bsdf_eval_pdf(*b, si.wi, wo, true, &val, &bpdf);
            float mis_em = mis_weight(ds.pdf, bpdf);
            Lr_dir = beta * mis_em * val * em_weight;
        }
        float s1 = rng.next_f32(), s2x = rng.next_f32(), s2y = rng.next_f32();
        BSample bs; memset(&bs, 0, sizeof(bs));
        if (b) bs = bsdf_sample(*b, si.wi, s1, s2x, s2y, active_next);
        L = L + Le + Lr_dir;
        if (si.valid) ray = spawn_ray(si, to_world(si, bs.wo));
        eta *= bs.eta; beta = beta * bs.weight;
        prev_p = si.p; prev_pdf = bs.pdf; prev_delta = (bs.type & EPSM_BF_DELTA) != 0;
        float bmax = max3(beta);
        active_next &= bmax != 0.f;
        float rr_prob = fminf(bmax * eta * eta, .95f);
        bool rr_active = depth >= rr_depth;
        if (rr_active) beta = beta * rcp_(rr_prob);
        bool rr_continue = rng.next_f32() < rr_prob;
        active_next &= !rr_active || rr_continue;
        if (si.valid) depth += 1;
        active = active_next;
    }
    *valid_out = depth != 0;
    return L;
}

// ------------------------------------------------------------------------------------------------
//  Film: imageblock.cpp:414-494 (coalesced put), hdrfilm.cpp:357-383 (develop)
// ------------------------------------------------------------------------------------------------
void film_put(const Scene &sc, std::vector<float> &film, uint32_t W, uint32_t H, float pos_x, float pos_y, V3 value) {
    const int n = 2, count = 5;                         // ceil(radius - .5) with radius = 2
    int pix = (int) floorf(pos_x) - n, piy = (int) floorf(pos_y) - n;
    float rel_x = (float) pix + .5f - pos_x, rel_y = (float) piy + .5f - pos_y;
    float wx[5], wy[5];
    for (int i = 0; i < count; ++i) { wx[i] = filter_eval(sc, rel_x); wy[i] = filter_eval(sc, rel_y); rel_x += 1; rel_y += 1; }
    float vals[4] = {value.x, value.y, value.z, 1.0f};
    for (int ys = 0; ys < count; ++ys) for (int xs = 0; xs < count; ++xs) {
        int x = pix + xs, y = piy + ys;
        if (x < 0 || y < 0 || x >= (int) W || y >= (int) H) continue;   // border pixels are cropped by put_block
        float w = wy[ys] * wx[xs];
        for (int k = 0; k < 4; ++k) film[((size_t) y * W + x) * 4 + k] += vals[k] * w;
    }
}

// ------------------------------------------------------------------------------------------------
//  EPSMIntegrator.sample_path (epsm.py:503-742) — the logging pass, 9 random numbers per bounce,
//  plus the quantities render_backward needs around it (epsm.py:238-272).
// ------------------------------------------------------------------------------------------------
} // namespace

extern "C" {

// One logged vertex (epsm.py:648-654) plus what the backward pass re-derives at the same bounce.
struct orc_vertex {
    float active, ismesh, active_em, eta;
    uint32_t bsdf_flags, prim, fi[3];
    float light[3], p0[3], p1[3], p2[3], p[3], b0, b1, sh_n[3], n0[3], n1[3], n2[3], hf[3];
    float bw[3];                 // barycentric weights of si.p on (fi[0],fi[1],fi[2])
    float lr_dir_sum;            // Σ_rgb Lr_dir (epsm.py:627)
    // si_direct: closest hit of si.spawn_ray(normalize(ds.p - si.p)), epsm.py:622-625
    uint32_t direct_valid, direct_fi[3]; float direct_bw[3];
    float direct_dis;            // epsm.py:614-615 (only meaningful for iteration 0)
    uint32_t direct_any;         // closest hit exists for the (active_next & Smooth) mask of epsm.py:611-613
    uint32_t mesh_flags, has_normals;
};
struct orc_path {
    float cam[3], d[3], dlduv[2], dldp1[3], L[3];
    uint32_t n_vertices;
    orc_vertex v[5];
};

orc_scene *orc_scene_create(const epsm_mesh_desc *meshes, uint32_t n_meshes, const epsm_bsdf_desc *bsdfs, uint32_t n_bsdfs,
                            const epsm_emitter_desc *emitters, uint32_t n_emitters,
                            const epsm_sensor_desc *sensors, uint32_t n_sensors) {
    orc_scene *os = new orc_scene(); Scene &sc = os->sc;
    for (uint32_t i = 0; i < n_meshes; ++i) {
        const epsm_mesh_desc &d = meshes[i]; Mesh m;
        m.v_off = (uint32_t) sc.pos.size(); m.f_off = (uint32_t) (sc.faces.size() / 3);
        m.v_cnt = d.vertex_count; m.f_cnt = d.face_count; m.bsdf = d.bsdf_id; m.emitter = d.emitter_id; m.flags = d.flags;
        m.has_normals = d.normals != nullptr; m.has_uv = d.texcoords != nullptr;
        for (uint32_t v = 0; v < d.vertex_count; ++v) {
            sc.pos.push_back(mk(d.positions[3 * v], d.positions[3 * v + 1], d.positions[3 * v + 2]));
            sc.nrm.push_back(d.normals ? mk(d.normals[3 * v], d.normals[3 * v + 1], d.normals[3 * v + 2]) : mk(0, 0, 0));
            sc.uv.push_back(d.texcoords ? d.texcoords[2 * v] : 0.f); sc.uv.push_back(d.texcoords ? d.texcoords[2 * v + 1] : 0.f);
        }
        for (uint32_t f = 0; f < d.face_count; ++f) {
            for (int k = 0; k < 3; ++k) sc.faces.push_back(d.faces[3 * f + k] + m.v_off);
            sc.face_mesh.push_back(i);
        }
        sc.meshes.push_back(m);
    }
    sc.bsdfs.assign(bsdfs, bsdfs + n_bsdfs);
    sc.textures.reserve(n_bsdfs);                                   // no reallocation: the descriptors point into it
    for (auto &b : sc.bsdfs) if (b.reflectance_texture) {
        Texture t; t.d = *b.reflectance_texture; t.texels.assign(t.d.texels, t.d.texels + 3 * (size_t) t.d.width * t.d.height);
        sc.textures.push_back(std::move(t)); sc.textures.back().d.texels = sc.textures.back().texels.data();
        b.reflectance_texture = &sc.textures.back().d;
    }
    sc.emitters.assign(emitters, emitters + n_emitters);
    for (auto &m : sc.meshes) build_pmf(sc, m);
    sc.sensors.resize(n_sensors);
    for (uint32_t i = 0; i < n_sensors; ++i) init_sensor(sc.sensors[i], sensors[i]);
    init_filter(sc);
    return os;
}
void orc_scene_destroy(orc_scene *s) { delete s; }
void orc_scene_update_vertices(orc_scene *s, uint32_t mesh_id, const float *positions) {
    Mesh &m = s->sc.meshes[mesh_id];
    for (uint32_t v = 0; v < m.v_cnt; ++v) s->sc.pos[m.v_off + v] = mk(positions[3 * v], positions[3 * v + 1], positions[3 * v + 2]);
    build_pmf(s->sc, m);
    if (s->use_bvh) bvh_build(s->sc, s->bvh);
}
// use_bvh = 1 switches the scene queries from the brute-force loops to the median-split BVH (CPU baseline only).
void orc_scene_set_accel(orc_scene *s, int use_bvh) {
    if (use_bvh && !s->bvh.built) bvh_build(s->sc, s->bvh);
    s->use_bvh = use_bvh != 0;
}

void orc_tea32(uint32_t v0, uint32_t v1, int rounds, uint32_t *out) { tea32(v0, v1, rounds, &out[0], &out[1]); }
float orc_tea_float32(uint32_t v0, uint32_t v1, int rounds) {                       // random.h:136-140 (uses .second)
    uint32_t a, b; tea32(v0, v1, rounds, &a, &b);
    uint32_t u = (b >> 9) | 0x3f800000u; float f; memcpy(&f, &u, 4); return f - 1.0f;
}
// PCG32(initstate, initseq).next_float32 stream — the scalar generator the reference tests compare against
void orc_pcg32_stream(uint64_t initstate, uint64_t initseq, uint32_t n, float *out) {
    PCG32 r; r.seed(initstate, initseq); for (uint32_t i = 0; i < n; ++i) out[i] = r.next_f32();
}
void orc_sampler_stream(uint32_t seed, uint64_t lane_begin, uint64_t lane_end, uint32_t n, float *out) {
    for (uint64_t l = lane_begin; l < lane_end; ++l) {
        PCG32 r = sampler_for_lane(seed, (uint32_t) l);
        for (uint32_t i = 0; i < n; ++i) out[(l - lane_begin) * n + i] = r.next_f32();
    }
}
void orc_sincos(float x, float *s, float *c) { sincos_dr(x, s, c); }
float orc_filter_eval(orc_scene *s, float x) { return filter_eval(s->sc, x); }

void orc_trace(orc_scene *s, const float *rays, uint64_t n, void *hits, int any_hit) {
    g_accel_owner = s;
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t) n; ++i) {
        const float *r = rays + 8 * i;
        Ray ray{mk(r[0], r[1], r[2]), mk(r[4], r[5], r[6]), r[3]};
        if (any_hit) { ((uint32_t *) hits)[i] = ray_test(s->sc, ray) ? 1u : 0u; continue; }
        PI pi = scene_closest(s->sc, ray);
        float *h = (float *) hits + 4 * i;
        h[0] = pi.valid ? pi.t : INFINITY; h[1] = pi.u; h[2] = pi.v;
        uint32_t prim = pi.valid ? pi.prim : 0xffffffffu; memcpy(&h[3], &prim, 4);
    }
}

// Surface-interaction probe for the KATs: fills 40 floats
// [t, p(3), n(3), sh_n(3), wi(3), b0, b1, ismesh, p0(3), p1(3), p2(3), n0(3), n1(3), n2(3), prim, bsdf_flags, u, v]
void orc_surface_interaction(orc_scene *s, const float *r, float *out) {
    g_accel_owner = s;
    Ray ray{mk(r[0], r[1], r[2]), mk(r[4], r[5], r[6]), r[3]};
    PI pi = scene_closest(s->sc, ray); SI si = compute_si(s->sc, ray, pi);
    float *o = out; auto put3 = [&](V3 v) { *o++ = v.x; *o++ = v.y; *o++ = v.z; };
    *o++ = si.t; put3(si.p); put3(si.n); put3(si.sh_n); put3(si.wi); *o++ = si.b0; *o++ = si.b1; *o++ = (float) si.ismesh;
    put3(si.p0); put3(si.p1); put3(si.p2); put3(si.n0); put3(si.n1); put3(si.n2);
    *o++ = (float) si.prim; *o++ = si.valid ? (float) bsdf_flags(s->sc.bsdfs[s->sc.meshes[si.mesh].bsdf]) : 0.f; *o++ = pi.u; *o++ = pi.v;
}
// Texture-side probe of a hit: out = [uv(2), sh_s(3), sh_t(3), diffuse reflectance at the hit (3), dp_du(3), dp_dv(3)]
void orc_surface_uv(orc_scene *s, const float *r, float *out) {
    g_accel_owner = s;
    Ray ray{mk(r[0], r[1], r[2]), mk(r[4], r[5], r[6]), r[3]};
    PI pi = scene_closest(s->sc, ray); SI si = compute_si(s->sc, ray, pi);
    for (int i = 0; i < 17; ++i) out[i] = 0.f;
    if (!si.valid) return;
    const epsm_bsdf_desc b = bsdf_at(s->sc, si);
    out[11] = si.dp_du.x; out[12] = si.dp_du.y; out[13] = si.dp_du.z; out[14] = si.dp_dv.x; out[15] = si.dp_dv.y; out[16] = si.dp_dv.z;
    out[0] = si.uv[0]; out[1] = si.uv[1]; out[2] = si.sh_s.x; out[3] = si.sh_s.y; out[4] = si.sh_s.z; out[5] = si.sh_t.x; out[6] = si.sh_t.y; out[7] = si.sh_t.z;
    out[8] = b.reflectance[0]; out[9] = b.reflectance[1]; out[10] = b.reflectance[2];
}
// Texture lookup on its own: n uv pairs → n RGB values
void orc_texture_eval(const epsm_texture_desc *t, const float *uv, uint32_t n, float *out) {
    for (uint32_t i = 0; i < n; ++i) { const V3 c = tex_eval(*t, uv + 2 * i); out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z; }
}
// BSDF probe: out = [wo(3), pdf, eta, type, weight(3), hf(3)]
void orc_bsdf_sample(const epsm_bsdf_desc *b, const float *wi, float s1, float s2x, float s2y, float *out) {
    BSample bs = bsdf_sample(*b, mk(wi[0], wi[1], wi[2]), s1, s2x, s2y, true);
    out[0] = bs.wo.x; out[1] = bs.wo.y; out[2] = bs.wo.z; out[3] = bs.pdf; out[4] = bs.eta; out[5] = (float) bs.type;
    out[6] = bs.weight.x; out[7] = bs.weight.y; out[8] = bs.weight.z; out[9] = bs.hf.x; out[10] = bs.hf.y; out[11] = bs.hf.z;
}
uint32_t orc_bsdf_flags(const epsm_bsdf_desc *b) { return bsdf_flags(*b); }
void orc_bsdf_eval_pdf(const epsm_bsdf_desc *b, const float *wi, const float *wo, float *out) {
    V3 v; float p; bsdf_eval_pdf(*b, mk(wi[0], wi[1], wi[2]), mk(wo[0], wo[1], wo[2]), true, &v, &p);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = p;
}
// Emitter-sampling probe from the first hit of ray r: out = [ds.p(3), ds.n(3), ds.d(3), pdf, dist, weight(3)]
void orc_sample_emitter(orc_scene *s, const float *r, float sx, float sy, float *out) {
    g_accel_owner = s;
    Ray ray{mk(r[0], r[1], r[2]), mk(r[4], r[5], r[6]), r[3]};
    PI pi = scene_closest(s->sc, ray); SI si = compute_si(s->sc, ray, pi);
    DS ds; V3 w; sample_emitter_direction(s->sc, si, sx, sy, si.valid, &ds, &w);
    float *o = out; auto put3 = [&](V3 v) { *o++ = v.x; *o++ = v.y; *o++ = v.z; };
    put3(ds.p); put3(ds.n); put3(ds.d); *o++ = ds.pdf; *o++ = ds.dist; put3(w);
}
// pdf_emitter_direction + Emitter::eval probe (scene.cpp:279-290, area.cpp:82-89, :173-200): the ray r starts at the
// reference point; out = [pdf, Le(3)] for the surface it hits (0 when that is no emitter)
void orc_pdf_emitter(orc_scene *s, const float *r, float *out) {
    g_accel_owner = s;
    Ray ray{mk(r[0], r[1], r[2]), mk(r[4], r[5], r[6]), r[3]};
    PI pi = scene_closest(s->sc, ray); SI si = compute_si(s->sc, ray, pi);
    out[0] = si.valid ? pdf_emitter_direction(s->sc, si, ray.o, true) : 0.f;
    V3 le = si.valid ? emitter_eval(s->sc, si) : mk(0, 0, 0);
    out[1] = le.x; out[2] = le.y; out[3] = le.z;
}

void orc_sample_rays(orc_scene *s, uint32_t sensor, uint32_t seed, uint32_t spp, uint64_t lane_begin, uint64_t lane_end, float *out) {
    const Sensor &se = s->sc.sensors[sensor];
    for (uint64_t l = lane_begin; l < lane_end; ++l) {
        PCG32 rng = sampler_for_lane(seed, (uint32_t) l);
        CamRay c = sample_camera_ray(se, (uint32_t) l, spp, rng);
        float *o = out + 14 * (l - lane_begin);
        o[0] = c.ray.o.x; o[1] = c.ray.o.y; o[2] = c.ray.o.z; o[3] = c.ray.d.x; o[4] = c.ray.d.y; o[5] = c.ray.d.z;
        o[6] = c.d_x.x; o[7] = c.d_x.y; o[8] = c.d_x.z; o[9] = c.d_y.x; o[10] = c.d_y.y; o[11] = c.d_y.z; o[12] = c.pos_x; o[13] = c.pos_y;
    }
}
// Camera-ray probe at an explicit film sample position in [0,1]^2 (perspective.cpp:238-279): out = o,d,d_x,d_y,maxt
void orc_camera_ray(orc_scene *s, uint32_t sensor, float u, float v, float *out) {
    const Sensor &se = s->sc.sensors[sensor];
    V3 near_p = xform_point(se.s2c, mk(u, v, 0.f)); V3 d = normalize(near_p);
    V3 o = mk(se.to_world[3], se.to_world[7], se.to_world[11]), dw = xform_vec(se.to_world, d);
    float inv_z = rcp_(d.z), near_t = se.near_clip * inv_z, far_t = se.far_clip * inv_z;
    V3 oo = o + dw * near_t, dx = xform_vec(se.to_world, normalize(near_p + se.dx)), dy = xform_vec(se.to_world, normalize(near_p + se.dy));
    float vals[13] = {oo.x, oo.y, oo.z, dw.x, dw.y, dw.z, dx.x, dx.y, dx.z, dy.x, dy.y, dy.z, far_t - near_t};
    memcpy(out, vals, sizeof(vals));
}

// EPSMIntegrator.render (epsm.py:13-82) for lanes [lane_begin, lane_end): accumulates into film_rgbw (H·W·4).
void orc_render_accumulate(orc_scene *s, uint32_t sensor, uint32_t seed, uint32_t spp, uint32_t max_depth, uint32_t rr_depth,
                           uint64_t lane_begin, uint64_t lane_end, float *film_rgbw, float *per_lane_L) {
    g_accel_owner = s;
    const Scene &sc = s->sc; const Sensor &se = sc.sensors[sensor];
    size_t n = lane_end - lane_begin;
    std::vector<float> Ls(n * 5);
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t) n; ++i) {
        uint32_t lane = (uint32_t) (lane_begin + i);
        PCG32 rng = sampler_for_lane(seed, lane);
        CamRay c = sample_camera_ray(se, lane, spp, rng);
        bool valid; V3 L = sample_primal(sc, c.ray, rng, max_depth, rr_depth, &valid);
        Ls[5 * i] = L.x; Ls[5 * i + 1] = L.y; Ls[5 * i + 2] = L.z; Ls[5 * i + 3] = c.pos_x; Ls[5 * i + 4] = c.pos_y;
    }
    std::vector<float> film((size_t) se.fw * se.fh * 4, 0.f);
    for (size_t i = 0; i < n; ++i) film_put(sc, film, se.fw, se.fh, Ls[5 * i + 3], Ls[5 * i + 4], mk(Ls[5 * i], Ls[5 * i + 1], Ls[5 * i + 2]));
    for (size_t i = 0; i < film.size(); ++i) film_rgbw[i] += film[i];
    if (per_lane_L) for (size_t i = 0; i < n; ++i) { per_lane_L[3 * i] = Ls[5 * i]; per_lane_L[3 * i + 1] = Ls[5 * i + 1]; per_lane_L[3 * i + 2] = Ls[5 * i + 2]; }
}
// HDRFilm::develop + epsm.py:76-82
void orc_film_develop(const float *film_rgbw, uint64_t n_pixels, float *out_hw5) {
    for (uint64_t i = 0; i < n_pixels; ++i) {
        float w = film_rgbw[4 * i + 3]; float d = w == 0.f ? 1.f : w;
        for (int k = 0; k < 3; ++k) out_hw5[5 * i + k] = film_rgbw[4 * i + k] / d;
        out_hw5[5 * i + 3] = 0.f; out_hw5[5 * i + 4] = 0.f;
    }
}
void orc_film_put(orc_scene *s, float *film_rgbw, uint32_t W, uint32_t H, float pos_x, float pos_y, const float *rgb) {
    std::vector<float> f(film_rgbw, film_rgbw + (size_t) W * H * 4);
    film_put(s->sc, f, W, H, pos_x, pos_y, mk(rgb[0], rgb[1], rgb[2]));
    memcpy(film_rgbw, f.data(), f.size() * 4);
}

// sample_path with log_path=True, plus dlduv/dldp1 (epsm.py:238-272) and the per-bounce quantities the
// Backward-mode replay uses (epsm.py:559-562, 609-627, 644-645).  grad_in: H·W·5 for this sensor.
void orc_path_log(orc_scene *s, uint32_t sensor, uint32_t seed, uint32_t spp, uint32_t max_depth_prop, uint32_t rr_depth,
                  const float *grad_in, uint64_t lane_begin, uint64_t lane_end, orc_path *out) {
    g_accel_owner = s;
    const Scene &sc = s->sc; const Sensor &se = sc.sensors[sensor];
    const uint32_t max_iter = std::min(max_depth_prop, 6u);                        // epsm.py:549
    const int prune_variant = g_prune_variant;
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t li = 0; li < (int64_t) (lane_end - lane_begin); ++li) {
        uint32_t lane = (uint32_t) (lane_begin + li);
        orc_path &P = out[li]; memset(&P, 0, sizeof(P));
        PCG32 rng = sampler_for_lane(seed, lane);
        CamRay c = sample_camera_ray(se, lane, spp, rng);
        Ray ray = c.ray;
        P.cam[0] = ray.o.x; P.cam[1] = ray.o.y; P.cam[2] = ray.o.z; P.d[0] = ray.d.x; P.d[1] = ray.d.y; P.d[2] = ray.d.z;
        // (2) pixel motion → direction tangent, epsm.py:250-257
        static const float g_zero[5] = {0, 0, 0, 0, 0};
        const int gfx = (int) c.px - se.border, gfy = (int) c.py - se.border;
        const float *g = (gfx >= 0 && gfy >= 0 && gfx < (int) se.fw && gfy < (int) se.fh) ? grad_in + ((size_t) gfy * se.fw + gfx) * 5 : g_zero;
        V3 grad_d = (c.d_x - ray.d) * g[3] + (c.d_y - ray.d) * g[4];
        uint32_t depth = 0; V3 L = mk(0, 0, 0), beta = mk(1, 1, 1); float eta = 1.f; bool active = true;
        V3 prev_p = mk(0, 0, 0); float prev_pdf = 1.f; bool prev_delta = true;
        bool chain = true;
        for (uint32_t it = 0; it < max_iter; ++it) {
            PI pi; SI si;
            if (active) { pi = scene_closest(sc, ray); si = compute_si(sc, ray, pi); }
            else { pi = PI{false, INFINITY, 0, 0, 0xffffffffu}; si = compute_si(sc, ray, pi); }
            if (it == 0 && si.valid && si.ismesh) {
                // forward-mode derivative of Möller–Trumbore w.r.t. ray.d (triangle detached), epsm.py:263-270
                V3 e1 = si.p1 - si.p0, e2 = si.p2 - si.p0, pvec = cross(ray.d, e2), tvec = ray.o - si.p0, qvec = cross(tvec, e1);
                float det = dot(e1, pvec), inv_det = 1.f / det;
                V3 dpvec = cross(grad_d, e2);
                float ddet = dot(e1, dpvec), dinv = -ddet * inv_det * inv_det;
                float du = dot(tvec, dpvec) * inv_det + dot(tvec, pvec) * dinv;
                float dv = dot(grad_d, qvec) * inv_det + dot(ray.d, qvec) * dinv;
                P.dlduv[0] = -(du + dv); P.dlduv[1] = du;                            // b0 = 1-u-v, b1 = u
                V3 dp = e1 * du + e2 * dv;
                P.dldp1[0] = dp.x; P.dldp1[1] = dp.y; P.dldp1[2] = dp.z;
            } else if (it == 0 && si.valid) {
                // tessellated analytic shape: si.b0/b1 are constant zero; si.p still moves with the ray
                const uint32_t *f = si.fi; V3 p0 = sc.pos[f[0]], p1 = sc.pos[f[1]], p2 = sc.pos[f[2]];
                V3 e1 = p1 - p0, e2 = p2 - p0, pvec = cross(ray.d, e2), tvec = ray.o - p0, qvec = cross(tvec, e1);
                float det = dot(e1, pvec), inv_det = 1.f / det;
                V3 dpvec = cross(grad_d, e2);
                float ddet = dot(e1, dpvec), dinv = -ddet * inv_det * inv_det;
                float du = dot(tvec, dpvec) * inv_det + dot(tvec, pvec) * dinv;
                float dv = dot(grad_d, qvec) * inv_det + dot(ray.d, qvec) * dinv;
                V3 dp = e1 * du + e2 * dv;
                P.dldp1[0] = dp.x; P.dldp1[1] = dp.y; P.dldp1[2] = dp.z;
            }
            float em_pdf = pdf_emitter_direction(sc, si, prev_p, !prev_delta);
            float mis = mis_weight(prev_pdf, em_pdf);
            V3 Le = beta * mis * emitter_eval(sc, si);
            bool active_next = (depth + 1 < max_depth_prop) && si.valid;
            epsm_bsdf_desc b_tex; const epsm_bsdf_desc *b = nullptr; if (si.valid) { b_tex = bsdf_at(sc, si); b = &b_tex; }
            uint32_t flags = b ? bsdf_flags(*b) : 0u;
            bool smooth = active_next && (flags & EPSM_BF_SMOOTH);
            bool active_em = smooth;
            float e1 = rng.next_f32(), e2 = rng.next_f32();
            DS ds; V3 em_weight;
            sample_emitter_direction(sc, si, e1, e2, active_em, &ds, &em_weight);
            active_em &= ds.pdf != 0.f;
            V3 Lr_dir = mk(0, 0, 0); V3 dir_to_light = mk(0, 0, 0);
            if (smooth) dir_to_light = normalize(ds.p - si.p);                       // epsm.py:596
            if (active_em) {
                // Backward mode re-evaluates the emitter with the recomputed direction (epsm.py:596-598)
                const epsm_emitter_desc &em = sc.emitters[ds.emitter];
                V3 em_val = dot(dir_to_light, ds.n) < 0.f ? mk(em.radiance[0], em.radiance[1], em.radiance[2]) : mk(0, 0, 0);
                V3 emw = em_val * (1.0f / ds.pdf);
                V3 wo = to_local(si, dir_to_light); V3 val; float bpdf;
                bsdf_eval_pdf(*b, si.wi, wo, true, &val, &bpdf);
                Lr_dir = beta * mis_weight(ds.pdf, bpdf) * val * emw;
            }
            // (discarded) detached BSDF sample, then the kept one: epsm.py:633-643
            rng.next_f32(); rng.next_f32(); rng.next_f32();
            float s1 = rng.next_f32(), s2x = rng.next_f32(), s2y = rng.next_f32();
            BSample bs; memset(&bs, 0, sizeof(bs));
            if (b) bs = bsdf_sample(*b, si.wi, s1, s2x, s2y, active_next);
            if (it < 5) {
                orc_vertex &v = P.v[it]; P.n_vertices = it + 1;
                auto set3 = [](float *d, V3 a) { d[0] = a.x; d[1] = a.y; d[2] = a.z; };
                v.active = (active && si.valid) ? 1.f : 0.f; v.bsdf_flags = flags; v.ismesh = (float) si.ismesh;
                set3(v.light, ds.p); v.active_em = active_em ? 1.f : 0.f;
                set3(v.p0, si.p0); set3(v.p1, si.p1); set3(v.p2, si.p2); set3(v.p, si.p); v.b0 = si.b0; v.b1 = si.b1;
                set3(v.sh_n, si.sh_n); set3(v.n0, si.n0); set3(v.n1, si.n1); set3(v.n2, si.n2); v.eta = bs.eta; set3(v.hf, bs.hf);
                v.prim = si.prim; for (int k = 0; k < 3; ++k) { v.fi[k] = si.fi[k]; v.bw[k] = si.bw[k]; }
                v.lr_dir_sum = Lr_dir.x + Lr_dir.y + Lr_dir.z;
                v.mesh_flags = si.valid ? sc.meshes[si.mesh].flags : 0u; v.has_normals = si.valid ? sc.meshes[si.mesh].has_normals : 0u;
                if (smooth) {
                    Ray rd = spawn_ray(si, dir_to_light);
                    PI pd = scene_closest(sc, rd); SI sd = compute_si(sc, rd, pd);
                    v.direct_any = sd.valid ? 1u : 0u;
                    v.direct_valid = (sd.valid && active_em) ? 1u : 0u;
                    for (int k = 0; k < 3; ++k) { v.direct_fi[k] = sd.fi[k]; v.direct_bw[k] = sd.bw[k]; }
                    if (sd.valid) {
                        float dis = norm(ds.p - sd.p) / norm(ds.p - si.p);
                        v.direct_dis = dis < 0.01f ? 0.f : dis;
                        // flag bit 8: gradient enabled on the hit mesh
                        v.direct_any |= (sc.meshes[sd.mesh].flags & EPSM_MESH_GRAD) ? 0x100u : 0u;
                    }
                }
            }
            L = L + Le + Lr_dir;
            if (prune_variant >= 0) {
                // gradient-liveness pruning, the rule of the product's wf_chain_vertex (epsm_device.cuh): calc_grad masks
                // everything behind the vertex that ends the specular chain, and render_backward consumes nothing else
                const bool diffuse = (flags & EPSM_BF_DIFFUSE) != 0, vact = active && si.valid;
                chain = chain && vact && si.ismesh && (prune_variant == 0 ? !diffuse : (it == 0 ? diffuse : !diffuse));
                if (!chain || it + 1 >= 5) break;                                   // the terminating vertex has been logged
            }
            if (si.valid) ray = spawn_ray(si, to_world(si, bs.wo));
            eta *= bs.eta; beta = beta * bs.weight;
            prev_p = si.p; prev_pdf = bs.pdf; prev_delta = (bs.type & EPSM_BF_DELTA) != 0;
            float bmax = max3(beta);
            active_next &= bmax != 0.f;
            float rr_prob = fminf(bmax * eta * eta, .95f);
            bool rr_active = depth >= rr_depth;
            if (rr_active) beta = beta * rcp_(rr_prob);
            bool rr_continue = rng.next_f32() < rr_prob;
            active_next &= !rr_active || rr_continue;
            if (si.valid) depth += 1;
            active = active_next;
        }
        P.L[0] = L.x; P.L[1] = L.y; P.L[2] = L.z;
    }
}
// The restatement behind the same switch as oracle/ref_probe.cpp (which drives the REFERENCE's header code): function
// numbering, input and output layouts are identical, so tests compare the two call for call.
int orc_probe(int fn, uint64_t n, const float *in_, int in_stride, float *out_, int out_stride) {
    for (uint64_t i = 0; i < n; ++i) {
        const float *in = in_ + i * (uint64_t) in_stride; float *out = out_ + i * (uint64_t) out_stride;
        auto mf = [&]() { return Microfacet{in[2] != 0.f, in[0], in[1], true}; };
        switch (fn) {
        case 0: { uint32_t v0, v1, o0, o1; memcpy(&v0, in, 4); memcpy(&v1, in + 1, 4); tea32(v0, v1, 4, &o0, &o1); memcpy(out, &o0, 4); memcpy(out + 1, &o1, 4); } break;
        case 2: sincos_dr(in[0], &out[0], &out[1]); break;
        case 3: { V3 v = square_to_cosine_hemisphere(in[0], in[1]); out[0] = v.x; out[1] = v.y; out[2] = v.z; } break;
        case 4: { V3 s, t; coordinate_system(mk(in[0], in[1], in[2]), &s, &t); out[0] = s.x; out[1] = s.y; out[2] = s.z; out[3] = t.x; out[4] = t.y; out[5] = t.z; } break;
        case 5: fresnel(in[0], in[1], &out[0], &out[1], &out[2], &out[3]); break;
        case 6: out[0] = fresnel_conductor(in[0], in[1], in[2]); break;
        case 7: { V3 m; mf().sample(in[6], in[7], &m, &out[3]); out[0] = m.x; out[1] = m.y; out[2] = m.z; } break;
        case 8: out[0] = mf().eval(mk(in[3], in[4], in[5])); break;
        case 9: out[0] = mf().smith_g1(mk(in[3], in[4], in[5]), mk(in[6], in[7], in[8])); break;
        case 10: { float t = safe_sqrt(1.f - in[0]); out[0] = 1.f - t; out[1] = t * in[1]; } break;       // as in sample_emitter_direction
        case 11: { V3 r = reflect(mk(in[0], in[1], in[2]), mk(in[3], in[4], in[5])); out[0] = r.x; out[1] = r.y; out[2] = r.z; } break;
        case 12: {
            SI si; si.sh_n = mk(in[0], in[1], in[2]); coordinate_system(si.sh_n, &si.sh_s, &si.sh_t);     // Frame3f(n), frame.h:21-26
            V3 v = mk(in[3], in[4], in[5]), l = to_local(si, v), w = to_world(si, v);
            out[0] = l.x; out[1] = l.y; out[2] = l.z; out[3] = w.x; out[4] = w.y; out[5] = w.z;
        } break;
        default: return -1;
        }
    }
    return 0;
}


// ------------------------------------------------------------------------------------------------
//  prb_reparam (src/python/python/ad/integrators/prb_reparam.py:277-607, reparam.py:10-463, common.py:291-422,
//  803-975): the DETACHED half of the differential pass — the random walk, every sampled quantity and the auxiliary
//  rays of the reparameterisation — logged per lane.  oracle/prb_reparam_ref.py re-evaluates the ATTACHED half on these
//  records with torch autograd.  Record layout: PRB_* offsets below (mirrored in prb_reparam_ref.py).
// ------------------------------------------------------------------------------------------------
enum { PRB_AUX = 8, PRB_NAUX_MAX = 64, PRB_HDR = 16, PRB_IT = 72 };
// aux record: hit, prim, b1, b2, w (= Z_i), d_w_omega(3)                                     reparam.py:83-118
// header:  o(3) d(3) pos(2) L(3) valid n_iter cam_active spare(2)   then n_rays aux records of the camera ray
// iteration: see the writes below; then n_rays aux records of the main ray and n_rays of the emitter ray

// warp.h:546-583 (approach 1.2), κ ≠ 0
static V3 square_to_von_mises_fisher(float sx, float sy_, float kappa) {
    float sy = fmaxf(1.f - sy_, 1e-6f);
    float cos_theta = 1.f + logf(fmaf(1.f - sy, expf(-2.f * kappa), sy)) / kappa;
    float s, c; sincos_dr(6.283185307179586f * sx, &s, &c);
    float sin_theta = safe_sqrt(1.f - cos_theta * cos_theta);
    return mk(c * sin_theta, s * sin_theta, cos_theta);
}
// Mesh::compute_surface_interaction with RayFlags::BoundaryTest (mesh.cpp:823-886) / Rectangle (rectangle.cpp:319-320)
static float boundary_test(const Scene &sc, const PI &pi, V3 ray_d) {
    uint32_t mid = sc.face_mesh[pi.prim]; const Mesh &m = sc.meshes[mid];
    const uint32_t *f = &sc.faces[3 * pi.prim];
    V3 p0 = sc.pos[f[0]], p1 = sc.pos[f[1]], p2 = sc.pos[f[2]];
    float b1 = pi.u, b2 = pi.v, b0 = 1.f - b1 - b2;
    V3 p = fmadd(p0, b0, fmadd(p1, b1, p2 * b2));
    if (m.flags & EPSM_MESH_RECTANGLE) {
        V3 c0 = sc.pos[m.v_off], ex = sc.pos[m.v_off + 1] - c0, ey = sc.pos[m.v_off + 3] - c0, r = p - c0;
        float u = dot(r, ex) / dot(ex, ex), v = dot(r, ey) / dot(ey, ey);
        return fminf(0.5f - fabsf(u - 0.5f), 0.5f - fabsf(v - 0.5f));
    }
    V3 dp0 = p1 - p0, dp1 = p2 - p0, rel = p - p0;
    float bb1 = dot(dp0, rel), bb2 = dot(dp1, rel), a11 = dot(dp0, dp0), a12 = dot(dp0, dp1), a22 = dot(dp1, dp1);
    float inv_det = rcp_(a11 * a22 - a12 * a12);
    float u = fmaf(a22, bb1, -(a12 * bb2)) * inv_det, v = fmaf(-a12, bb1, a11 * bb2) * inv_det, w = 1.f - u - v;
    if (!m.has_normals) {
        const float h = 0.5f * sqrtf(3.f);
        float px = 0.f * w + 1.f * u + 0.5f * v, py = 0.f * w + 0.f * u + h * v;
        const float tx[3] = {0.f, 1.f, 0.5f}, ty[3] = {0.f, 0.f, h};
        const float ex[3] = {tx[1] - tx[0], tx[2] - tx[1], tx[0] - tx[2]}, ey[3] = {ty[1] - ty[0], ty[2] - ty[1], ty[0] - ty[2]};
        float s = (ex[0] * ey[2] - ey[0] * ex[2]) < 0.f ? -1.f : 1.f;
        float dx = INFINITY, dy = INFINITY;
        for (int k = 0; k < 3; ++k) {
            float vx = px - tx[k], vy = py - ty[k];
            float tt = fminf(fmaxf((vx * ex[k] + vy * ey[k]) / (ex[k] * ex[k] + ey[k] * ey[k]), 0.f), 1.f);
            float qx = vx - ex[k] * tt, qy = vy - ey[k] * tt;
            dx = fminf(dx, qx * qx + qy * qy); dy = fminf(dy, s * (vx * ey[k] - vy * ex[k]));
        }
        return sqrtf(dx) / (sqrtf(3.f) / 6.f);
    }
    V3 n = fmadd(sc.nrm[f[0]], w, fmadd(sc.nrm[f[1]], u, sc.nrm[f[2]] * v));
    float dp = dot(n, -ray_d);
    return dp * dp;
}
// One auxiliary ray of _sample_warp_field (reparam.py:83-118), detached quantities only.
static void warp_field_sample(const Scene &sc, const Ray &ray, float sx, float sy, float kappa, float exponent, float *rec) {
    V3 om = square_to_von_mises_fisher(sx, sy, kappa);
    V3 fs, ft; coordinate_system(ray.d, &fs, &ft);                                 // Frame3f(ray.d)
    V3 ad = fmadd(ray.d, om.z, fmadd(ft, om.y, fs * om.x));                         // frame.to_world
    Ray aux{ray.o, ad, FLT_MAX};
    PI pi = scene_closest(sc, aux);
    float B = pi.valid ? boundary_test(sc, pi, ad) : 1.0f;
    float inv_vmf = rcp_(fmaf(sy, expf(-2.f * kappa), 1.f - sy));
    float w_denom = inv_vmf - 1.f + B;
    float w_denom_rcp = w_denom > 1e-4f ? rcp_(w_denom) : 0.f;
    float w = powf(w_denom_rcp, exponent) * inv_vmf;
    float tmp1 = inv_vmf * w * w_denom_rcp * kappa * exponent;
    tmp1 = fminf(fmaxf(tmp1, -1e10f), 1e10f);
    V3 tmp2 = fmadd(ray.d, 0.f, fmadd(ft, om.y, fs * om.x));
    rec[0] = pi.valid ? 1.f : 0.f; rec[1] = pi.valid ? (float) pi.prim : 0.f; rec[2] = pi.u; rec[3] = pi.v; rec[4] = w;
    rec[5] = tmp1 * tmp2.x; rec[6] = tmp1 * tmp2.y; rec[7] = tmp1 * tmp2.z;
}
static void reparam_log(const Scene &sc, const Ray &ray, bool active, PCG32 &rng, uint32_t n_rays, float kappa, float exponent, float *out) {
    memset(out, 0, sizeof(float) * PRB_AUX * n_rays);
    if (!active) return;                                                            // masked lanes do not advance the generator
    for (uint32_t i = 0; i < n_rays; ++i) {
        float sx = rng.next_f32(), sy = rng.next_f32();
        warp_field_sample(sc, ray, sx, sy, kappa, exponent, out + PRB_AUX * i);
    }
}
uint32_t orc_prb_record_floats(uint32_t n_iter, uint32_t n_rays) { return PRB_HDR + PRB_AUX * n_rays + n_iter * (PRB_IT + 2 * PRB_AUX * n_rays); }

// Lanes [lane_begin, lane_end) of RBIntegrator.render_backward's two sample() passes for `prb_reparam`.  n_iter = record
// capacity in iterations (≥ min(max_depth, what the test expects)); deeper iterations are still walked but not logged.
void orc_prb_log(orc_scene *os, uint32_t sensor, uint32_t seed, uint32_t spp, uint32_t max_depth, uint32_t rr_depth,
                 uint32_t reparam_max_depth, uint32_t n_rays, float kappa, float exponent,
                 uint64_t lane_begin, uint64_t lane_end, uint32_t n_iter, float *out) {
    const Scene &sc = os->sc; const Sensor &sn = sc.sensors[sensor];
    g_accel_owner = os;
    const uint32_t stride = orc_prb_record_floats(n_iter, n_rays);
    #pragma omp parallel for schedule(dynamic, 16)
    for (int64_t lane_ = (int64_t) lane_begin; lane_ < (int64_t) lane_end; ++lane_) {
        uint32_t lane = (uint32_t) lane_;
        float *R = out + (uint64_t) (lane - lane_begin) * stride; memset(R, 0, sizeof(float) * stride);
        PCG32 rng = sampler_for_lane(seed, lane);
        PCG32 aux = sampler_for_lane(0xffffffffu ^ seed, lane);                      // _ReparamWrapper, common.py:1200-1203
        CamRay cr = sample_camera_ray(sn, lane, spp, rng);
        // primal pass with the cloned sampler (common.py:905-915)
        bool valid; PCG32 rng_p = rng;
        V3 Ltot = sample_primal(sc, cr.ray, rng_p, max_depth, rr_depth, &valid);
        R[0] = cr.ray.o.x; R[1] = cr.ray.o.y; R[2] = cr.ray.o.z; R[3] = cr.ray.d.x; R[4] = cr.ray.d.y; R[5] = cr.ray.d.z;
        R[6] = cr.pos_x; R[7] = cr.pos_y; R[8] = Ltot.x; R[9] = Ltot.y; R[10] = Ltot.z; R[11] = valid ? 1.f : 0.f;
        // sample_rays' reparameterisation of the camera ray (depth 0), common.py:372-397
        bool cam_active = 0 < reparam_max_depth;
        R[13] = cam_active ? 1.f : 0.f;
        Ray cam_aux{cr.ray.o, cr.ray.d, FLT_MAX};
        reparam_log(sc, cam_aux, cam_active, aux, n_rays, kappa, exponent, R + PRB_HDR);
        // differential pass, prb_reparam.py:329-602 (detached quantities)
        uint32_t depth = 0; V3 L = Ltot, beta = mk(1, 1, 1); float eta = 1.f, mis_em = 1.f; bool active = true;
        Ray ray = cr.ray; PI pi = scene_closest(sc, ray);
        uint32_t it = 0;
        for (; it < max_depth && active; ++it) {
            float *I = it < n_iter ? R + PRB_HDR + PRB_AUX * n_rays + (uint64_t) it * (PRB_IT + 2 * PRB_AUX * n_rays) : nullptr;
            SI si = compute_si(sc, ray, pi);
            V3 Le = beta * mis_em * emitter_eval(sc, si);
            bool active_next = (depth + 1 < max_depth) && si.valid;
            epsm_bsdf_desc b_tex; const epsm_bsdf_desc *b = nullptr; if (si.valid) { b_tex = bsdf_at(sc, si); b = &b_tex; }
            uint32_t flags = b ? bsdf_flags(*b) : 0u;
            bool active_em = active_next && (flags & EPSM_BF_SMOOTH);
            float e1 = rng.next_f32(), e2 = rng.next_f32();
            DS ds; V3 em_weight;
            sample_emitter_direction(sc, si, e1, e2, active_em, &ds, &em_weight);
            active_em &= ds.pdf != 0.f;
            Ray em_ray{mk(0, 0, 0), mk(0, 0, 1), 0.f};
            V3 Lr_dir = mk(0, 0, 0); float mis_direct = 0.f;
            if (active_em) {
                em_ray = spawn_ray_to(si, ds.p);
                V3 wo = to_local(si, ds.d); V3 val; float bpdf;
                bsdf_eval_pdf(*b, si.wi, wo, true, &val, &bpdf);
                mis_direct = mis_weight(ds.pdf, bpdf);
                Lr_dir = beta * mis_direct * val * em_weight;
            }
            float s1 = rng.next_f32(), s2x = rng.next_f32(), s2y = rng.next_f32();
            BSample bs; memset(&bs, 0, sizeof(bs));
            if (b) bs = bsdf_sample(*b, si.wi, s1, s2x, s2y, active_next);
            V3 beta_in = beta; float eta_in = eta;
            eta *= bs.eta; beta = beta * bs.weight;
            V3 L_prev = L; L = L - Le - Lr_dir;
            float bmax = max3(beta);
            active_next &= bmax != 0.f;
            float rr_prob = fminf(bmax * eta * eta, .95f);
            bool rr_active = depth >= rr_depth;
            if (rr_active) beta = beta * rcp_(rr_prob);
            bool rr_continue = rng.next_f32() < rr_prob;
            active_next &= !rr_active || rr_continue;
            Ray ray_next = ray; PI pi_next; memset(&pi_next, 0, sizeof(pi_next));
            SI si_next; memset(&si_next, 0, sizeof(si_next));
            if (si.valid) ray_next = spawn_ray(si, to_world(si, bs.wo));
            if (active_next) pi_next = scene_closest(sc, ray_next);                  // ray_intersect_preliminary(ray_next, active=active_next)
            si_next = compute_si(sc, ray_next, pi_next);
            float pdf_em = pdf_emitter_direction(sc, si_next, si.p, !(bs.type & EPSM_BF_DELTA));
            float mis_next = mis_weight(bs.pdf, pdf_em);
            // run-ahead with a clone of the sampler, prb_reparam.py:484-520
            PCG32 clone = rng;
            bool active_next_next = active_next && si_next.valid && (depth + 2 < max_depth);
            epsm_bsdf_desc bn_tex; const epsm_bsdf_desc *bn = nullptr; if (si_next.valid) { bn_tex = bsdf_at(sc, si_next); bn = &bn_tex; }
            bool active_em_next = active_next_next && bn && (bsdf_flags(*bn) & EPSM_BF_SMOOTH);
            float c1 = clone.next_f32(), c2 = clone.next_f32();
            DS dsn; V3 emw_n;
            sample_emitter_direction(sc, si_next, c1, c2, active_em_next, &dsn, &emw_n);
            active_em_next &= dsn.pdf != 0.f;
            V3 Lr_dir_next = mk(0, 0, 0);
            if (active_em_next) {
                V3 won = to_local(si_next, dsn.d), valn; float bpdfn;
                bsdf_eval_pdf(*bn, si_next.wi, won, true, &valn, &bpdfn);
                Lr_dir_next = beta * mis_weight(dsn.pdf, bpdfn) * valn * emw_n;
            }
            float c3 = clone.next_f32(), c4 = clone.next_f32(), c5 = clone.next_f32();
            BSample bsn; memset(&bsn, 0, sizeof(bsn));
            if (bn) bsn = bsdf_sample(*bn, si_next.wi, c3, c4, c5, active_next_next);
            V3 Le_next = active_next ? beta * mis_next * emitter_eval(sc, si_next) : mk(0, 0, 0);
            V3 L_next = L - Le_next - Lr_dir_next;
            // reparameterisations: the emitter ray's custom op is created later in the iteration, so its backward runs first
            // (reverse traversal order of the AD graph); both draw from the wrapper's one generator.
            bool em_rp_active = active_em && (depth + 1 < reparam_max_depth);
            bool main_rp_active = depth < reparam_max_depth;
            float *aux_main = I ? I + PRB_IT : nullptr, *aux_em = I ? I + PRB_IT + PRB_AUX * n_rays : nullptr;
            std::vector<float> scratch(PRB_AUX * n_rays);
            reparam_log(sc, Ray{em_ray.o, em_ray.d, FLT_MAX}, em_rp_active, aux, n_rays, kappa, exponent, aux_em ? aux_em : scratch.data());
            reparam_log(sc, Ray{ray.o, ray.d, FLT_MAX}, main_rp_active, aux, n_rays, kappa, exponent, aux_main ? aux_main : scratch.data());
            if (I) {
                I[0] = si.valid ? 1.f : 0.f; I[1] = si.valid ? (float) pi.prim : 0.f; I[2] = pi.u; I[3] = pi.v; I[4] = pi.t;
                I[5] = ray.o.x; I[6] = ray.o.y; I[7] = ray.o.z; I[8] = ray.d.x; I[9] = ray.d.y; I[10] = ray.d.z;
                I[11] = beta_in.x; I[12] = beta_in.y; I[13] = beta_in.z; I[14] = mis_em; I[15] = eta_in;
                I[16] = active_next ? 1.f : 0.f; I[17] = active_em ? 1.f : 0.f;
                I[18] = ds.p.x; I[19] = ds.p.y; I[20] = ds.p.z; I[21] = ds.n.x; I[22] = ds.n.y; I[23] = ds.n.z; I[24] = ds.pdf;
                I[25] = em_weight.x; I[26] = em_weight.y; I[27] = em_weight.z; I[28] = mis_direct;
                I[29] = bs.wo.x; I[30] = bs.wo.y; I[31] = bs.wo.z; I[32] = bs.pdf; I[33] = bs.weight.x; I[34] = bs.weight.y; I[35] = bs.weight.z;
                I[36] = (bs.type & EPSM_BF_DELTA) ? 1.f : 0.f; I[37] = ray_next.d.x; I[38] = ray_next.d.y; I[39] = ray_next.d.z;
                I[40] = L_prev.x; I[41] = L_prev.y; I[42] = L_prev.z; I[43] = L.x; I[44] = L.y; I[45] = L.z;
                I[46] = si_next.valid ? 1.f : 0.f; I[47] = si_next.valid ? (float) pi_next.prim : 0.f; I[48] = pi_next.u; I[49] = pi_next.v;
                I[50] = L_next.x; I[51] = L_next.y; I[52] = L_next.z; I[53] = bsn.wo.x; I[54] = bsn.wo.y; I[55] = bsn.wo.z;
                I[56] = active_next_next ? 1.f : 0.f; I[57] = main_rp_active ? 1.f : 0.f; I[58] = em_rp_active ? 1.f : 0.f;
                I[59] = em_ray.o.x; I[60] = em_ray.o.y; I[61] = em_ray.o.z; I[62] = em_ray.d.x; I[63] = em_ray.d.y; I[64] = em_ray.d.z;
                I[65] = Le.x; I[66] = Le.y; I[67] = Le.z; I[68] = Lr_dir.x; I[69] = Lr_dir.y; I[70] = Lr_dir.z; I[71] = (float) depth;
            }
            mis_em = mis_next;
            pi = pi_next; ray = ray_next;
            if (si.valid) depth += 1;
            active = active_next;
        }
        R[12] = (float) it;
    }
}

// CPU-baseline timing only: replay each path up to the vertex that ends its gradient chain (same rule as the product), so
// that the CPU arm and the GPU arm of bench.py do equal work.  P.L is then the radiance of the prefix only.
void orc_set_prune(int variant) { g_prune_variant = variant; }

uint32_t orc_sizeof_path() { return (uint32_t) sizeof(orc_path); }
uint32_t orc_sizeof_vertex() { return (uint32_t) sizeof(orc_vertex); }

} // extern "C"
