// epsm_device.cuh — per-lane device code of the EPSM manifold path for sm_100a.
//
// Everything here is __host__ __device__ so that tests/hostsim (a test-only harness) can run the
// very same lane code on the CPU while developing without a GPU; the product only ever launches it
// from the kernels in epsm_kernels.cu.  Compiled with -fmad=false: the only fused multiply-adds are
// the explicit fmaf() calls, which follow Dr.Jit's array code (dot/cross/normalize) so that the
// arithmetic is the reference's (ext/drjit/include/drjit/array_base.h:603-628, array_router.h:635-666).
//
// Reference map (paths relative to the reference root):
//   RNG            include/mitsuba/core/random.h:76-90, ext/drjit/include/drjit/random.h:55-132, src/render/sampler.cpp:115-130
//   ray generation src/python/python/ad/integrators/common.py:291-422, src/sensors/perspective.cpp:238-279
//   triangle test  include/mitsuba/render/mesh.h:342-367
//   surface hit    src/render/mesh.cpp:632-892, include/mitsuba/render/interaction.h:136-165,262-275,482-505
//   BSDFs          src/bsdfs/{diffuse,twosided,dielectric,conductor,roughconductor}.cpp, include/mitsuba/render/{fresnel,microfacet}.h
//   emitters       src/render/scene.cpp:161-290, src/emitters/area.cpp:82-160, src/render/shape.cpp:352-383, src/render/mesh.cpp:516-571
//   path loops     src/python/python/ad/integrators/epsm.py:308-501 (sample), :503-742 (sample_path)
//   manifold solve epsm.py:745-946 (manifold), :952-1200 (manifold_caustic)
//   film           src/render/imageblock.cpp:414-494, src/rfilters/gaussian.cpp:48-96
#pragma once
#include <math.h>
#include <stddef.h>
#include <float.h>
#include <string.h>
#include "epsm_types.h"
#include "../../include/epsm_b200.h"

// ------------------------------------------------------------------------------------------------
//  Scalar / vector helpers
// ------------------------------------------------------------------------------------------------
EPSM_HD uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
EPSM_HD float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
EPSM_HD f4 ldg4(const f4 *p) {
#ifdef __CUDA_ARCH__
    float4 v = __ldg(reinterpret_cast<const float4 *>(p));     // ld.global.nc.v4.f32 (128-bit)
    return f4{v.x, v.y, v.z, v.w};
#else
    return *p;
#endif
}
// 256-bit read-only load (sm_100a: LDG.E.256.CONSTANT — one L1 data-pipe wavefront per thread instead of two);
// `p` must be 32-byte aligned.
struct f8 { f4 a, b; };
EPSM_HD f8 ldg8(const f4 *p) {
#ifdef __CUDA_ARCH__
    f8 r;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
    return r;
#else
    return f8{p[0], p[1]};
#endif
}
struct f2 { float x, y; };
EPSM_HD f2 ldg2(const f4 *p) {
#ifdef __CUDA_ARCH__
    float2 v = __ldg(reinterpret_cast<const float2 *>(p));
    return f2{v.x, v.y};
#else
    return f2{p->x, p->y};
#endif
}
EPSM_HD f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
EPSM_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
EPSM_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
EPSM_HD f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
EPSM_HD f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
EPSM_HD f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
EPSM_HD float dot3(f3 a, f3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
EPSM_HD f3 cross3(f3 a, f3 b) {
    return mk3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
EPSM_HD float sqn3(f3 a) { return fmaf(a.z, a.z, fmaf(a.y, a.y, a.x * a.x)); }
EPSM_HD float rsqrt_e(float x) { return 1.0f / sqrtf(x); }
EPSM_HD float rcp_e(float x) { return 1.0f / x; }
EPSM_HD f3 norm3(f3 a) { return a * rsqrt_e(sqn3(a)); }
EPSM_HD float len3(f3 a) { return sqrtf(sqn3(a)); }
EPSM_HD f3 fma3(f3 a, float b, f3 c) { return mk3(fmaf(a.x, b, c.x), fmaf(a.y, b, c.y), fmaf(a.z, b, c.z)); }
EPSM_HD float mulsign(float a, float b) { return u2f(f2u(a) ^ (f2u(b) & 0x80000000u)); }
EPSM_HD float mulsign_neg(float a, float b) { return u2f(f2u(a) ^ (~f2u(b) & 0x80000000u)); }
EPSM_HD float safe_sqrt(float x) { return sqrtf(fmaxf(x, 0.0f)); }
EPSM_HD float sqr(float x) { return x * x; }
EPSM_HD float max3(f3 a) { return fmaxf(fmaxf(a.x, a.y), a.z); }
// Three-input min / max: one FMNMX3 on sm_100a (the slab tests are ALU-pipe heavy); same value as the nested form.
EPSM_HD float fmax3(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
#else
    return fmaxf(fmaxf(a, b), c);
#endif
}
EPSM_HD float fmin3(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
#else
    return fminf(fminf(a, b), c);
#endif
}
EPSM_HD f3 ld3(const float *p, uint32_t i) { return mk3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
EPSM_HD f3 a3(const float *p) { return mk3(p[0], p[1], p[2]); }

#define EPSM_PI      3.14159265358979323846f
#define EPSM_INV_PI  0.31830988618379067154f
#define EPSM_RAY_EPS (1500.0f * 5.9604644775390625e-8f)      // include/mitsuba/core/math.h:18-22
#define EPSM_SHADOW_EPS (EPSM_RAY_EPS * 10.0f)

// ------------------------------------------------------------------------------------------------
//  RNG: TEA-seeded PCG32 stream per lane (integer pipes; bit-exact contract)
// ------------------------------------------------------------------------------------------------
struct Pcg { uint64_t state, inc; };
EPSM_HD void tea4(uint32_t v0, uint32_t v1, uint32_t &o0, uint32_t &o1) {
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    o0 = v0; o1 = v1;
}
EPSM_HD uint32_t pcg_next(Pcg &r) {
    uint64_t old = r.state;
    r.state = old * 0x5851f42d4c957f2dULL + r.inc;
    uint32_t xs = (uint32_t) (((old >> 18u) ^ old) >> 27u), rot = (uint32_t) (old >> 59u);
    return (xs >> rot) | (xs << ((0u - rot) & 31u));
}
EPSM_HD Pcg pcg_for_lane(uint32_t seed, uint32_t lane) {
    uint32_t v0, v1; tea4(seed, lane, v0, v1);
    Pcg r; r.state = 0; r.inc = ((uint64_t) v1 << 1) | 1u;
    pcg_next(r); r.state += (uint64_t) v0; pcg_next(r);
    return r;
}
EPSM_HD float pcg_f32(Pcg &r) { return u2f((pcg_next(r) >> 9) | 0x3f800000u) - 1.0f; }

// ------------------------------------------------------------------------------------------------
//  Dr.Jit single-precision sincos (CEPHES polynomials, ext/drjit/include/drjit/math.h:76-186)
// ------------------------------------------------------------------------------------------------
EPSM_HD void sincos_dr(float x, float &so, float &co) {
    float xa = fabsf(x);
    int32_t j = (int32_t) (xa * 1.2732395447351626862f);
    j = (j + 1) & ~1;
    float y = (float) j;
    uint32_t sign_sin = ((uint32_t) j << 29) ^ f2u(x);
    uint32_t sign_cos = (uint32_t) (~(j - 2)) << 29;
    y = xa - y * 0.78515625f - y * 2.4187564849853515625e-4f - y * 3.77489497744594108e-8f;
    float z = y * y, z2 = z * z;
    float s = fmaf(z2, -1.9515295891e-4f, fmaf(z, 8.3321608736e-3f, -1.6666654611e-1f)) * z;
    float c = fmaf(z2, 2.443315711809948e-5f, fmaf(z, -1.388731625493765e-3f, 4.166664568298827e-2f)) * z;
    s = fmaf(s, y, y);
    c = fmaf(c, z, fmaf(z, -0.5f, 1.0f));
    bool poly = (j & 2) == 0;
    so = u2f(f2u(poly ? s : c) ^ (sign_sin & 0x80000000u));
    co = u2f(f2u(poly ? c : s) ^ (sign_cos & 0x80000000u));
}

// include/mitsuba/core/warp.h:54-95, 320-328
EPSM_HD f3 square_to_cosine_hemisphere(float sx, float sy) {
    float x = fmaf(2.0f, sx, -1.0f), y = fmaf(2.0f, sy, -1.0f);
    bool is_zero = (x == 0.0f) && (y == 0.0f), q13 = fabsf(x) < fabsf(y);
    float r = q13 ? y : x, rp = q13 ? x : y;
    float phi = 0.25f * EPSM_PI * rp / r;
    if (q13) phi = 0.5f * EPSM_PI - phi;
    if (is_zero) phi = 0.0f;
    float s, c; sincos_dr(phi, s, c);
    float px = r * c, py = r * s;
    return mk3(px, py, safe_sqrt(1.0f - fmaf(py, py, px * px)));
}
// include/mitsuba/core/vector.h:116-136
EPSM_HD void coordinate_system(f3 n, f3 &s, f3 &t) {
    float sign = (f2u(n.z) & 0x80000000u) ? -1.0f : 1.0f;
    float a = -rcp_e(sign + n.z), b = n.x * n.y * a;
    s = mk3(mulsign(sqr(n.x) * a, n.z) + 1.0f, mulsign(b, n.z), mulsign_neg(n.x, n.z));
    t = mk3(b, fmaf(n.y, n.y * a, sign), -n.y);
}
// drjit estrin scheme for 10 coefficients (math.h:26-42), used by GaussianFilter::eval (gaussian.cpp:94-96)
EPSM_HD float filter_eval(const float *c, float xx) {
    float x = xx * xx, x2 = x * x, x4 = x2 * x2, x8 = x4 * x4;
    float r0 = fmaf(x, c[1], c[0]), r1 = fmaf(x, c[3], c[2]), r2 = fmaf(x, c[5], c[4]),
          r3 = fmaf(x, c[7], c[6]), r4 = fmaf(x, c[9], c[8]);
    float q0 = fmaf(x2, r1, r0), q1 = fmaf(x2, r3, r2);
    float p0 = fmaf(x4, q1, q0);
    return fmaxf(fmaf(x8, r4, p0), 0.0f);
}

// ------------------------------------------------------------------------------------------------
//  Ray / triangle / BVH4 traversal
// ------------------------------------------------------------------------------------------------
struct Ray { f3 o, d; float maxt; };
struct Hit { float t, u, v; uint32_t prim; };         // prim == 0xffffffff: miss

// Möller–Trumbore with the reference's expression order (mesh.h:347-364).  e1, e2 come precomputed.
EPSM_HD bool tri_test(const Ray &ray, f3 p0, f3 e1, f3 e2, float maxt, float &t, float &u, float &v) {
    f3 pvec = cross3(ray.d, e2);
    float inv_det = rcp_e(dot3(e1, pvec));
    f3 tvec = ray.o - p0;
    u = dot3(tvec, pvec) * inv_det;
    bool active = u >= 0.f && u <= 1.f;
    f3 qvec = cross3(tvec, e1);
    v = dot3(ray.d, qvec) * inv_det;
    active &= v >= 0.f && u + v <= 1.f;
    t = dot3(e2, qvec) * inv_det;
    active &= t >= 0.f && t <= maxt;
    return active;
}

// Per-thread traversal stack: a short stack in shared memory (entry i of this thread lives at
// s_ref[i * stride]; consecutive threads are consecutive words ⇒ conflict-free) that spills into a
// thread-local array.
#define EPSM_SSTACK 12
#define EPSM_LSTACK 84
struct TStack {
    uint32_t *s_ref; float *s_t; int stride;
    uint32_t l_ref[EPSM_LSTACK]; float l_t[EPSM_LSTACK];
    int sp;
    EPSM_HD void push(uint32_t ref, float t) {
        if (sp < EPSM_SSTACK) { s_ref[sp * stride] = ref; s_t[sp * stride] = t; }
        else if (sp < EPSM_SSTACK + EPSM_LSTACK) { l_ref[sp - EPSM_SSTACK] = ref; l_t[sp - EPSM_SSTACK] = t; }
        else return;                                   // deeper than any BVH4 this builder produces
        ++sp;
    }
    EPSM_HD void pop(uint32_t &ref, float &t) {
        --sp;
        if (sp < EPSM_SSTACK) { ref = s_ref[sp * stride]; t = s_t[sp * stride]; }
        else { ref = l_ref[sp - EPSM_SSTACK]; t = l_t[sp - EPSM_SSTACK]; }
    }
};

EPSM_HD void cswap(uint32_t &a, uint32_t &b) { uint32_t lo = a < b ? a : b, hi = a < b ? b : a; a = lo; b = hi; }

// ------------------------------------------------------------------------------------------------
//  BVH4 refit after a vertex update (the per-iteration scene update of EPSM/optim.py:110-123 →
//  Mesh::parameters_changed, mesh.cpp:82-102 → Scene::parameters_changed, scene.cpp:304-336, which rebuilds
//  the Embree / OptiX structure).  The topology of the last host build is kept; triangle records and boxes
//  are recomputed with the host builder's expressions (epsm_bvh.cpp), so a refit with unchanged positions
//  reproduces the build bit for bit.  Shared by k_refit_* (epsm_kernels.cu) and the test-only host harness.
// ------------------------------------------------------------------------------------------------
// Rewrites triangle slot `slot` (p0 | e1 | e2, prim id kept); returns max |coordinate| of its vertices.
EPSM_HD float refit_tri_slot(f4 *tris, const uint32_t *faces, const float *pos, uint32_t slot) {
    f4 *tp = tris + (size_t) slot * EPSM_TRI_F4;
    const float pf = tp[0].w; const uint32_t prim = f2u(pf);
    const uint32_t *f = faces + 3 * (size_t) prim;
    const f3 p0 = ld3(pos, f[0]), p1 = ld3(pos, f[1]), p2 = ld3(pos, f[2]);
    tp[0] = f4{p0.x, p0.y, p0.z, pf};
    tp[1] = f4{p1.x - p0.x, p1.y - p0.y, p1.z - p0.z, p2.x - p0.x};
    tp[2] = f4{p2.y - p0.y, p2.z - p0.z, 0.f, 0.f};
    float m = fmaxf(fmaxf(fabsf(p0.x), fabsf(p0.y)), fabsf(p0.z));
    m = fmaxf(m, fmaxf(fmaxf(fabsf(p1.x), fabsf(p1.y)), fabsf(p1.z)));
    return fmaxf(m, fmaxf(fmaxf(fabsf(p2.x), fabsf(p2.y)), fabsf(p2.z)));
}
// Recomputes the boxes of the leaf children of node `idx` (inner-child slots must already hold their subtree boxes),
// stores them, and returns the union of all child boxes in lo/hi and its half-area.
EPSM_HD float refit_node(f4 *nodes, uint32_t idx, const f4 *tris, const uint32_t *faces, const float *pos, float pad, float *lo, float *hi) {
    volatile float *n = &nodes[(size_t) idx * EPSM_NODE_F4].x;                // slots written by other CTAs: no L1
    for (int a = 0; a < 3; ++a) { lo[a] = 1e30f; hi[a] = -1e30f; }
    for (int s = 0; s < 4; ++s) {
        const int32_t ref = (int32_t) f2u(n[24 + s]);
        if (ref == 0x7fffffff) continue;
        float clo[3], chi[3];
        if (ref < 0) {
            const uint32_t code = ~(uint32_t) ref, first = code >> EPSM_LEAF_BITS, cnt = (code & (EPSM_MAX_LEAF - 1u)) + 1u;
            for (int a = 0; a < 3; ++a) { clo[a] = 3.402823466e38f; chi[a] = -3.402823466e38f; }
            for (uint32_t i = 0; i < cnt; ++i) {
                const uint32_t prim = f2u(tris[(size_t) (first + i) * EPSM_TRI_F4].w);
                const uint32_t *f = faces + 3 * (size_t) prim;
                for (int k = 0; k < 3; ++k) {
                    const float *p = pos + 3 * (size_t) f[k];
                    for (int a = 0; a < 3; ++a) { clo[a] = fminf(clo[a], p[a]); chi[a] = fmaxf(chi[a], p[a]); }
                }
            }
            for (int a = 0; a < 3; ++a) { clo[a] = clo[a] - pad; chi[a] = chi[a] + pad; n[4 * a + s] = clo[a]; n[12 + 4 * a + s] = chi[a]; }
        } else {
            for (int a = 0; a < 3; ++a) { clo[a] = n[4 * a + s]; chi[a] = n[12 + 4 * a + s]; }
        }
        for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], clo[a]); hi[a] = fmaxf(hi[a], chi[a]); }
    }
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return (dx >= 0.f && dy >= 0.f && dz >= 0.f) ? dx * dy + dy * dz + dz * dx : 0.f;
}
// Stores the subtree box of a finished node into its slot of the parent node.
EPSM_HD void refit_store_parent(f4 *nodes, uint32_t parent_link, const float *lo, const float *hi) {
    volatile float *pn = &nodes[(size_t) (parent_link >> 2) * EPSM_NODE_F4].x; const uint32_t s = parent_link & 3u;
    for (int a = 0; a < 3; ++a) { pn[4 * a + s] = lo[a]; pn[12 + 4 * a + s] = hi[a]; }
}
EPSM_HD float refit_pad(float max_abs) { return 1e-5f * max_abs + 1e-7f; }    // epsm_bvh.cpp: Möller–Trumbore slack

// Closest hit (ANY = false) or occlusion test (ANY = true) through the BVH4.
// Ties in t are resolved towards the larger primitive index, which is what scanning primitives in
// order with `t <= maxt` (mesh.h:363; kdtree.h:2188-2290) produces; box culling is inclusive and the
// boxes are padded, so every triangle the exact test would accept is reached.
// COUNT adds the algorithmic-byte counters of SURVEY.md §8d (nodes fetched, triangles tested).
template <bool ANY, bool COUNT>
EPSM_HDN Hit bvh_trace(const DScene &sc, const Ray &ray, TStack &st, uint32_t &n_nodes, uint32_t &n_tris) {
    Hit hit; hit.t = ray.maxt; hit.u = hit.v = 0.f; hit.prim = 0xffffffffu;
    if (sc.n_nodes == 0) return hit;
    float dx = fabsf(ray.d.x) < 1e-20f ? mulsign(1e-20f, ray.d.x) : ray.d.x;
    float dy = fabsf(ray.d.y) < 1e-20f ? mulsign(1e-20f, ray.d.y) : ray.d.y;
    float dz = fabsf(ray.d.z) < 1e-20f ? mulsign(1e-20f, ray.d.z) : ray.d.z;
    const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;
    const float ox = ray.o.x * ix, oy = ray.o.y * iy, oz = ray.o.z * iz;
    bool found = false;
    st.sp = 0;
    st.push(0u, 0.0f);
    while (st.sp > 0) {
        uint32_t ref; float tn;
        st.pop(ref, tn);
        if (tn > hit.t) continue;
        if ((int32_t) ref < 0) {
            // ---- leaf: up to 4 triangles, one 256-bit + one 64-bit load each ----
            uint32_t code = ~ref, first = code >> EPSM_LEAF_BITS, cnt = (code & (EPSM_MAX_LEAF - 1u)) + 1u;
            for (uint32_t i = 0; i < cnt; ++i) {
                const f4 *tp = sc.tris + (size_t) (first + i) * EPSM_TRI_F4;
                const f8 ab = ldg8(tp); const f2 c = ldg2(tp + 2); const f4 a = ab.a, b = ab.b;
                if (COUNT) ++n_tris;
                float t, u, v;
                if (tri_test(ray, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(b.w, c.x, c.y), hit.t, t, u, v)) {
                    if (ANY) { hit.t = t; hit.prim = f2u(a.w); return hit; }
                    uint32_t prim = f2u(a.w);
                    if (!found || t < hit.t || prim > hit.prim) { hit.t = t; hit.u = u; hit.v = v; hit.prim = prim; found = true; }
                }
            }
            continue;
        }
        // ---- inner node: one 128-byte line as three 256-bit loads + one 128-bit load ----
        const f4 *np = sc.nodes + (size_t) ref * EPSM_NODE_F4;
        const f8 n0 = ldg8(np), n1 = ldg8(np + 2), n2 = ldg8(np + 4);
        const f4 lx = n0.a, ly = n0.b, lz = n1.a, hx = n1.b, hy = n2.a, hz = n2.b, ch = ldg4(np + 6);
        if (COUNT) ++n_nodes;
        uint32_t key[4];
        const uint32_t cref[4] = {f2u(ch.x), f2u(ch.y), f2u(ch.z), f2u(ch.w)};     // 0x7fffffff marks an empty slot
#define EPSM_SLAB(k, LX, LY, LZ, HX, HY, HZ)                                                        \
        {                                                                                           \
            float ax = fmaf(LX, ix, -ox), bx = fmaf(HX, ix, -ox);                                   \
            float ay = fmaf(LY, iy, -oy), by = fmaf(HY, iy, -oy);                                   \
            float az = fmaf(LZ, iz, -oz), bz = fmaf(HZ, iz, -oz);                                   \
            float t0 = fmaxf(fmax3(fminf(ax, bx), fminf(ay, by), fminf(az, bz)), 0.0f);             \
            float t1 = fminf(fmin3(fmaxf(ax, bx), fmaxf(ay, by), fmaxf(az, bz)), hit.t);            \
            key[k] = (t0 <= t1 * 1.0000004f && cref[k] != 0x7fffffffu) ? ((f2u(t0) & ~3u) | (uint32_t) k) : 0xffffffffu; \
        }
        EPSM_SLAB(0, lx.x, ly.x, lz.x, hx.x, hy.x, hz.x)
        EPSM_SLAB(1, lx.y, ly.y, lz.y, hx.y, hy.y, hz.y)
        EPSM_SLAB(2, lx.z, ly.z, lz.z, hx.z, hy.z, hz.z)
        EPSM_SLAB(3, lx.w, ly.w, lz.w, hx.w, hy.w, hz.w)
#undef EPSM_SLAB
        // sort ascending by entry distance (5-comparator network), push far → near
        cswap(key[0], key[1]); cswap(key[2], key[3]); cswap(key[0], key[2]); cswap(key[1], key[3]); cswap(key[1], key[2]);
#pragma unroll
        for (int i = 3; i >= 0; --i) {
            if (key[i] != 0xffffffffu) {
                uint32_t k = key[i] & 3u;
                uint32_t r = k == 0 ? cref[0] : k == 1 ? cref[1] : k == 2 ? cref[2] : cref[3];
                // the key carries the entry distance (two mantissa bits cleared: never larger, so the pop test stays conservative)
                float t = u2f(key[i] & ~3u);
                st.push(r, t);
            }
        }
    }
    if (!found) hit.t = INFINITY;
    return hit;
}

// ------------------------------------------------------------------------------------------------
//  Surface interaction
// ------------------------------------------------------------------------------------------------
struct SI {
    bool valid; float t; f3 p, n, sh_n, sh_s, sh_t, wi;
    uint32_t prim, mesh, fi[3]; float bw[3];
    uint32_t mesh_flags, has_normals; int32_t bsdf, emitter;
    // EPSM fields (interaction.h:221-224)
    float b0, b1; f3 p0, p1, p2, n0, n1, n2; uint32_t ismesh;
    float uv[2];                 // mesh.cpp:731-744
};
EPSM_HD f3 to_local(const SI &si, f3 v) { return mk3(dot3(v, si.sh_s), dot3(v, si.sh_t), dot3(v, si.sh_n)); }
EPSM_HD f3 to_world(const SI &si, f3 v) { return fma3(si.sh_n, v.z, fma3(si.sh_t, v.y, si.sh_s * v.x)); }

EPSM_HDN void compute_si(const DScene &sc, const Ray &ray, const Hit &h, SI &si) {
    const f3 z = mk3(0, 0, 0);
    si.valid = h.prim != 0xffffffffu; si.t = INFINITY; si.prim = 0xffffffffu; si.mesh = 0;
    si.p = si.n = si.sh_n = si.sh_s = si.sh_t = z; si.wi = -ray.d;
    si.fi[0] = si.fi[1] = si.fi[2] = 0; si.bw[0] = si.bw[1] = si.bw[2] = 0.f;
    si.mesh_flags = 0; si.has_normals = 0; si.bsdf = -1; si.emitter = -1;
    si.b0 = si.b1 = 0.f; si.p0 = si.p1 = si.p2 = si.n0 = si.n1 = si.n2 = z; si.ismesh = 0; si.uv[0] = si.uv[1] = 0.f;
    if (!si.valid) return;
    const uint32_t mid = sc.face_mesh[h.prim]; const DMesh m = sc.meshes[mid];
    const uint32_t f0 = sc.faces[3 * h.prim], f1 = sc.faces[3 * h.prim + 1], f2 = sc.faces[3 * h.prim + 2];
    const f3 p0 = ld3(sc.pos, f0), p1 = ld3(sc.pos, f1), p2 = ld3(sc.pos, f2);
    const float b1 = h.u, b2 = h.v, b0 = 1.f - b1 - b2;
    const f3 dp0 = p1 - p0, dp1 = p2 - p0;
    si.t = h.t; si.prim = h.prim; si.mesh = mid; si.mesh_flags = m.flags; si.has_normals = m.has_normals;
    si.bsdf = m.bsdf; si.emitter = m.emitter;
    si.fi[0] = f0; si.fi[1] = f1; si.fi[2] = f2; si.bw[0] = b0; si.bw[1] = b1; si.bw[2] = b2;
    si.p = fma3(p0, b0, fma3(p1, b1, p2 * b2));                                    // mesh.cpp:707
    si.n = norm3(cross3(dp0, dp1));                                                // mesh.cpp:729
    f3 dp_du, dp_dv; coordinate_system(si.n, dp_du, dp_dv);                        // mesh.cpp:734
    si.uv[0] = b1; si.uv[1] = b2;                                                  // mesh.cpp:731
    if (m.has_uv) {                                                                // mesh.cpp:736-763 (RayFlags::All): interpolated uv, dp/du, dp/dv
        const float *t0 = sc.uv + 2 * (size_t) f0, *t1 = sc.uv + 2 * (size_t) f1, *t2 = sc.uv + 2 * (size_t) f2;
        const float u0 = t0[0], v0 = t0[1], u1 = t1[0], v1 = t1[1], u2 = t2[0], v2 = t2[1];
        si.uv[0] = fmaf(u2, b2, fmaf(u1, b1, u0 * b0)); si.uv[1] = fmaf(v2, b2, fmaf(v1, b1, v0 * b0));
        const float du0 = u1 - u0, dv0 = v1 - v0, du1 = u2 - u0, dv1 = v2 - v0;
        const float det = fmaf(du0, dv1, -(dv0 * du1));
        if (det != 0.f) {
            const float inv_det = rcp_e(det);
            dp_du = mk3(fmaf(dv1, dp0.x, -(dv0 * dp1.x)), fmaf(dv1, dp0.y, -(dv0 * dp1.y)), fmaf(dv1, dp0.z, -(dv0 * dp1.z))) * inv_det;
            dp_dv = mk3(fmaf(-du1, dp0.x, du0 * dp1.x), fmaf(-du1, dp0.y, du0 * dp1.y), fmaf(-du1, dp0.z, du0 * dp1.z)) * inv_det;
        }
    }
    f3 n0, n1, n2;
    if (m.has_normals) {                                                           // mesh.cpp:771-800
        n0 = ld3(sc.nrm, f0); n1 = ld3(sc.nrm, f1); n2 = ld3(sc.nrm, f2);
        f3 n = fma3(n2, b2, fma3(n1, b1, n0 * b0));
        si.sh_n = n * rsqrt_e(sqn3(n));
    } else {                                                                       // mesh.cpp:809-815
        si.sh_n = si.n; n0 = n1 = n2 = si.n;
    }
    if (m.flags & EPSM_MESH_FLIP_NORMALS) { si.n = -si.n; si.sh_n = -si.sh_n; n0 = -n0; n1 = -n1; n2 = -n2; }
    si.sh_s = norm3(fma3(si.sh_n, -dot3(si.sh_n, dp_du), dp_du));                  // interaction.h:262-275
    si.sh_t = cross3(si.sh_n, si.sh_s);
    si.wi = to_local(si, -ray.d);                                                  // interaction.h:499
    if (m.flags & EPSM_MESH_ISMESH) {                                              // mesh.cpp:712-720
        si.b0 = b0; si.b1 = b1; si.p0 = p0; si.p1 = p1; si.p2 = p2; si.n0 = n0; si.n1 = n1; si.n2 = n2; si.ismesh = 1;
    }
}

// Bitmap reflectance at si.uv: Dr.Jit Texture<Float, 2> without hardware acceleration (ext/drjit/include/drjit/texture.h: wrap
// :1143-1168, nearest :346-360, bilinear :362-405) under src/textures/bitmap.cpp:582-625.
EPSM_HD int tex_wrap(int p, int n, uint32_t mode) {
    if (mode == EPSM_TEX_CLAMP) return p < 0 ? 0 : (p > n - 1 ? n - 1 : p);
    int div = p / n; if (p < 0 && div * n != p) --div;                             // floor division
    int mod = p - div * n;
    if (mode == EPSM_TEX_MIRROR && (div & 1)) mod = n - 1 - mod;
    return mod;
}
EPSM_HD f3 tex_fetch(const DScene &sc, const DTexture &t, int x, int y) {
    const float *c = sc.texels + t.off + 3 * ((size_t) tex_wrap(y, (int) t.h, t.wrap) * t.w + (size_t) tex_wrap(x, (int) t.w, t.wrap));
    return mk3(c[0], c[1], c[2]);
}
EPSM_HD f3 tex_eval(const DScene &sc, const DTexture &t, const float *uv_in) {
    const float u = fmaf(t.to_uv[0], uv_in[0], fmaf(t.to_uv[1], uv_in[1], t.to_uv[2])), v = fmaf(t.to_uv[3], uv_in[0], fmaf(t.to_uv[4], uv_in[1], t.to_uv[5]));
    if (t.filter == EPSM_TEX_NEAREST) return tex_fetch(sc, t, (int) floorf(u * (float) t.w), (int) floorf(v * (float) t.h));
    const float fx = fmaf(u, (float) t.w, -.5f), fy = fmaf(v, (float) t.h, -.5f);
    const int ix = (int) floorf(fx), iy = (int) floorf(fy);
    const float w1x = fx - (float) ix, w1y = fy - (float) iy, w0x = 1.f - w1x, w0y = 1.f - w1y;
    const f3 v00 = tex_fetch(sc, t, ix, iy), v10 = tex_fetch(sc, t, ix + 1, iy), v01 = tex_fetch(sc, t, ix, iy + 1), v11 = tex_fetch(sc, t, ix + 1, iy + 1);
    const f3 r0 = fma3(v00, w0x, v10 * w1x), r1 = fma3(v01, w0x, v11 * w1x);
    return fma3(r0, w0y, r1 * w1y);
}
// The BSDF of a hit with its textured parameter evaluated at the hit's uv (diffuse.cpp:107-140: m_reflectance->eval(si)).
EPSM_HD DBsdf bsdf_at(const DScene &sc, int32_t bsdf, const float *uv) {
    DBsdf b = sc.bsdfs[bsdf];
    if (b.tex >= 0) { const f3 c = tex_eval(sc, sc.textures[b.tex], uv); b.reflectance[0] = c.x; b.reflectance[1] = c.y; b.reflectance[2] = c.z; }
    return b;
}

EPSM_HD f3 offset_p(const SI &si, f3 d) {                                          // interaction.h:161-165
    float mag = (1.f + max3(mk3(fabsf(si.p.x), fabsf(si.p.y), fabsf(si.p.z)))) * EPSM_RAY_EPS;
    mag = mulsign(mag, dot3(si.n, d));
    return fma3(si.n, mag, si.p);
}
EPSM_HD Ray spawn_ray(const SI &si, f3 d) { return Ray{offset_p(si, d), d, FLT_MAX}; }
EPSM_HD Ray spawn_ray_to(const SI &si, f3 t) {                                     // interaction.h:141-147
    f3 o = offset_p(si, t - si.p); f3 d = t - o; float dist = len3(d);
    d = d * (1.0f / dist);
    return Ray{o, d, dist * (1.f - EPSM_SHADOW_EPS)};
}

// ------------------------------------------------------------------------------------------------
//  BSDFs
// ------------------------------------------------------------------------------------------------
struct BSample { f3 wo, hf, weight; float pdf, eta; uint32_t type; };

EPSM_HD void fresnel_diel(float cos_theta_i, float eta, float &r, float &cos_theta_t, float &eta_it, float &eta_ti) {   // fresnel.h:35-72
    bool outside = cos_theta_i >= 0.f;
    float rcp_eta = rcp_e(eta); eta_it = outside ? eta : rcp_eta; eta_ti = outside ? rcp_eta : eta;
    float cos_theta_t_sqr = fmaf(-fmaf(-cos_theta_i, cos_theta_i, 1.f), eta_ti * eta_ti, 1.f);
    float ci = fabsf(cos_theta_i), ct = safe_sqrt(cos_theta_t_sqr);
    bool index_matched = eta == 1.f, special = index_matched || ci == 0.f;
    float a_s = fmaf(-eta_it, ct, ci) / fmaf(eta_it, ct, ci);
    float a_p = fmaf(-eta_it, ci, ct) / fmaf(eta_it, ci, ct);
    r = 0.5f * (sqr(a_s) + sqr(a_p));
    if (special) r = index_matched ? 0.f : 1.f;
    cos_theta_t = mulsign_neg(ct, cos_theta_i);
}
EPSM_HD float fresnel_cond(float cos_theta_i, float eta_r, float eta_i) {          // fresnel.h:93-117
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
EPSM_HD f3 fresnel_cond3(float c, const DBsdf &b) {
    return mk3(fresnel_cond(c, b.eta[0], b.k[0]), fresnel_cond(c, b.eta[1], b.k[1]), fresnel_cond(c, b.eta[2], b.k[2]));
}
// microfacet.h:183-210 (eval), :365-392 (smith_g1), :244-302 (sample — the EPSM fork's `if (true)` branch)
EPSM_HD float mf_eval(const DBsdf &b, f3 m) {
    bool ggx = (b.flags & EPSM_BSDF_GGX) != 0;
    float au = b.alpha_u, av = b.alpha_v, alpha_uv = au * av, ct = m.z, ct2 = sqr(ct), result;
    if (!ggx) result = (float) exp((double) (-(sqr(m.x / au) + sqr(m.y / av)) / ct2)) / (EPSM_PI * alpha_uv * sqr(ct2));
    else result = rcp_e(EPSM_PI * alpha_uv * sqr(sqr(m.x / au) + sqr(m.y / av) + sqr(m.z)));
    return result * ct > 1e-20f ? result : 0.f;
}
EPSM_HD float mf_g1(const DBsdf &b, f3 v, f3 m) {
    bool ggx = (b.flags & EPSM_BSDF_GGX) != 0;
    float xy_alpha_2 = sqr(b.alpha_u * v.x) + sqr(b.alpha_v * v.y), tan_theta_alpha_2 = xy_alpha_2 / sqr(v.z), result;
    if (!ggx) {
        float a = rsqrt_e(tan_theta_alpha_2), a_sqr = sqr(a);
        result = a >= 1.6f ? 1.f : (3.535f * a + 2.181f * a_sqr) / (1.f + 2.276f * a + 2.577f * a_sqr);
    } else {
        result = 2.f / (1.f + sqrtf(1.f + tan_theta_alpha_2));
    }
    if (xy_alpha_2 == 0.f) result = 1.f;
    if (dot3(v, m) * v.z <= 0.f) result = 0.f;
    return result;
}
EPSM_HD void mf_sample(const DBsdf &b, float sx, float sy, f3 &m, float &pdf) {
    bool ggx = (b.flags & EPSM_BSDF_GGX) != 0;
    float au = b.alpha_u, av = b.alpha_v, sin_phi, cos_phi, cos_theta, cos_theta_2, alpha_2;
    if (au == av) {
        sincos_dr((2.f * EPSM_PI) * sy, sin_phi, cos_phi);
        alpha_2 = au * au;
    } else {
        float ratio = av / au, tmp = ratio * (float) tan((double) ((2.f * EPSM_PI) * sy));
        cos_phi = rsqrt_e(fmaf(tmp, tmp, 1.f));
        cos_phi = mulsign(cos_phi, fabsf(sy - .5f) - .25f);
        sin_phi = cos_phi * tmp;
        alpha_2 = rcp_e(sqr(cos_phi / au) + sqr(sin_phi / av));
    }
    if (!ggx) {
        cos_theta = rsqrt_e(fmaf(-alpha_2, (float) log((double) (1.f - sx)), 1.f));
        cos_theta_2 = sqr(cos_theta);
        float cos_theta_3 = fmaxf(cos_theta_2 * cos_theta, 1e-20f);
        pdf = (1.f - sx) / (EPSM_PI * au * av * cos_theta_3);
    } else {
        float tan_theta_m_2 = alpha_2 * sx / (1.f - sx);
        cos_theta = rsqrt_e(1.f + tan_theta_m_2);
        cos_theta_2 = sqr(cos_theta);
        float temp = 1.f + tan_theta_m_2 / alpha_2, cos_theta_3 = fmaxf(cos_theta_2 * cos_theta, 1e-20f);
        pdf = rcp_e(EPSM_PI * au * av * cos_theta_3 * sqr(temp));
    }
    float sin_theta = sqrtf(1.f - cos_theta_2);
    m = mk3(cos_phi * sin_theta, sin_phi * sin_theta, cos_theta);
}
EPSM_HD f3 reflect_m(f3 wi, f3 m) { float s = 2.f * dot3(wi, m); return mk3(fmaf(m.x, s, -wi.x), fmaf(m.y, s, -wi.y), fmaf(m.z, s, -wi.z)); }

EPSM_HDN void bsdf_sample(const DBsdf &b, f3 wi_in, float s1, float s2x, float s2y, bool active, BSample &bs) {
    const f3 z = mk3(0, 0, 0);
    bs.wo = bs.hf = bs.weight = z; bs.pdf = 0.f; bs.eta = 0.f; bs.type = 0;
    const bool two = (b.flags & EPSM_BSDF_TWOSIDED) != 0;                          // twosided.cpp:124-127
    f3 wi = wi_in; if (two) wi.z = fabsf(wi.z);
    const f3 refl = a3(b.reflectance);
    switch (b.kind) {
    case EPSM_BSDF_DIFFUSE: {                                                      // diffuse.cpp:101-125
        if (active && wi.z > 0.f) {
            bs.wo = square_to_cosine_hemisphere(s2x, s2y);
            bs.pdf = EPSM_INV_PI * bs.wo.z; bs.eta = 1.f; bs.type = EPSM_BF_DIFFUSE_REFLECTION;
            if (bs.pdf > 0.f) bs.weight = refl;
        }
    } break;
    case EPSM_BSDF_DIELECTRIC: {                                                   // dielectric.cpp:250-366
        float r_i, cos_theta_t, eta_it, eta_ti;
        fresnel_diel(wi.z, b.eta[0], r_i, cos_theta_t, eta_it, eta_ti);
        float t_i = 1.f - r_i;
        bool sel_r = s1 <= r_i && active, sel_t = !sel_r && active;
        bs.pdf = sel_r ? r_i : t_i;
        bs.type = sel_r ? EPSM_BF_DELTA_REFLECTION : EPSM_BF_DELTA_TRANSMISSION;
        bs.wo = sel_r ? mk3(-wi.x, -wi.y, wi.z) : mk3(-eta_ti * wi.x, -eta_ti * wi.y, cos_theta_t);
        bs.eta = sel_r ? 1.f : eta_it;
        f3 w = mk3(1, 1, 1);
        if (sel_r) w = w * refl;
        if (sel_t) { w = w * a3(b.transmittance); w = w * sqr(eta_ti); }
        if (active) bs.weight = w;
    } break;
    case EPSM_BSDF_CONDUCTOR: {                                                    // conductor.cpp:248-300
        if (active && wi.z > 0.f) {
            bs.type = EPSM_BF_DELTA_REFLECTION; bs.wo = mk3(-wi.x, -wi.y, wi.z); bs.eta = 1.f; bs.pdf = 1.f;
            bs.weight = refl * fresnel_cond3(wi.z, b);
        }
    } break;
    case EPSM_BSDF_ROUGHCONDUCTOR: {                                               // roughconductor.cpp:226-305
        bool act = active && wi.z > 0.f;
        if (act) {
            f3 m; mf_sample(b, s2x, s2y, m, bs.pdf);
            bs.wo = reflect_m(wi, m); bs.eta = 1.f; bs.type = EPSM_BF_GLOSSY_REFLECTION; bs.hf = m;
            act &= bs.pdf != 0.f && bs.wo.z > 0.f;
            float weight = (b.flags & EPSM_BSDF_SAMPLE_VISIBLE) ? mf_g1(b, bs.wo, m)
                                                                : mf_g1(b, wi, m) * mf_g1(b, bs.wo, m) * dot3(wi, m) / (wi.z * m.z);
            bs.pdf /= 4.f * dot3(bs.wo, m);
            f3 F = fresnel_cond3(dot3(wi, m), b);
            if (act) bs.weight = F * weight * refl;
        }
    } break;
    }
    if (two) bs.wo.z = mulsign(bs.wo.z, wi_in.z);
}

EPSM_HDN void bsdf_eval_pdf(const DBsdf &b, f3 wi, f3 wo, f3 &value, float &pdf) {
    value = mk3(0, 0, 0); pdf = 0.f;
    if (b.flags & EPSM_BSDF_TWOSIDED) { wo.z = mulsign(wo.z, wi.z); wi.z = fabsf(wi.z); }   // twosided.cpp:226-232
    const f3 refl = a3(b.reflectance);
    if (b.kind == EPSM_BSDF_DIFFUSE) {                                             // diffuse.cpp:158-175
        if (wi.z > 0.f && wo.z > 0.f) { value = refl * EPSM_INV_PI * wo.z; pdf = EPSM_INV_PI * wo.z; }
    } else if (b.kind == EPSM_BSDF_ROUGHCONDUCTOR) {                               // roughconductor.cpp:417-492
        f3 H = norm3(wo + wi);
        bool active = wi.z > 0.f && wo.z > 0.f && dot3(wi, H) > 0.f && dot3(wo, H) > 0.f;
        if (!active) return;
        float D = mf_eval(b, H);
        active &= D != 0.f;
        float g1 = mf_g1(b, wi, H), G = g1 * mf_g1(b, wo, H);
        float val = D * G / (4.f * wi.z);
        f3 F = fresnel_cond3(dot3(wi, H), b);
        float p = (b.flags & EPSM_BSDF_SAMPLE_VISIBLE) ? D * g1 / (4.f * wi.z) : D * H.z / (4.f * dot3(wo, H));
        if (active) { value = F * (refl * val); pdf = p; }
    }
}

// ------------------------------------------------------------------------------------------------
//  Emitters
// ------------------------------------------------------------------------------------------------
struct DSamp { f3 p, n, d; float pdf, dist; int32_t emitter; };

EPSM_HD f3 emitter_eval(const DScene &sc, const SI &si) {                           // area.cpp:82-89
    if (!si.valid || si.emitter < 0 || !(si.wi.z > 0.f)) return mk3(0, 0, 0);
    return a3(sc.emitters[si.emitter].radiance);
}
EPSM_HD float mis_weight(float a, float b) { float a2 = a * a; return a > 0.f ? a2 / fmaf(b, b, a2) : 0.f; }   // common.py:1224-1230

// Scene::pdf_emitter_direction for the DirectionSample built from a surface hit (records.h:173-180)
EPSM_HD float pdf_emitter_direction(const DScene &sc, const SI &si, f3 ref_p, bool active) {
    if (!si.valid || !active || si.emitter < 0) return 0.f;
    f3 rel = si.p - ref_p; float dist = len3(rel); f3 d = rel * (1.0f / dist);
    float dp = dot3(d, si.sh_n);
    if (!(dp < 0.f)) return 0.f;                                                   // area.cpp:140-142
    float area_norm = sc.meshes[sc.emitters[si.emitter].mesh].area_norm;
    float adp = fabsf(dp);
    float pdf = area_norm * (adp != 0.f ? (dist * dist) / adp : 0.f);              // shape.cpp:372-383
    return pdf * (1.0f / (float) sc.n_emitters);
}

// Scene::sample_emitter_direction(ref, sample, test_visibility=true, active); `occl` does the shadow query.
template <typename OcclFn>
EPSM_HDN void sample_emitter_direction(const DScene &sc, const SI &ref, float sx, float sy, bool active, DSamp &ds, f3 &spec, OcclFn occl) {
    ds.p = ds.n = ds.d = mk3(0, 0, 0); ds.pdf = 0.f; ds.dist = 0.f; ds.emitter = -1; spec = mk3(0, 0, 0);
    if (sc.n_emitters == 0 || !active) return;
    uint32_t index = 0; float emitter_weight = 1.f, pmf = 1.f;
    if (sc.n_emitters > 1) {                                                       // scene.cpp:161-178
        float ef = (float) sc.n_emitters, scaled = sx * ef;
        uint32_t idx = (uint32_t) scaled; index = idx < sc.n_emitters - 1u ? idx : sc.n_emitters - 1u;
        sx = scaled - (float) index; emitter_weight = ef; pmf = 1.f / ef;
    }
    const DEmitter em = sc.emitters[index]; const DMesh m = sc.meshes[em.mesh];
    f3 n;
    if (m.flags & EPSM_MESH_RECTANGLE) {                                           // rectangle.cpp:152-166
        f3 c0 = ld3(sc.pos, m.v_off), ex = ld3(sc.pos, m.v_off + 1) - c0, ey = ld3(sc.pos, m.v_off + 3) - c0;
        ds.p = fma3(ex, sx, fma3(ey, sy, c0));
        n = norm3(cross3(ex, ey));
    } else {                                                                       // mesh.cpp:516-571
        const float *cdf = sc.cdf + m.cdf_off, *pm = sc.pmf + m.cdf_off;
        float value = sy * m.area_sum;
        uint32_t lo = m.valid_lo, hi = m.valid_hi;                                 // dr::binary_search (distr_1d.h:117-128)
        while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (cdf[mid] < value) lo = (mid + 1 < hi ? mid + 1 : hi); else hi = mid; }
        uint32_t face = lo;
        float pmv = pm[face] * m.area_norm, cdv = face > 0 ? cdf[face - 1] * m.area_norm : 0.f;   // distr_1d.h:167-176
        sy = (sy - cdv) / pmv;
        const uint32_t g = m.f_off + face;
        const uint32_t f0 = sc.faces[3 * g], f1 = sc.faces[3 * g + 1], f2 = sc.faces[3 * g + 2];
        f3 p0 = ld3(sc.pos, f0), e0 = ld3(sc.pos, f1) - p0, e1 = ld3(sc.pos, f2) - p0;
        float t = safe_sqrt(1.f - sx), bx = 1.f - t, by = t * sy;                  // warp.h:153-156
        ds.p = fma3(e0, bx, fma3(e1, by, p0));
        if (m.has_normals) n = fma3(ld3(sc.nrm, f0), (1.f - bx - by), fma3(ld3(sc.nrm, f1), bx, ld3(sc.nrm, f2) * by));
        else n = cross3(e0, e1);
        n = norm3(n);
    }
    if (m.flags & EPSM_MESH_FLIP_NORMALS) n = -n;
    ds.n = n; ds.pdf = m.area_norm; ds.emitter = (int32_t) index;
    ds.d = ds.p - ref.p;                                                           // shape.cpp:352-370
    float dist2 = sqn3(ds.d); ds.dist = sqrtf(dist2);
    ds.d = ds.d * (1.0f / ds.dist);
    float dp = fabsf(dot3(ds.d, ds.n)), x = dist2 / dp;
    ds.pdf *= (fabsf(x) <= FLT_MAX) ? x : 0.f;                                     // isfinite
    bool act = dot3(ds.d, ds.n) < 0.f && ds.pdf != 0.f;                            // area.cpp:127
    f3 s = act ? a3(em.radiance) * (1.0f / ds.pdf) : mk3(0, 0, 0);
    ds.pdf *= pmf; s = s * emitter_weight;
    if (ds.pdf != 0.f) {                                                           // scene.cpp:253-263
        if (occl(spawn_ray_to(ref, ds.p))) { s = mk3(0, 0, 0); ds.pdf = 0.f; }
    }
    spec = s;
}

// ------------------------------------------------------------------------------------------------
//  Camera rays
// ------------------------------------------------------------------------------------------------
struct CamRay { Ray ray; f3 d_x, d_y; float pos_x, pos_y; uint32_t px, py; };
EPSM_HD f3 xform_point(const float *m, f3 p) {
    float x = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
    float y = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    float z = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
    float w = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    float iw = 1.0f / w; return mk3(x * iw, y * iw, z * iw);
}
EPSM_HD f3 xform_vec(const float *m, f3 v) {
    return mk3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z);
}
// Lane → pixel → jittered film position → ray + differentials.  The lane's first two random numbers are the jitter.
EPSM_HDN void camera_ray(const DSensor &s, uint32_t px, uint32_t py, Pcg &rng, CamRay &c) {
    float jx = pcg_f32(rng), jy = pcg_f32(rng);
    float pos_x = (float) ((int) px - s.border) + jx, pos_y = (float) ((int) py - s.border) + jy;     // pos -= border_size (common.py:337-338)
    float sx = 1.0f / (float) s.fw, sy = 1.0f / (float) s.fh;
    float ax = fmaf(pos_x, sx, -0.f * sx), ay = fmaf(pos_y, sy, -0.f * sy);
    f3 near_p = xform_point(s.s2c, mk3(ax, ay, 0.f));
    f3 d = norm3(near_p);
    f3 o = mk3(s.to_world[3], s.to_world[7], s.to_world[11]);
    f3 dw = xform_vec(s.to_world, d);
    float inv_z = rcp_e(d.z), near_t = s.near_clip * inv_z, far_t = s.far_clip * inv_z;
    c.ray.o = o + dw * near_t; c.ray.d = dw; c.ray.maxt = far_t - near_t;
    c.d_x = xform_vec(s.to_world, norm3(near_p + s.dx));
    c.d_y = xform_vec(s.to_world, norm3(near_p + s.dy));
    c.pos_x = pos_x; c.pos_y = pos_y; c.px = px; c.py = py;
}

// ------------------------------------------------------------------------------------------------
//  Traversal counters
// ------------------------------------------------------------------------------------------------
struct Tracer {
    const DScene &sc; TStack &st; uint32_t n_nodes, n_tris, n_rays;
    EPSM_HD Tracer(const DScene &s, TStack &t) : sc(s), st(t), n_nodes(0), n_tris(0), n_rays(0) {}
};
template <bool COUNT> EPSM_HD Hit trace_closest(Tracer &tr, const Ray &r) {
    if (COUNT) ++tr.n_rays;
    return bvh_trace<false, COUNT>(tr.sc, r, tr.st, tr.n_nodes, tr.n_tris);
}
template <bool COUNT> EPSM_HD bool trace_any(Tracer &tr, const Ray &r) {
    if (COUNT) ++tr.n_rays;
    return bvh_trace<true, COUNT>(tr.sc, r, tr.st, tr.n_nodes, tr.n_tris).prim != 0xffffffffu;
}

// ------------------------------------------------------------------------------------------------
//  EPSMIntegrator.sample — primal path (epsm.py:308-501): 6 random numbers per bounce
// ------------------------------------------------------------------------------------------------
template <bool COUNT>
EPSM_HDN f3 path_primal(const DScene &sc, Tracer &tr, Ray ray, Pcg &rng, uint32_t max_depth, uint32_t rr_depth, bool &valid_out) {
    uint32_t depth = 0; f3 L = mk3(0, 0, 0), beta = mk3(1, 1, 1); float eta = 1.f; bool active = true;
    f3 prev_p = mk3(0, 0, 0); float prev_pdf = 1.f; bool prev_delta = true;
    SI si;
    for (uint32_t it = 0; it < max_depth && active; ++it) {
        Hit h = trace_closest<COUNT>(tr, ray);
        compute_si(sc, ray, h, si);
        float em_pdf = pdf_emitter_direction(sc, si, prev_p, !prev_delta);
        float mis = mis_weight(prev_pdf, em_pdf);
        f3 Le = beta * mis * emitter_eval(sc, si);
        bool active_next = (depth + 1 < max_depth) && si.valid;
        DBsdf b; uint32_t flags = 0;
        if (si.valid) { b = bsdf_at(sc, si.bsdf, si.uv); flags = b.bflags; }
        bool active_em = active_next && (flags & EPSM_BF_SMOOTH);
        float e1 = pcg_f32(rng), e2 = pcg_f32(rng);
        DSamp ds; f3 em_weight;
        sample_emitter_direction(sc, si, e1, e2, active_em, ds, em_weight, [&](const Ray &r) { return trace_any<COUNT>(tr, r); });
        active_em &= ds.pdf != 0.f;
        f3 Lr_dir = mk3(0, 0, 0);
        if (active_em) {
            f3 wo = to_local(si, ds.d); f3 val; float bpdf;
            bsdf_eval_pdf(b, si.wi, wo, val, bpdf);
            Lr_dir = beta * mis_weight(ds.pdf, bpdf) * val * em_weight;
        }
        float s1 = pcg_f32(rng), s2x = pcg_f32(rng), s2y = pcg_f32(rng);
        BSample bs; bs.wo = bs.hf = bs.weight = mk3(0, 0, 0); bs.pdf = 0.f; bs.eta = 0.f; bs.type = 0;
        if (si.valid) bsdf_sample(b, si.wi, s1, s2x, s2y, active_next, bs);
        L = L + Le + Lr_dir;
        if (si.valid) ray = spawn_ray(si, to_world(si, bs.wo));
        eta *= bs.eta; beta = beta * bs.weight;
        prev_p = si.p; prev_pdf = bs.pdf; prev_delta = (bs.type & EPSM_BF_DELTA) != 0;
        float bmax = max3(beta);
        active_next &= bmax != 0.f;
        float rr_prob = fminf(bmax * eta * eta, .95f);
        bool rr_active = depth >= rr_depth;
        if (rr_active) beta = beta * rcp_e(rr_prob);
        bool rr_continue = pcg_f32(rng) < rr_prob;
        active_next &= !rr_active || rr_continue;
        if (si.valid) depth += 1;
        active = active_next;
    }
    valid_out = depth != 0;
    return L;
}

// ------------------------------------------------------------------------------------------------
//  The same primal path as a wavefront stage (k_pr_* in epsm_kernels.cu).  One stage per bounce: the emitter sample's
//  visibility is the only thing the rest of the bounce does not need (it gates Lr_dir alone, epsm.py:590-605 via
//  scene.cpp:259-263), so Lr_dir is parked in `pend` together with a shadow ray and added by the next stage (or by the
//  film kernel) once the any-hit query has answered.  Random-number order and arithmetic are those of path_primal.
// ------------------------------------------------------------------------------------------------
#define PR_ACTIVE 1u
#define PR_PREV_DELTA 2u
#define PR_PEND 4u
struct PrState {
    // the first PR_INIT_WORDS words are what ray generation hands over; pr_state_defaults() restores the rest
    Pcg rng;
    Ray ray;
    f3 beta, L, prev_p, pend;
    float eta, prev_pdf;
    uint32_t depth, bits;
};
#define PR_INIT_WORDS 11          // rng (4) + ray (7)
EPSM_HD void pr_state_defaults(PrState &S) {
    S.beta = mk3(1, 1, 1); S.L = S.prev_p = S.pend = mk3(0, 0, 0); S.eta = 1.f; S.prev_pdf = 1.f; S.depth = 0; S.bits = PR_ACTIVE | PR_PREV_DELTA;
}
EPSM_HD void pr_resolve_pending(PrState &S, bool occluded) {
    if (S.bits & PR_PEND) { if (!occluded) S.L = S.L + S.pend; S.bits &= ~PR_PEND; }
}
// One bounce after the closest hit `h` of S.ray (pending NEE of the previous bounce already resolved).  Returns true
// when `shadow` must be traced for S.pend; (S.bits & PR_ACTIVE) tells whether S.ray is the next path ray.
EPSM_HDN bool pr_stage(const DScene &sc, PrState &S, const Hit &h, uint32_t max_depth, uint32_t rr_depth, Ray &shadow) {
    SI si; compute_si(sc, S.ray, h, si);
    const bool prev_delta = (S.bits & PR_PREV_DELTA) != 0;
    float em_pdf = pdf_emitter_direction(sc, si, S.prev_p, !prev_delta);
    float mis = mis_weight(S.prev_pdf, em_pdf);
    f3 Le = S.beta * mis * emitter_eval(sc, si);
    bool active_next = (S.depth + 1 < max_depth) && si.valid;
    DBsdf b; uint32_t flags = 0;
    if (si.valid) { b = bsdf_at(sc, si.bsdf, si.uv); flags = b.bflags; }
    bool active_em = active_next && (flags & EPSM_BF_SMOOTH);
    float e1 = pcg_f32(S.rng), e2 = pcg_f32(S.rng);
    DSamp ds; f3 em_weight; bool need_shadow = false;
    sample_emitter_direction(sc, si, e1, e2, active_em, ds, em_weight, [&](const Ray &r) { shadow = r; need_shadow = true; return false; });
    active_em = active_em && ds.pdf != 0.f;
    f3 Lr_dir = mk3(0, 0, 0);
    if (active_em) {
        f3 wo = to_local(si, ds.d); f3 val; float bpdf;
        bsdf_eval_pdf(b, si.wi, wo, val, bpdf);
        Lr_dir = S.beta * mis_weight(ds.pdf, bpdf) * val * em_weight;
    }
    float s1 = pcg_f32(S.rng), s2x = pcg_f32(S.rng), s2y = pcg_f32(S.rng);
    BSample bs; bs.wo = bs.hf = bs.weight = mk3(0, 0, 0); bs.pdf = 0.f; bs.eta = 0.f; bs.type = 0;
    if (si.valid) bsdf_sample(b, si.wi, s1, s2x, s2y, active_next, bs);
    // L = L + Le + Lr_dir (epsm.py:466): Lr_dir joins when its visibility is known — same operands, same order
    S.L = S.L + Le;
    if (need_shadow) { S.pend = Lr_dir; S.bits |= PR_PEND; } else S.L = S.L + Lr_dir;
    if (si.valid) S.ray = spawn_ray(si, to_world(si, bs.wo));
    S.eta *= bs.eta; S.beta = S.beta * bs.weight;
    S.prev_p = si.p; S.prev_pdf = bs.pdf;
    S.bits = (bs.type & EPSM_BF_DELTA) ? (S.bits | PR_PREV_DELTA) : (S.bits & ~PR_PREV_DELTA);
    float bmax = max3(S.beta);
    active_next = active_next && bmax != 0.f;
    float rr_prob = fminf(bmax * S.eta * S.eta, .95f);
    bool rr_active = S.depth >= rr_depth;
    if (rr_active) S.beta = S.beta * rcp_e(rr_prob);
    bool rr_continue = pcg_f32(S.rng) < rr_prob;
    active_next = active_next && (!rr_active || rr_continue);
    if (si.valid) S.depth += 1;
    S.bits = active_next ? (S.bits | PR_ACTIVE) : (S.bits & ~PR_ACTIVE);
    return need_shadow;
}

// ------------------------------------------------------------------------------------------------
//  Gradient pass: sample_path replay (epsm.py:503-742) with the per-path log kept in thread-local
//  memory, then the manifold solve, then the scatter seeds.
// ------------------------------------------------------------------------------------------------
EPSM_HD float nan_to_num(float x) {                       // torch.nan_to_num defaults
    if (x != x) return 0.f;
    if (x > FLT_MAX) return FLT_MAX;
    if (x < -FLT_MAX) return -FLT_MAX;
    return x;
}
EPSM_HD f3 ntn3(f3 v) { return mk3(nan_to_num(v.x), nan_to_num(v.y), nan_to_num(v.z)); }
EPSM_HD float outlier(float x) { return (x > 0.1f || x < -0.1f) ? 0.f : x; }   // epsm.py:933-944
EPSM_HD f3 outlier3(f3 v) { return mk3(outlier(v.x), outlier(v.y), outlier(v.z)); }

struct VLog {
    bool active, active_em, direct_valid, direct_any, direct_grad, direct_traced;
    uint32_t flags, ismesh, mesh_flags, has_normals, prim;
    uint32_t fi[3];
    float bw[3];
    float eta, b0, b1, lr_sum;
    f3 light, p0, p1, p2, n0, n1, n2, p;
    // ---- the first VLOG_STAGE_WORDS words are what stage B knows and the wavefront keeps in HBM; the rest is filled by
    //      wf_direct_hit in the finish kernel (dfi, dbw, dis) or is never read (hf: its seed is identically zero)
    uint32_t dfi[3];
    float dbw[3];
    float dis;
    f3 hf;
};
#define VLOG_STAGE_WORDS 41       // 2 (flags) + 5 + 3 + 3 + 4 + 8 × 3
static_assert(offsetof(VLog, dfi) == 4 * VLOG_STAGE_WORDS, "VLog prefix");
struct PathHead { f3 cam, d, dldp1, L; float dlduv0, dlduv1; };
// NV = vertex slots held: EPSM_MAXV in general; the short-chain instantiation of the gradient pass (k_wf_finish) holds 3
// (two logged vertices + the inactive terminator), which lets the compiler keep the whole solve in registers.
template <int NV> struct PathLogT : PathHead {
    VLog v[NV];
};
typedef PathLogT<EPSM_MAXV> PathLog;

// Per-lane path state that persists between the wavefront stages (k_wf_* in epsm_kernels.cu).  The stages are
// cut exactly where sample_path queries the scene, so the trace kernels run between them on compacted queues:
//
//   wf_begin → [closest: path ray] → wf_stage_a → [any-hit: shadow ray] → wf_stage_b → [closest: direct ray ∥ next path ray] → …
//
// The random-number draw order is that of epsm.py:503-742 (2 emitter + 3 discarded + 3 kept + 1 RR per bounce).
struct WfState {
    // the first WF_INIT_WORDS words are all that wf_begin needs to hand over; wf_state_defaults() restores the rest
    Pcg rng;
    Ray ray;                      // current path ray
    f3 grad_d;
    f3 beta, L, prev_p;
    float eta, prev_pdf;
    uint32_t depth;
    bool active, prev_delta;
    bool chain;                   // every vertex so far continues the gradient chain of calc_grad (see wf_chain_vertex)
    // stage A → stage B
    bool active_next, smooth, need_shadow, chain_vertex;
    DSamp ds; f3 em_weight;
    Ray shadow_ray;
};

#define WF_INIT_WORDS 14          // rng (4) + ray (7) + grad_d (3)
EPSM_HD void wf_state_defaults(WfState &S) {
    S.beta = mk3(1, 1, 1); S.L = mk3(0, 0, 0); S.prev_p = mk3(0, 0, 0); S.eta = 1.f; S.prev_pdf = 1.f; S.depth = 0;
    S.active = true; S.prev_delta = true; S.active_next = S.smooth = S.need_shadow = false; S.chain = true; S.chain_vertex = false;
}

EPSM_HD uint32_t wf_iterations(uint32_t max_depth_prop) { return max_depth_prop < EPSM_MAX_ITER ? max_depth_prop : EPSM_MAX_ITER; }   // epsm.py:549

// Iterations of the replay: min(max_depth, 6) (epsm.py:549); only the first EPSM_MAXV vertices are logged (epsm.py:648), so a
// pruned replay never needs the sixth.
EPSM_HD uint32_t wf_replay_iterations(uint32_t max_depth_prop, bool prune) {
    const uint32_t m = wf_iterations(max_depth_prop);
    return prune && m > EPSM_MAXV ? EPSM_MAXV : m;
}

EPSM_HDN void wf_begin(const DSensor &sensor, uint32_t px, uint32_t py, uint32_t seed, uint32_t lane, float gx, float gy, WfState &S, PathHead &P) {
    S.rng = pcg_for_lane(seed, lane);
    CamRay c; camera_ray(sensor, px, py, S.rng, c);
    S.ray = c.ray;
    S.grad_d = (c.d_x - c.ray.d) * gx + (c.d_y - c.ray.d) * gy;                    // pixel motion → tangent of ray.d (epsm.py:250-257)
    wf_state_defaults(S);
    P.cam = c.ray.o; P.d = c.ray.d; P.dlduv0 = P.dlduv1 = 0.f; P.dldp1 = mk3(0, 0, 0); P.L = mk3(0, 0, 0);
}

// Gradient liveness.  calc_grad walks the logged vertices in order and stops producing outputs at the first vertex that
// ends the specular chain: `manifold` stops at the first vertex that is diffuse, not on a mesh or inactive
// (epsm.py:793-796,855,862,920,928 — `valid`, `hasdiffuse > 0`, `active` masks); `manifold_caustic` needs a diffuse
// first vertex (epsm.py:999,1008) and then stops at the second diffuse / non-mesh / inactive vertex.  That terminating
// vertex is still read as the "next vertex" of section B (position, triangle, flags), but nothing behind it and none
// of its own emitter / BSDF data reach path_grad, light_grad or diffuse_grad, which are the only things
// render_backward's scatter consumes.  With `prune` the replay therefore stops there (and skips emitter sampling,
// shadow and light-direction rays wherever their results cannot reach an output); the gradients are unchanged and
// tests hold the pruned pass to the full replay.  Every random number is still drawn in the reference order.
EPSM_HD bool wf_chain_vertex(const WfState &S, int variant, uint32_t it, bool vact, uint32_t ismesh, uint32_t flags) {
    const bool diffuse = (flags & EPSM_BF_DIFFUSE) != 0;
    if (variant == EPSM_VARIANT_MANIFOLD) return S.chain && vact && ismesh && !diffuse;
    return S.chain && vact && ismesh && (it == 0 ? diffuse : !diffuse);
}

// After the closest hit `h` of S.ray: surface interaction, first-hit tangents, emission, emitter sampling up to
// (not including) its visibility test.  Returns true when S.shadow_ray must be tested.
// `si` returns the surface interaction of the hit (the first-vertex fast path of k_wf_stage_a scatters from it).
EPSM_HDN bool wf_stage_a(const DScene &sc, WfState &S, PathHead &P, const Hit &h_in, uint32_t it, uint32_t max_depth_prop, int variant, bool prune, SI &si) {
    Hit h = h_in;
    if (!S.active) { h.t = INFINITY; h.u = h.v = 0.f; h.prim = 0xffffffffu; }
    compute_si(sc, S.ray, h, si);
    if (it == 0 && si.valid) {
        // forward-mode derivative of the first hit w.r.t. ray.d with the triangle detached (epsm.py:263-270)
        f3 p0 = ld3(sc.pos, si.fi[0]), e1 = ld3(sc.pos, si.fi[1]) - p0, e2 = ld3(sc.pos, si.fi[2]) - p0;
        f3 pvec = cross3(S.ray.d, e2), tvec = S.ray.o - p0, qvec = cross3(tvec, e1);
        float det = dot3(e1, pvec), inv_det = 1.f / det;
        f3 dpvec = cross3(S.grad_d, e2);
        float ddet = dot3(e1, dpvec), dinv = -ddet * inv_det * inv_det;
        float du = dot3(tvec, dpvec) * inv_det + dot3(tvec, pvec) * dinv;
        float dv = dot3(S.grad_d, qvec) * inv_det + dot3(S.ray.d, qvec) * dinv;
        if (si.ismesh) { P.dlduv0 = -(du + dv); P.dlduv1 = du; }                   // b0 = 1-u-v, b1 = u
        P.dldp1 = e1 * du + e2 * dv;
    }
    float em_pdf = pdf_emitter_direction(sc, si, S.prev_p, !S.prev_delta);
    float mis = mis_weight(S.prev_pdf, em_pdf);
    f3 Le = S.beta * mis * emitter_eval(sc, si);
    S.L = S.L + Le;
    S.active_next = (S.depth + 1 < max_depth_prop) && si.valid;
    uint32_t flags = si.valid ? sc.bsdfs[si.bsdf].bflags : 0u;
    const bool vact = S.active && si.valid;
    S.chain_vertex = wf_chain_vertex(S, variant, it, vact, si.ismesh, flags);
    bool need_nee = true, need_occl = true;
    if (prune) {
        // active_em is read by section A of `manifold` only (epsm.py:855-866); ds.p and si_direct of the first
        // vertex feed the shadow seed when max_depth ≤ 3 (epsm.py:609-620)
        need_occl = variant == EPSM_VARIANT_MANIFOLD && S.chain_vertex;
        need_nee = need_occl || (it == 0 && vact && wf_iterations(max_depth_prop) <= 3);
    }
    S.smooth = S.active_next && (flags & EPSM_BF_SMOOTH) && need_nee;
    float e1r = pcg_f32(S.rng), e2r = pcg_f32(S.rng);
    S.need_shadow = false;
    // the visibility query is deferred: record the shadow ray, assume visible, fix up in stage B (scene.cpp:259-263)
    sample_emitter_direction(sc, si, e1r, e2r, S.smooth, S.ds, S.em_weight, [&](const Ray &r) { if (need_occl) { S.shadow_ray = r; S.need_shadow = true; } return false; });
    return S.need_shadow;
}

EPSM_HDN bool wf_stage_a(const DScene &sc, WfState &S, PathHead &P, const Hit &h_in, uint32_t it, uint32_t max_depth_prop, int variant, bool prune) {
    SI si; return wf_stage_a(sc, S, P, h_in, it, max_depth_prop, variant, prune, si);
}

// A path whose first vertex already ends the gradient chain and needs no light-direction ray (see wf_chain_vertex and
// need_nee above) has exactly one output: diffuse_grad[0] = dldp1 on a diffuse first hit (epsm.py:789-792, :1056 in
// the caustic variant), ±0.1 outlier filter applied (epsm.py:933-944), scattered by seed (ii) onto the hit triangle.
EPSM_HD bool wf_first_vertex_final(const WfState &S, uint32_t it, bool prune) { return prune && it == 0 && !S.chain_vertex && !S.smooth; }
// … its value and whether the hit mesh takes gradients (the scatter targets are si.fi[c] with weights si.bw[c]).
EPSM_HD bool wf_first_vertex_seed(const DScene &sc, const WfState &S, const PathHead &H, const SI &si, f3 &dg) {
    const bool diffuse = si.valid && (sc.bsdfs[si.bsdf].bflags & EPSM_BF_DIFFUSE) != 0;
    dg = outlier3(diffuse ? H.dldp1 : mk3(0, 0, 0));
    return S.active && si.valid && (si.mesh_flags & EPSM_MESH_GRAD);
}

// After the visibility test: NEE contribution, the two BSDF samples, the vertex log, the ray towards the light
// sample (returned in `direct`, traced later as a closest-hit query) and the next path ray.
// `v` receives the logged vertex when it < EPSM_MAXV (epsm.py:648).
EPSM_HDN void wf_stage_b(const DScene &sc, WfState &S, VLog &v, const Hit &h_in, bool occluded, uint32_t it, uint32_t rr_depth, bool prune,
                         Ray &direct, bool &direct_valid) {
    Hit h = h_in;
    if (!S.active) { h.t = INFINITY; h.u = h.v = 0.f; h.prim = 0xffffffffu; }
    SI si; compute_si(sc, S.ray, h, si);
    DBsdf b; uint32_t flags = 0;
    if (si.valid) { b = bsdf_at(sc, si.bsdf, si.uv); flags = b.bflags; }
    if (S.need_shadow && occluded) { S.em_weight = mk3(0, 0, 0); S.ds.pdf = 0.f; }
    bool active_em = S.smooth && S.ds.pdf != 0.f;
    f3 Lr_dir = mk3(0, 0, 0), dir_to_light = mk3(0, 0, 0);
    if (S.smooth) dir_to_light = norm3(S.ds.p - si.p);                             // epsm.py:596
    if (active_em) {
        // Backward mode re-evaluates the emitter for the recomputed direction (epsm.py:596-598)
        f3 em_val = dot3(dir_to_light, S.ds.n) < 0.f ? a3(sc.emitters[S.ds.emitter].radiance) : mk3(0, 0, 0);
        f3 emw = em_val * (1.0f / S.ds.pdf);
        f3 wo = to_local(si, dir_to_light); f3 val; float bpdf;
        bsdf_eval_pdf(b, si.wi, wo, val, bpdf);
        Lr_dir = S.beta * mis_weight(S.ds.pdf, bpdf) * val * emw;
    }
    pcg_f32(S.rng); pcg_f32(S.rng); pcg_f32(S.rng);                                // discarded detached sample (epsm.py:633-636)
    float s1 = pcg_f32(S.rng), s2x = pcg_f32(S.rng), s2y = pcg_f32(S.rng);
    BSample bs; bs.wo = bs.hf = bs.weight = mk3(0, 0, 0); bs.pdf = 0.f; bs.eta = 0.f; bs.type = 0;
    if (si.valid) bsdf_sample(b, si.wi, s1, s2x, s2y, S.active_next, bs);
    direct_valid = false;
    if (it < EPSM_MAXV) {                                                          // epsm.py:648-654
        v.active = S.active && si.valid; v.flags = flags; v.ismesh = si.ismesh; v.light = S.ds.p; v.active_em = active_em;
        v.p0 = si.p0; v.p1 = si.p1; v.p2 = si.p2; v.p = si.p; v.b0 = si.b0; v.b1 = si.b1;
        v.n0 = si.n0; v.n1 = si.n1; v.n2 = si.n2; v.eta = bs.eta; v.hf = bs.hf;
        v.prim = si.prim; v.mesh_flags = si.mesh_flags; v.has_normals = si.has_normals;
        for (int k = 0; k < 3; ++k) { v.fi[k] = si.fi[k]; v.bw[k] = si.bw[k]; v.dfi[k] = 0; v.dbw[k] = 0.f; }
        v.lr_sum = Lr_dir.x + Lr_dir.y + Lr_dir.z;
        v.direct_valid = v.direct_any = v.direct_grad = false; v.dis = 0.f;
        v.direct_traced = S.smooth;
        if (S.smooth) { direct = spawn_ray(si, dir_to_light); direct_valid = true; }    // si_direct, epsm.py:609-625
    }
    S.L = S.L + Lr_dir;
    if (si.valid) S.ray = spawn_ray(si, to_world(si, bs.wo));
    S.eta *= bs.eta; S.beta = S.beta * bs.weight;
    S.prev_p = si.p; S.prev_pdf = bs.pdf; S.prev_delta = (bs.type & EPSM_BF_DELTA) != 0;
    float bmax = max3(S.beta);
    bool active_next = S.active_next && bmax != 0.f;
    float rr_prob = fminf(bmax * S.eta * S.eta, .95f);
    bool rr_active = S.depth >= rr_depth;
    if (rr_active) S.beta = S.beta * rcp_e(rr_prob);
    bool rr_continue = pcg_f32(S.rng) < rr_prob;
    active_next = active_next && (!rr_active || rr_continue);
    if (si.valid) S.depth += 1;
    // the gradient chain ends at this vertex: nothing behind it reaches an output of calc_grad (see wf_chain_vertex)
    S.chain = S.chain && S.chain_vertex;
    if (prune && !S.chain) active_next = false;
    S.active = active_next;
}

// Closest hit of the ray towards the light sample (FollowShape): barycentric weights on the hit triangle, the
// shadow-term distance ratio of epsm.py:614-615.
EPSM_HD void wf_direct_hit(const DScene &sc, VLog &v, const Hit &hd) {
    if (hd.prim == 0xffffffffu) return;
    const uint32_t g0 = sc.faces[3 * hd.prim], g1 = sc.faces[3 * hd.prim + 1], g2 = sc.faces[3 * hd.prim + 2];
    const float w1 = hd.u, w2 = hd.v, w0 = 1.f - w1 - w2;
    f3 q = fma3(ld3(sc.pos, g0), w0, fma3(ld3(sc.pos, g1), w1, ld3(sc.pos, g2) * w2));
    v.direct_any = true; v.direct_valid = v.active_em;
    v.direct_grad = (sc.meshes[sc.face_mesh[hd.prim]].flags & EPSM_MESH_GRAD) != 0;
    v.dfi[0] = g0; v.dfi[1] = g1; v.dfi[2] = g2; v.dbw[0] = w0; v.dbw[1] = w1; v.dbw[2] = w2;
    float dis = len3(v.light - q) / len3(v.light - v.p);
    v.dis = dis < 0.01f ? 0.f : dis;
}


// The whole replay for one lane with immediate scene queries (the test-only host harness and the record probe use
// this; the production kernels run the same stages as a wavefront).
template <bool COUNT>
EPSM_HDN void path_log(const DScene &sc, Tracer &tr, const DSensor &sensor, uint32_t px, uint32_t py, uint32_t seed, uint32_t lane,
                       float gx, float gy, uint32_t max_depth_prop, uint32_t rr_depth, int variant, bool prune, PathLog &P) {
    WfState S; wf_begin(sensor, px, py, seed, lane, gx, gy, S, P);
    const uint32_t max_iter = wf_replay_iterations(max_depth_prop, prune);
    uint32_t it = 0;
    for (; it < max_iter; ++it) {
        if (prune && !S.active) break;                 // the remaining logged vertices are inactive
        Hit h; h.t = INFINITY; h.u = h.v = 0.f; h.prim = 0xffffffffu;
        if (S.active) h = trace_closest<COUNT>(tr, S.ray);
        bool occluded = false;
        if (wf_stage_a(sc, S, P, h, it, max_depth_prop, variant, prune)) occluded = trace_any<COUNT>(tr, S.shadow_ray);
        Ray direct; bool dv; VLog scratch;
        VLog &v = it < EPSM_MAXV ? P.v[it] : scratch;
        wf_stage_b(sc, S, v, h, occluded, it, rr_depth, prune, direct, dv);
        if (dv) wf_direct_hit(sc, v, trace_closest<COUNT>(tr, direct));
    }
    for (; it < EPSM_MAXV; ++it) { memset(&P.v[it], 0, sizeof(VLog)); P.v[it].prim = 0xffffffffu; }
    P.L = S.L;
}

// ---- the manifold constraint and its Jacobian (epsm.py:805-821, 868-883) ---------------------------
// res = normalize(R·ŵi + η·R·ŵo) with R = [t; b; n̂] built from the interpolated normal (epsm.py:746-756).
// Outputs, for the two constrained components i ∈ {0,1}, the gradients of res[i] w.r.t. the previous
// vertex position, this vertex position, the next (or light) position and the interpolated normal.
struct RowJac { f3 gp[2], gc[2], gx[2], gn[2]; };

EPSM_HD void frame_from_normal(f3 n, f3 &t, f3 &b, f3 &nh, float &inv_ln, float &inv_lv) {
    float ln = len3(n); inv_ln = 1.0f / ln; nh = n * inv_ln;
    f3 v = mk3(0.f, -nh.z, nh.y); float lv = len3(v); inv_lv = 1.0f / lv; t = v * inv_lv;
    b = cross3(nh, t);
}
// gradient of L w.r.t. n given gradients w.r.t. the frame rows (gt, gb, gnh)
EPSM_HD f3 frame_backward(f3 t, f3 nh, float inv_ln, float inv_lv, f3 gt, f3 gb, f3 gnh) {
    // b = nh × t
    gnh = gnh + cross3(t, gb);
    gt = gt + cross3(gb, nh);
    // t = v/|v|, v = (0, -nh.z, nh.y)
    f3 gv = (gt - t * dot3(t, gt)) * inv_lv;
    gnh.z -= gv.y; gnh.y += gv.z;
    // nh = n/|n|
    return (gnh - nh * dot3(nh, gnh)) * inv_ln;
}
// half-vector constraint rows.  `attached`: differentiate through the frame (false in caustic section A, epsm.py:1022)
EPSM_HD void halfvector_jac(f3 xp, f3 xc, f3 xn, f3 n, float eta, bool attached, RowJac &J) {
    f3 t, b, nh; float inv_ln, inv_lv; frame_from_normal(n, t, b, nh, inv_ln, inv_lv);
    f3 di = xp - xc, dO = xn - xc;
    float ili = 1.0f / len3(di), ilo = 1.0f / len3(dO);
    f3 wi = di * ili, wo = dO * ilo;
    f3 wi2 = mk3(dot3(t, wi), dot3(b, wi), dot3(nh, wi)), wo2 = mk3(dot3(t, wo), dot3(b, wo), dot3(nh, wo));
    f3 hv = wi2 + wo2 * eta;
    float ilh = 1.0f / len3(hv); f3 res = hv * ilh;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        f3 e = i == 0 ? mk3(1, 0, 0) : mk3(0, 1, 0);
        float ri = i == 0 ? res.x : res.y;
        f3 a = (e - res * ri) * ilh;                       // ∂res_i/∂h
        // h = R wi + η R wo
        f3 gwi = t * a.x + b * a.y + nh * a.z;             // Rᵀ a
        f3 gwo = gwi * eta;
        f3 gdi = (gwi - wi * dot3(wi, gwi)) * ili, gdo = (gwo - wo * dot3(wo, gwo)) * ilo;
        J.gp[i] = gdi; J.gx[i] = gdo; J.gc[i] = -(gdi + gdo);
        if (attached) {
            f3 g = wi + wo * eta;                          // rows of R see this vector
            J.gn[i] = frame_backward(t, nh, inv_ln, inv_lv, g * a.x, g * a.y, g * a.z);
        } else J.gn[i] = mk3(0, 0, 0);
    }
}
// res2 = R·ŵo − const (epsm.py:1028,1116): rows of the "outgoing direction stays fixed" constraint
EPSM_HD void outdir_jac(f3 xc, f3 xn, f3 n, bool attached, RowJac &J) {
    f3 t, b, nh; float inv_ln, inv_lv; frame_from_normal(n, t, b, nh, inv_ln, inv_lv);
    f3 dO = xn - xc; float ilo = 1.0f / len3(dO); f3 wo = dO * ilo;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        f3 gwo = i == 0 ? t : b;
        f3 gdo = (gwo - wo * dot3(wo, gwo)) * ilo;
        J.gp[i] = mk3(0, 0, 0); J.gx[i] = gdo; J.gc[i] = -gdo;
        if (attached) J.gn[i] = frame_backward(t, nh, inv_ln, inv_lv, i == 0 ? wo : mk3(0, 0, 0), i == 1 ? wo : mk3(0, 0, 0), mk3(0, 0, 0));
        else J.gn[i] = mk3(0, 0, 0);
    }
}

// 2×2 helpers (row-major m[0] m[1]; m[2] m[3])
struct M2 { float a, b, c, d; };
EPSM_HD M2 m2_T(M2 m) { return M2{m.a, m.c, m.b, m.d}; }
EPSM_HD M2 m2_mul(M2 x, M2 y) { return M2{x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d}; }
EPSM_HD M2 m2_add(M2 x, M2 y) { return M2{x.a + y.a, x.b + y.b, x.c + y.c, x.d + y.d}; }
EPSM_HD M2 m2_neg(M2 x) { return M2{-x.a, -x.b, -x.c, -x.d}; }
EPSM_HD M2 m2_inv(M2 m) { float id = 1.0f / (m.a * m.d - m.b * m.c); return M2{m.d * id, -m.b * id, -m.c * id, m.a * id}; }
struct V2 { float x, y; };
EPSM_HD V2 m2_mv(M2 m, V2 v) { return V2{m.a * v.x + m.b * v.y, m.c * v.x + m.d * v.y}; }
EPSM_HD V2 v2_neg(V2 v) { return V2{-v.x, -v.y}; }
EPSM_HD V2 v2_add(V2 a, V2 b) { return V2{a.x + b.x, a.y + b.y}; }
// block of ∂c/∂u: rows i ∈ {0,1}, columns (b0, b1) — position part g·(p0−p2), g·(p1−p2)
EPSM_HD M2 uv_block(const f3 g[2], f3 a0, f3 a1) { return M2{dot3(g[0], a0), dot3(g[0], a1), dot3(g[1], a0), dot3(g[1], a1)}; }


template <int NV> struct PathGradsT {
    f3 pg[NV][5];                // path_grad[5k + {0,1,2: p0 p1 p2, 3: n, 4: m}]
    f3 lg[NV];                   // light_grad[k]
    f3 dg[NV];                   // diffuse_grad[k]
};
typedef PathGradsT<EPSM_MAXV> PathGrads;
// Loops over the vertex slots are fully unrolled: in the short-chain instantiation every index becomes static and the
// whole solve lives in registers (244, no stack frame); a counted `unroll (N)` keeps the loop and the local arrays.
#define EPSM_UNROLL_NV _Pragma("unroll")

EPSM_HD f3 vpoint(const VLog &v) { return v.p0 * v.b0 + v.p1 * v.b1 + v.p2 * (1.f - v.b0 - v.b1); }
EPSM_HD f3 vnormal(const VLog &v) { return v.n0 * v.b0 + v.n1 * v.b1 + v.n2 * (1.f - v.b0 - v.b1); }

struct RowStore { RowJac J; M2 Lb, Db, Ub; };            // a stored section-B row block

// manifold: ManifoldIntegrator.calc_grad (epsm.py:745-946)
template <int NV>
EPSM_HD void calc_grad_manifold(const PathLogT<NV> &P, int n_logged, PathGradsT<NV> &G) {
    RowStore rows[NV];                                    // rows[r-1]: B-row of vertex r
    int hasdiffuse = 0; bool valid = true;
    const V2 rhs = V2{P.dlduv0, P.dlduv1};
    EPSM_UNROLL_NV
    for (int id = 1; id <= NV; ++id) {
        if (id > n_logged) break;
        const VLog &v = P.v[id - 1];
        const bool isdiffuse = (v.flags & EPSM_BF_DIFFUSE) != 0;
        if (id == 1) { G.dg[0] = isdiffuse ? P.dldp1 : mk3(0, 0, 0); valid = v.ismesh > 0; }
        else valid = valid && (v.ismesh > 0);
        hasdiffuse += isdiffuse ? 1 : 0;
        valid = valid && hasdiffuse < 2;
        // every output of this and all later vertices carries the masks valid ∧ active ∧ hasdiffuse == 0
        if (!valid || hasdiffuse > 0 || !v.active) break;
        const f3 xp = id == 1 ? P.cam : vpoint(P.v[id - 2]);
        const f3 xc = vpoint(v), n = vnormal(v);
        const f3 a0 = v.p0 - v.p2, a1 = v.p1 - v.p2, nn0 = v.n0 - v.n2, nn1 = v.n1 - v.n2;
        for (int sec = 0; sec < 2; ++sec) {
            const bool secB = sec == 1;
            if (secB && (id >= n_logged || id >= NV)) break;   // id < len(path_info) − 1 (epsm.py:868); NV: slots held
            const f3 xn = secB ? vpoint(P.v[id]) : v.light;
            RowStore cur;
            halfvector_jac(xp, xc, xn, n, v.eta, true, cur.J);
            cur.Db = m2_add(uv_block(cur.J.gc, a0, a1), uv_block(cur.J.gn, nn0, nn1));
            if (id >= 2) { const VLog &q = P.v[id - 2]; cur.Lb = uv_block(cur.J.gp, q.p0 - q.p2, q.p1 - q.p2); }
            else cur.Lb = M2{0, 0, 0, 0};
            if (secB) { const VLog &q = P.v[id]; cur.Ub = uv_block(cur.J.gx, q.p0 - q.p2, q.p1 - q.p2); }
            else cur.Ub = M2{0, 0, 0, 0};
            const bool ok = secB ? (P.v[id].active && (P.v[id].flags & EPSM_BF_DIFFUSE) != 0) : v.active_em;
            if (ok) {
                // Solve Cᵀ w = (dlduv, 0, …) for the block-tridiagonal C = [B-rows 1..id−1 ; this row]:
                // eliminate from the last block upwards (S_j = D_jᵀ + L_{j+1}ᵀ G_{j+1}, G_j = −S_j⁻¹ U_{j−1}ᵀ).
                M2 Gm[NV + 1];
                M2 S = m2_T(cur.Db);
                EPSM_UNROLL_NV
                for (int j = id; j >= 2; --j) {
                    Gm[j] = m2_neg(m2_mul(m2_inv(S), m2_T(rows[j - 2].Ub)));
                    const M2 Lj = j == id ? cur.Lb : rows[j - 1].Lb;
                    S = m2_add(m2_T(rows[j - 2].Db), m2_mul(m2_T(Lj), Gm[j]));
                }
                V2 w[NV + 1];
                w[1] = m2_mv(m2_inv(S), rhs);
                EPSM_UNROLL_NV
                for (int j = 2; j <= id; ++j) w[j] = m2_mv(Gm[j], w[j - 1]);
                // parameter gradients: θ ∈ {p0,p1,p2,n,m} of vertices 1..id
                EPSM_UNROLL_NV
                for (int j = 1; j <= id; ++j) {
                    const RowJac &Jj = j == id ? cur.J : rows[j - 1].J;
                    f3 a = Jj.gc[0] * w[j].x + Jj.gc[1] * w[j].y;
                    if (j + 1 <= id) { const RowJac &Jn = j + 1 == id ? cur.J : rows[j].J; a = a + (Jn.gp[0] * w[j + 1].x + Jn.gp[1] * w[j + 1].y); }
                    const VLog &q = P.v[j - 1];
                    G.pg[j - 1][0] = G.pg[j - 1][0] + ntn3(-(a * q.b0));
                    G.pg[j - 1][1] = G.pg[j - 1][1] + ntn3(-(a * q.b1));
                    G.pg[j - 1][2] = G.pg[j - 1][2] + ntn3(-(a * (1.f - q.b0 - q.b1)));
                    G.pg[j - 1][3] = G.pg[j - 1][3] + ntn3(-(Jj.gn[0] * w[j].x + Jj.gn[1] * w[j].y));
                    if (j < id || secB) G.pg[j - 1][4] = G.pg[j - 1][4] + ntn3(mk3(w[j].x, w[j].y, 0.f));
                }
                const f3 gx = ntn3(-(cur.J.gx[0] * w[id].x + cur.J.gx[1] * w[id].y));
                if (secB) G.dg[id] = gx; else G.lg[id - 1] = gx;
            }
            if (secB) rows[id - 1] = cur;
        }
    }
}

// manifold_caustic: ManifoldCausticIntegrator.calc_grad (epsm.py:952-1200).  Its section A never reaches an
// output (dlduv is zeroed unless vertex 1 is diffuse, epsm.py:999, and section A masks hasdiffuse > 0,
// epsm.py:1083,1090), so only section B is evaluated: rows 2..id are half-vector rows, row 1 (the diffuse
// vertex) is replaced by the fixed-outgoing-direction constraint of the current vertex (epsm.py:1141-1163).
template <int NV>
EPSM_HD void calc_grad_caustic(const PathLogT<NV> &P, int n_logged, PathGradsT<NV> &G) {
    RowStore rows[NV];
    const VLog &v1 = P.v[0];
    const bool d1 = (v1.flags & EPSM_BF_DIFFUSE) != 0;
    G.dg[0] = d1 ? P.dldp1 : mk3(0, 0, 0);
    if (!d1) return;
    const V2 rhs = V2{P.dlduv0, P.dlduv1};
    int hasdiffuse = 0; bool valid = true;
    EPSM_UNROLL_NV
    for (int id = 1; id < NV; ++id) {
        if (id >= n_logged) break;
        const VLog &v = P.v[id - 1];
        hasdiffuse += (v.flags & EPSM_BF_DIFFUSE) ? 1 : 0;
        valid = valid && (v.ismesh > 0) && hasdiffuse < 2;
        if (!valid || !v.active) break;
        const VLog &nx = P.v[id];
        const f3 xp = id == 1 ? P.cam : vpoint(P.v[id - 2]);
        const f3 xc = vpoint(v), xn = vpoint(nx), n = vnormal(v);
        const f3 a0 = v.p0 - v.p2, a1 = v.p1 - v.p2, nn0 = v.n0 - v.n2, nn1 = v.n1 - v.n2;
        RowStore cur;
        halfvector_jac(xp, xc, xn, n, v.eta, true, cur.J);
        cur.Db = m2_add(uv_block(cur.J.gc, a0, a1), uv_block(cur.J.gn, nn0, nn1));
        if (id >= 2) { const VLog &q = P.v[id - 2]; cur.Lb = uv_block(cur.J.gp, q.p0 - q.p2, q.p1 - q.p2); } else cur.Lb = M2{0, 0, 0, 0};
        cur.Ub = uv_block(cur.J.gx, nx.p0 - nx.p2, nx.p1 - nx.p2);
        rows[id - 1] = cur;
        RowJac J2; outdir_jac(xc, xn, n, true, J2);
        const M2 E = m2_add(uv_block(J2.gc, a0, a1), uv_block(J2.gn, nn0, nn1));
        const bool nd = (nx.flags & EPSM_BF_DIFFUSE) != 0, nnull = (nx.flags & EPSM_BF_NULL) != 0;
        if (!nx.active || !(nd || nnull)) continue;
        // Solve wᵀC = (dlduv, 0, …): with the rows ordered [2, 3, …, id, 1] the matrix is block upper
        // triangular with diagonal blocks L_2 … L_id, E — forward substitution over column blocks.
        V2 w[NV + 1];
        if (id == 1) {
            w[1] = m2_mv(m2_inv(m2_T(E)), rhs);
        } else {
            w[2] = m2_mv(m2_inv(m2_T(rows[1].Lb)), rhs);
            EPSM_UNROLL_NV
            for (int j = 2; j < id; ++j) {
                V2 acc = m2_mv(m2_T(rows[j - 1].Db), w[j]);
                if (j - 1 >= 2) acc = v2_add(acc, m2_mv(m2_T(rows[j - 2].Ub), w[j - 1]));
                w[j + 1] = v2_neg(m2_mv(m2_inv(m2_T(rows[j].Lb)), acc));
            }
            V2 acc = m2_mv(m2_T(rows[id - 1].Db), w[id]);
            if (id - 1 >= 2) acc = v2_add(acc, m2_mv(m2_T(rows[id - 2].Ub), w[id - 1]));
            w[1] = v2_neg(m2_mv(m2_inv(m2_T(E)), acc));
        }
        if (nd) {
            EPSM_UNROLL_NV
            for (int j = 1; j <= id; ++j) {
                // half-vector rows exist for r ≥ 2 (r = 1 only when id == 1 … then row 1 is the res2 row)
                f3 a = mk3(0, 0, 0), gn = mk3(0, 0, 0), gm = mk3(0, 0, 0);
                if (j >= 2) { const RowJac &Jj = rows[j - 1].J; a = a + (Jj.gc[0] * w[j].x + Jj.gc[1] * w[j].y); gn = gn + (Jj.gn[0] * w[j].x + Jj.gn[1] * w[j].y); gm = mk3(w[j].x, w[j].y, 0.f); }
                if (j + 1 <= id && j + 1 >= 2) { const RowJac &Jn = rows[j].J; a = a + (Jn.gp[0] * w[j + 1].x + Jn.gp[1] * w[j + 1].y); }
                if (j == id) { a = a + (J2.gc[0] * w[1].x + J2.gc[1] * w[1].y); gn = gn + (J2.gn[0] * w[1].x + J2.gn[1] * w[1].y); }
                const VLog &q = P.v[j - 1];
                G.pg[j - 1][0] = G.pg[j - 1][0] + ntn3(-(a * q.b0));
                G.pg[j - 1][1] = G.pg[j - 1][1] + ntn3(-(a * q.b1));
                G.pg[j - 1][2] = G.pg[j - 1][2] + ntn3(-(a * (1.f - q.b0 - q.b1)));
                G.pg[j - 1][3] = G.pg[j - 1][3] + ntn3(-gn);
                G.pg[j - 1][4] = G.pg[j - 1][4] + ntn3(gm);
            }
        }
        f3 gx = J2.gx[0] * w[1].x + J2.gx[1] * w[1].y;
        // Reference quirk kept for parity: in section B, point_next.grad is not cleared after the res2[:,0]
        // backward (epsm.py:1154-1157 vs :1062-1064), so the row of res[:,1] in param_diffuse_grad also
        // contains ∂res2[0]/∂x_next (epsm.py:1139).
        if (id >= 2) gx = gx + (cur.J.gx[0] * w[id].x + (cur.J.gx[1] + J2.gx[0]) * w[id].y);
        G.dg[id] = ntn3(-gx);
    }
}

// `n_avail` (optional): only vertices 0 … n_avail−1 of the log are active — outputs are produced for those alone, so the
// zero-fill and the outlier pass stop there (the gradient pass passes the loaded prefix; default: all).
template <int NV>
EPSM_HD void calc_grad(const PathLogT<NV> &P, int variant, int n_logged, PathGradsT<NV> &G, int n_avail = NV) {
    EPSM_UNROLL_NV
    for (int k = 0; k < NV; ++k) { if (k >= n_avail) break; for (int c = 0; c < 5; ++c) G.pg[k][c] = mk3(0, 0, 0); G.lg[k] = G.dg[k] = mk3(0, 0, 0); }
    if (variant == EPSM_VARIANT_MANIFOLD) calc_grad_manifold<NV>(P, n_logged, G);
    else calc_grad_caustic<NV>(P, n_logged, G);
    EPSM_UNROLL_NV
    for (int k = 0; k < NV; ++k) { if (k >= n_avail) break; for (int c = 0; c < 5; ++c) G.pg[k][c] = outlier3(G.pg[k][c]); G.lg[k] = outlier3(G.lg[k]); G.dg[k] = outlier3(G.dg[k]); }
}
