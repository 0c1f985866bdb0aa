// ref_probe.cpp — TEST INFRASTRUCTURE ONLY.  A thin harness over the REFERENCE's own header-only code, compiled
// from the headers where they lie under /root/reference (never copied into this repository) by `make -C oracle ref`
// into oracle/_ref/libepsm_refprobe.so.  It is used (a) to validate the CPU restatement in epsm_oracle.cpp function
// by function on seeded inputs and (b) to generate the golden vectors tests/golden/ref_probe.npz
// (tests/golden/make_ref_probe.py).  Nothing in the product path loads it.
//
// What it reaches (all `scalar_rgb` instantiations — the variant whose transcendental functions are Dr.Jit's own
// polynomials, like llvm_ad_rgb):
//   fn 0  sample_tea_32                        include/mitsuba/core/random.h:76-90
//   fn 1  PCG32 stream, sampler seeding        ext/drjit/include/drjit/random.h:55-132 + src/render/sampler.cpp:122-130
//   fn 2  dr::sincos                           ext/drjit/include/drjit/math.h:76-186
//   fn 3  warp::square_to_cosine_hemisphere    include/mitsuba/core/warp.h:54-95,320-328
//   fn 4  coordinate_system                    include/mitsuba/core/vector.h:116-136
//   fn 5  fresnel (dielectric)                 include/mitsuba/render/fresnel.h:35-72
//   fn 6  fresnel_conductor                    include/mitsuba/render/fresnel.h:93-117
//   fn 7  MicrofacetDistribution::sample       include/mitsuba/render/microfacet.h:244-327
//   fn 8  MicrofacetDistribution::eval         include/mitsuba/render/microfacet.h:183-210
//   fn 9  MicrofacetDistribution::smith_g1     include/mitsuba/render/microfacet.h:365-392
//   fn 10 warp::square_to_uniform_triangle     include/mitsuba/core/warp.h:153-156
//   fn 11 reflect(wi, m)                       include/mitsuba/render/fresnel.h:282-284
//   fn 12 Frame3f(n) + to_local / to_world     include/mitsuba/core/frame.h:21-50
// The triangle test, Mesh::compute_surface_interaction, the BSDF plugins, the area emitter and the film live in
// .cpp files behind libmitsuba's object model (cmake build) and are pinned by tests/golden/reference_kats.json instead.
// Instantiation note.  Float = plain `float` evaluates every explicit dr::fmadd fused (as the LLVM backend's llvm.fma) and
// everything else unfused (-ffp-contract=off; the LLVM backend emits plain fmul/fadd), which is the arithmetic of
// llvm_ad_rgb — except for dr::dot / squared_norm / normalize on static arrays: with array-valued entries (the JIT
// variants) Dr.Jit evaluates them as an fmadd chain (ext/drjit/include/drjit/array_base.h:603-631) whereas plain-float
// entries take the SSE dp_ps route (packet_sse42.h:370).  The restatement follows the JIT form, so functions built on
// dot() (fn 3, 11, 12) agree with this probe to the last 1–2 ulp instead of bit for bit; all others are bit-exact
// (tests/test_ref_probe.py states the tolerance per function).  Dr.Jit's 1-wide host arrays are no way around this:
// their generic fmadd_ is the unfused `a*b+c` (array_base.h:429-460,509).
#include <mitsuba/core/random.h>
#include <mitsuba/core/warp.h>
#include <mitsuba/core/frame.h>
#include <mitsuba/core/vector.h>
#include <mitsuba/render/fresnel.h>
#include <mitsuba/render/microfacet.h>
#include <drjit/random.h>
#include <drjit/math.h>
#include <cstdint>
#include <cstring>

using namespace mitsuba;
namespace dr = drjit;
using Spectrum3 = Color<float, 3>;
using MF = MicrofacetDistribution<float, Spectrum3>;
using V3f = Vector<float, 3>;
using N3f = Normal<float, 3>;
using P2f = Point<float, 2>;

static MF make_mf(const float *in) {
    return MF(in[2] != 0.f ? MicrofacetType::GGX : MicrofacetType::Beckmann, in[0], in[1], true);
}

extern "C" {

// in: n × in_stride floats, out: n × out_stride floats (integer functions pass bit patterns).
int ref_probe(int fn, uint64_t n, const float *in_, int in_stride, float *out_, int out_stride) {
    for (uint64_t i = 0; i < n; ++i) {
        const float *in = in_ + i * (uint64_t) in_stride; float *out = out_ + i * (uint64_t) out_stride;
        switch (fn) {
        case 0: {
            uint32_t v0, v1; memcpy(&v0, in, 4); memcpy(&v1, in + 1, 4);
            auto r = sample_tea_32<uint32_t>(v0, v1, 4);
            memcpy(out, &r.first, 4); memcpy(out + 1, &r.second, 4);
        } break;
        case 2: { auto [s, c] = dr::sincos(in[0]); out[0] = s; out[1] = c; } break;
        case 3: { V3f v = warp::square_to_cosine_hemisphere(P2f(in[0], in[1])); out[0] = v.x(); out[1] = v.y(); out[2] = v.z(); } break;
        case 4: {
            auto [s, t] = coordinate_system(V3f(in[0], in[1], in[2]));
            out[0] = s.x(); out[1] = s.y(); out[2] = s.z(); out[3] = t.x(); out[4] = t.y(); out[5] = t.z();
        } break;
        case 5: {
            auto [r, cos_t, eta_it, eta_ti] = fresnel(in[0], in[1]);
            out[0] = r; out[1] = cos_t; out[2] = eta_it; out[3] = eta_ti;
        } break;
        case 6: out[0] = fresnel_conductor(in[0], dr::Complex<float>(in[1], in[2])); break;
        case 7: {
            MF d = make_mf(in);
            auto [m, pdf] = d.sample(V3f(in[3], in[4], in[5]), P2f(in[6], in[7]));
            out[0] = m.x(); out[1] = m.y(); out[2] = m.z(); out[3] = pdf;
        } break;
        case 8: { MF d = make_mf(in); out[0] = d.eval(V3f(in[3], in[4], in[5])); } break;
        case 9: { MF d = make_mf(in); out[0] = d.smith_g1(V3f(in[3], in[4], in[5]), V3f(in[6], in[7], in[8])); } break;
        case 10: { P2f b = warp::square_to_uniform_triangle(P2f(in[0], in[1])); out[0] = b.x(); out[1] = b.y(); } break;
        case 11: { V3f r = reflect(V3f(in[0], in[1], in[2]), N3f(in[3], in[4], in[5])); out[0] = r.x(); out[1] = r.y(); out[2] = r.z(); } break;
        case 12: {
            Frame<float> f(N3f(in[0], in[1], in[2])); V3f v(in[3], in[4], in[5]);
            V3f l = f.to_local(v), w = f.to_world(v);
            out[0] = l.x(); out[1] = l.y(); out[2] = l.z(); out[3] = w.x(); out[4] = w.y(); out[5] = w.z();
        } break;
        default: return -1;
        }
    }
    return 0;
}

// fn 1: the first n_floats of lanes [lane_begin, lane_begin + n_lanes) of the independent sampler's JIT-variant
// seeding (sampler.cpp:122-130: idx → TEA(seed_value, idx) → PCG32::seed(initstate=v0, initseq=v1)), each lane
// driven through the reference's own PCG32 and TEA.
int ref_sampler_stream(uint32_t seed_value, uint64_t lane_begin, uint64_t n_lanes, uint32_t n_floats, float *out) {
    for (uint64_t i = 0; i < n_lanes; ++i) {
        auto t = sample_tea_32<uint32_t>(seed_value, (uint32_t) (lane_begin + i), 4);
        dr::PCG32<uint32_t> rng; rng.seed(1, t.first, t.second);
        for (uint32_t k = 0; k < n_floats; ++k) out[i * n_floats + k] = rng.next_float32();
    }
    return 0;
}

} // extern "C"
