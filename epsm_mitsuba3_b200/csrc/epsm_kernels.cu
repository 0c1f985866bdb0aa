// epsm_kernels.cu — sm_100a kernels of the EPSM manifold path and their launchers.
//
//   k_trace            Scene::ray_intersect_preliminary / ray_test on a ray batch (BVH4, one thread per ray)
//   k_render           EPSMIntegrator.render as one kernel: ray gen + primal path loop + tile-local film (EPSM_PRIMAL_MEGAKERNEL=1)
//   k_pr_stage / k_pr_film   the default primal render: one stage per bounce over the persistent traversal kernel, tile-local film
//   k_film_develop     HDRFilm::develop + the two zero channels of epsm.py:76-82
//   k_wf_trace / k_wf_trace_primary   persistent BVH4 traversal over a ray queue / over the camera rays of a chunk
//   k_wf_stage_a / k_wf_stage_b / k_wf_finish   EPSMIntegrator.render_backward as a wavefront: sample_path replay cut at every
//                      scene query, path log in HBM (word-major), manifold solve in registers, warp-aggregated scatter
//   k_prb_pass / k_prb_backward   prb_reparam's reverse pass (hybrid phase 2): replay wavefront with checkpoints / per-lane megakernel
//   k_sampler_stream / k_sample_rays   parity probes for the RNG and ray-generation contracts
//   k_normals_*        Mesh::recompute_vertex_normals (mesh.cpp:257-330) and its adjoint
//   k_refit_tris / k_refit_nodes   per-iteration BVH refit
//
// Work decomposition: a CTA of 128 threads owns an 8×4 pixel tile; thread i handles pixel (i & 31)
// of the tile and the samples s = (i >> 5) + 4k of that pixel, so a warp is always 32 neighbouring
// pixels at the same sample index (coherent primary rays) and the lane id — which alone determines
// the random stream (sampler.cpp:122-130) — is (pixel · spp + s) exactly as in the reference
// wavefront (common.py:326-343).
#include <cuda_runtime.h>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <cooperative_groups.h>
#include <stdlib.h>
#include <stdint.h>
#include "epsm_device.cuh"
#include "epsm_prb.cuh"
#include "epsm_launch.h"

#define TILE_W 8
#define TILE_H 4
#define CTA 128
#ifndef EPSM_PATH_MINBLOCKS
#define EPSM_PATH_MINBLOCKS 2      // resident CTAs per SM requested for the path kernels (register cap = 65536 / (128·N))
#endif

// ------------------------------------------------------------------------------------------------
//  Shared-memory traversal stack for a CTA
// ------------------------------------------------------------------------------------------------
struct SharedStack { uint32_t ref[EPSM_SSTACK][CTA]; float t[EPSM_SSTACK][CTA]; };

__device__ __forceinline__ void bind_stack(TStack &st, SharedStack &ss) {
    st.s_ref = &ss.ref[0][threadIdx.x]; st.s_t = &ss.t[0][threadIdx.x]; st.stride = CTA; st.sp = 0;
}

__device__ __forceinline__ void flush_counters(const Tracer &tr, uint32_t lanes, unsigned long long *counters) {
    unsigned long long a = tr.n_nodes, b = tr.n_tris, c = tr.n_rays, d = lanes;
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o);
        c += __shfl_down_sync(0xffffffffu, c, o); d += __shfl_down_sync(0xffffffffu, d, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(counters, a); atomicAdd(counters + 1, b); atomicAdd(counters + 2, c); atomicAdd(counters + 3, d); }
}

// ------------------------------------------------------------------------------------------------
//  Ray batches
// ------------------------------------------------------------------------------------------------
template <bool ANY, bool COUNT>
__global__ void __launch_bounds__(CTA) k_trace(const __grid_constant__ DScene sc, const f4 *__restrict__ rays, uint64_t n, uint32_t *__restrict__ hits,
                                               unsigned long long *counters) {
    __shared__ SharedStack ss;
    TStack st; bind_stack(st, ss);
    Tracer tr(sc, st);
    uint64_t i = (uint64_t) blockIdx.x * CTA + threadIdx.x;
    if (i < n) {
        f4 a = ldg4(rays + 2 * i), b = ldg4(rays + 2 * i + 1);
        Ray r{mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), a.w};
        if (ANY) {
            hits[i] = trace_any<COUNT>(tr, r) ? 1u : 0u;
        } else {
            Hit h = trace_closest<COUNT>(tr, r);
            uint4 o = make_uint4(f2u(h.t), f2u(h.u), f2u(h.v), h.prim);
            reinterpret_cast<uint4 *>(hits)[i] = o;
        }
    }
    if (COUNT) flush_counters(tr, i < n ? 1u : 0u, counters);
}

// ------------------------------------------------------------------------------------------------
//  Primal render
// ------------------------------------------------------------------------------------------------
// Tile-local film block: 2-pixel apron below, 3 above.  The sample position px + ξ is an fp32 sum and rounds up to
// px + 1 for ξ within half an ulp of 1 (≈ 3e-5 of the samples at px ≈ 1000), which moves the 5×5 footprint one pixel on.
#define FT_W (TILE_W + 5)
#define FT_H (TILE_H + 5)

template <bool COUNT>
__global__ void __launch_bounds__(CTA, EPSM_PATH_MINBLOCKS) k_render(const __grid_constant__ DScene sc, const __grid_constant__ DLaunch ln, float *__restrict__ film, unsigned long long *counters) {
    __shared__ SharedStack ss;
    __shared__ float ftile[FT_H][FT_W][4];
    TStack st; bind_stack(st, ss);
    Tracer tr(sc, st);
    for (int i = threadIdx.x; i < FT_H * FT_W * 4; i += CTA) (&ftile[0][0][0])[i] = 0.f;
    __syncthreads();
    const uint32_t W = ln.sensor.w, H = ln.sensor.h;                  // sampling grid (film + border)
    const int B = ln.sensor.border, FW = (int) ln.sensor.fw, FH = (int) ln.sensor.fh;
    const uint32_t tiles_x = (W + TILE_W - 1) / TILE_W;
    const uint32_t tx = blockIdx.x % tiles_x, ty = ln.tile_y0 + blockIdx.x / tiles_x;
    const uint32_t pin = threadIdx.x & 31, slot = threadIdx.x >> 5;
    const uint32_t px = tx * TILE_W + (pin & (TILE_W - 1)), py = ty * TILE_H + pin / TILE_W;
    uint32_t lanes = 0;
    if (px < W && py < H) {
        const uint64_t pixel = (uint64_t) py * W + px;
        for (uint32_t s = ln.s_begin + slot; s < ln.s_end; s += CTA / 32) {
            const uint64_t lane = pixel * ln.spp + s;
            if (lane < ln.lane_begin || lane >= ln.lane_end) continue;
            Pcg rng = pcg_for_lane(ln.seed, (uint32_t) lane);
            CamRay c; camera_ray(ln.sensor, px, py, rng, c);
            bool valid;
            f3 L = path_primal<COUNT>(sc, tr, c.ray, rng, ln.max_depth, ln.rr_depth, valid);
            ++lanes;
            // ImageBlock::put, coalesced branch (imageblock.cpp:414-494): 5×5 footprint, separable weights
            const int pix = (int) floorf(c.pos_x) - 2, piy = (int) floorf(c.pos_y) - 2;
            float rel_x = (float) pix + .5f - c.pos_x, rel_y = (float) piy + .5f - c.pos_y;
            float wx[5], wy[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) { wx[k] = filter_eval(sc.filter, rel_x); wy[k] = filter_eval(sc.filter, rel_y); rel_x += 1; rel_y += 1; }
            const int lx = pix - ((int) (tx * TILE_W) - B) + 2, ly = piy - ((int) (ty * TILE_H) - B) + 2;   // tile-local, 2-px apron
#pragma unroll
            for (int ys = 0; ys < 5; ++ys)
#pragma unroll
                for (int xs = 0; xs < 5; ++xs) {
                    const float w = wy[ys] * wx[xs];
                    float *d = ftile[ly + ys][lx + xs];
                    atomicAdd(d + 0, L.x * w); atomicAdd(d + 1, L.y * w); atomicAdd(d + 2, L.z * w); atomicAdd(d + 3, 1.0f * w);
                }
        }
    }
    __syncthreads();
    // flush the tile (with apron) into the RGBW film; pixels outside the film are the cropped border
    for (int i = threadIdx.x; i < FT_H * FT_W; i += CTA) {
        const int fy = (int) (ty * TILE_H) - B + i / FT_W - 2, fx = (int) (tx * TILE_W) - B + i % FT_W - 2;      // film coordinates
        if (fx < 0 || fy < 0 || fx >= FW || fy >= FH) continue;
        const float *s = ftile[i / FT_W][i % FT_W];
        if (s[3] == 0.f && s[0] == 0.f && s[1] == 0.f && s[2] == 0.f) continue;
        float *d = film + ((size_t) fy * FW + fx) * 4;
        atomicAdd(d + 0, s[0]); atomicAdd(d + 1, s[1]); atomicAdd(d + 2, s[2]); atomicAdd(d + 3, s[3]);
    }
    if (COUNT) flush_counters(tr, lanes, counters);
}

__global__ void k_film_develop(const float *__restrict__ film, uint64_t n_pixels, float *__restrict__ out) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    float4 v = reinterpret_cast<const float4 *>(film)[i];
    float d = v.w == 0.f ? 1.f : v.w;                                             // hdrfilm.cpp:372
    out[5 * i + 0] = v.x / d; out[5 * i + 1] = v.y / d; out[5 * i + 2] = v.z / d; out[5 * i + 3] = 0.f; out[5 * i + 4] = 0.f;
}

// ------------------------------------------------------------------------------------------------
//  Warp-aggregated gradient scatter (the adjoint of the vertex gathers; cf. the reference's scatter_reduce
//  codegen, ext/drjit/ext/drjit-core/src/cuda_eval.cpp:808-905)
// ------------------------------------------------------------------------------------------------
// BFLY: also test for the whole warp adding to one vertex (worth it where that is common — the first-vertex fast path and the
// short-chain finish kernel; the 255-register long-chain finish kernel only pays registers for it)
template <bool BFLY = true>
__device__ __forceinline__ void scatter3(float *__restrict__ buf, uint32_t idx, f3 v, bool pred) {
    pred = pred && (v.x != 0.f || v.y != 0.f || v.z != 0.f);
    if (!pred) return;
    // Opportunistic aggregation over the lanes that arrived together, partitioned by target vertex (match.any); each
    // partition is summed and its first lane issues the atomics.  This is what cooperative_groups' coalesced_threads() +
    // labeled_partition() expand to (an __activemask() followed by __match_any_sync on it, then shuffles on the partition's
    // own mask) written out: cg::reduce / cg shuffles over a non-contiguous partition pay a rank→lane search per shuffle —
    // measured on the room pass (profiles/r02_b_warp_aggregation.md): this form 44.6 ms, cg::reduce 49.8 ms, rank-rotating
    // cg shuffles 159.8 ms.  Every lane named in `peers` executes the same shuffles (the loop runs popc(peers) times for
    // all of them), and nothing between the mask query and the match can split the lanes.
    const unsigned peers = __match_any_sync(__activemask(), idx);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    if (BFLY && peers == 0xffffffffu) {
        // the whole warp adds to one vertex (pixel-major gradient pass: a warp's samples of one pixel hit one triangle): butterfly
        // sum, 15 shuffles instead of the 96 of the general loop below
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { v.x += __shfl_xor_sync(0xffffffffu, v.x, o); v.y += __shfl_xor_sync(0xffffffffu, v.y, o); v.z += __shfl_xor_sync(0xffffffffu, v.z, o); }
    } else if (peers != (1u << lane)) {
        f3 sum = mk3(0, 0, 0);
        for (unsigned rem = peers; rem; rem &= rem - 1) {
            const int src = __ffs(rem) - 1;
            sum.x += __shfl_sync(peers, v.x, src); sum.y += __shfl_sync(peers, v.y, src); sum.z += __shfl_sync(peers, v.z, src);
        }
        v = sum;
    }
    if (lane == leader) {
        float *d = buf + 3 * (size_t) idx;
        atomicAdd(d, v.x); atomicAdd(d + 1, v.y); atomicAdd(d + 2, v.z);          // red.global.add.f32
    }
}

// The Backward-mode replay seeds of sample_path (epsm.py:559-562, 609-627, 644-645)
// `n_avail`: vertices 0 … n_avail−1 of the log are present; the rest are inactive and contribute nothing.
template <int NV>
__device__ void scatter_path(const PathLogT<NV> &P, const PathGradsT<NV> &G, int variant, int n_logged, int n_avail, uint32_t max_depth_prop,
                             float *__restrict__ grad_pos, float *__restrict__ grad_nrm) {
    // len(path_grad) = 5·n_logged (manifold) or 5·(n_logged−1)+3 (caustic: n, m of the last vertex are never added)
    const int n_pg = variant == EPSM_VARIANT_MANIFOLD ? 5 * n_logged : 5 * (n_logged - 1) + 3;
    const bool shadow_term = (max_depth_prop < EPSM_MAX_ITER ? max_depth_prop : EPSM_MAX_ITER) <= 3;
    EPSM_UNROLL_NV
    for (int k = 0; k < NV; ++k) {
        if (k >= n_logged || k >= n_avail) break;
        const VLog &v = P.v[k];
        const bool sel = v.active && (v.mesh_flags & EPSM_MESH_GRAD);
        if (5 * k + 4 < n_pg) {
#pragma unroll
            for (int c = 0; c < 3; ++c) scatter3<(NV < EPSM_MAXV)>(grad_pos, v.fi[c], G.pg[k][c], sel);                   // (i)
            const f3 g3 = G.pg[k][3];                                                                     // (iv)
            const bool any_g3 = sel && (g3.x != 0.f || g3.y != 0.f || g3.z != 0.f);
            const float flip = (v.mesh_flags & EPSM_MESH_FLIP_NORMALS) ? -1.f : 1.f;
            if (any_g3 && v.has_normals) {
                // si_follow.sh_frame.n = normalize(Σ b_k n_k) with uv detached (mesh.cpp:788-791)
                const f3 nt = v.n0 * v.bw[0] + v.n1 * v.bw[1] + v.n2 * v.bw[2];
                const float il = 1.0f / len3(nt); const f3 nh = nt * il;
                const f3 gt = (g3 - nh * dot3(nh, g3)) * (il * flip);
#pragma unroll
                for (int c = 0; c < 3; ++c) scatter3<(NV < EPSM_MAXV)>(grad_nrm, v.fi[c], gt * v.bw[c], true);
            } else if (any_g3 && v.ismesh) {
                // flat-shaded mesh: sh_frame.n = normalize(cross(p1−p0, p2−p0)) → vertex positions (mesh.cpp:729,811)
                const f3 e1 = v.p1 - v.p0, e2 = v.p2 - v.p0, cr = cross3(e1, e2);
                const float il = 1.0f / len3(cr); const f3 nh = cr * il;
                const f3 gc = (g3 - nh * dot3(nh, g3)) * (il * flip);
                const f3 ge1 = cross3(e2, gc), ge2 = cross3(gc, e1);
                scatter3<(NV < EPSM_MAXV)>(grad_pos, v.fi[1], ge1, true); scatter3<(NV < EPSM_MAXV)>(grad_pos, v.fi[2], ge2, true); scatter3<(NV < EPSM_MAXV)>(grad_pos, v.fi[0], -(ge1 + ge2), true);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) scatter3<(NV < EPSM_MAXV)>(grad_pos, v.fi[c], G.dg[k] * v.bw[c], sel);                 // (ii)
        if (k == 0 && shadow_term) {                                                                      // epsm.py:609-620
            const f3 g = G.dg[0] * v.dis;
#pragma unroll
            for (int c = 0; c < 3; ++c) scatter3<(NV < EPSM_MAXV)>(grad_pos, v.dfi[c], g * v.dbw[c], v.direct_any && v.direct_grad);
        }
        const f3 gl = G.lg[k] * v.lr_sum;                                                                 // (iii) epsm.py:622-627
#pragma unroll
        for (int c = 0; c < 3; ++c) scatter3<(NV < EPSM_MAXV)>(grad_pos, v.dfi[c], gl * v.dbw[c], v.direct_valid && v.direct_grad);
    }
}

__device__ void write_record(const PathLog &P, const PathGrads &G, int n_logged, float *__restrict__ r) {
    for (int i = 0; i < EPSM_REC_FLOATS; ++i) r[i] = 0.f;
    for (int k = 0; k < n_logged; ++k) {
        const VLog &v = P.v[k]; float *o = r + k * EPSM_REC_VSTRIDE;
        o[0] = v.active ? 1.f : 0.f; o[1] = u2f(v.prim); o[2] = v.p.x; o[3] = v.p.y; o[4] = v.p.z; o[5] = v.b0; o[6] = v.b1;
        o[7] = v.active_em ? 1.f : 0.f; o[8] = v.light.x; o[9] = v.light.y; o[10] = v.light.z; o[11] = u2f(v.flags); o[12] = v.eta;
        o[13] = v.lr_sum; o[14] = v.dis; o[15] = u2f((v.direct_any ? 1u : 0u) | (v.direct_valid ? 2u : 0u) | (v.direct_grad ? 4u : 0u));
        float *g = r + EPSM_REC_GRAD_OFF + k * EPSM_REC_GSTRIDE;
        for (int c = 0; c < 5; ++c) { g[3 * c] = G.pg[k][c].x; g[3 * c + 1] = G.pg[k][c].y; g[3 * c + 2] = G.pg[k][c].z; }
        g[15] = G.lg[k].x; g[16] = G.lg[k].y; g[17] = G.lg[k].z; g[18] = G.dg[k].x; g[19] = G.dg[k].y; g[20] = G.dg[k].z;
    }
    float *h = r + EPSM_REC_HEAD_OFF;
    h[0] = P.cam.x; h[1] = P.cam.y; h[2] = P.cam.z; h[3] = P.d.x; h[4] = P.d.y; h[5] = P.d.z;
    h[6] = P.dlduv0; h[7] = P.dlduv1; h[8] = P.dldp1.x; h[9] = P.dldp1.y; h[10] = P.dldp1.z; h[11] = P.L.x; h[12] = P.L.y; h[13] = P.L.z;
}

// ------------------------------------------------------------------------------------------------
//  Gradient pass — wavefront
//
//  The replay of sample_path is cut at every scene query (epsm_device.cuh: wf_begin / wf_stage_a / wf_stage_b /
//  wf_direct_hit).  Between the stages a lean persistent traversal kernel drains a compacted ray queue, so the
//  traversal runs with full warps, ~56 registers and a few KB of code instead of sitting inside a divergent
//  > 64 KB megakernel (profiles/r01_a_*).  Per-lane state, the path log and the queues live in an HBM workspace
//  sized for one chunk of the wavefront; chunks are processed back to back.
// ------------------------------------------------------------------------------------------------
struct QRay { float ox, oy, oz, maxt, dx, dy, dz; uint32_t out; };      // 32 B, out = index into the result arena

// Per-lane records live in HBM word-major ("SoA by 32-bit word"): word w of lane i sits at base[w·capacity + i], so
// a warp moving its 32 lanes' records issues fully coalesced 128-byte transactions (the AoS first version wrote one
// 32-byte sector per thread per store: profiles/r01_b_*).
#define WORDS_OF(T) ((sizeof(T) + 3) / 4)
template <typename T> __device__ __forceinline__ void soa_load(T &dst, const uint32_t *__restrict__ base, uint32_t stride, uint32_t i) {
    uint32_t w[WORDS_OF(T)];
#pragma unroll
    for (uint32_t k = 0; k < WORDS_OF(T); ++k) w[k] = base[(size_t) k * stride + i];
    memcpy(&dst, w, sizeof(T));
}
template <typename T> __device__ __forceinline__ void soa_store(const T &src, uint32_t *__restrict__ base, uint32_t stride, uint32_t i) {
    uint32_t w[WORDS_OF(T)];
    w[WORDS_OF(T) - 1] = 0u;
    memcpy(w, &src, sizeof(T));
#pragma unroll
    for (uint32_t k = 0; k < WORDS_OF(T); ++k) base[(size_t) k * stride + i] = w[k];
}

// the first NW words only (a record whose tail is reconstructed by the consumer)
template <uint32_t NW, typename T> __device__ __forceinline__ void soa_load_prefix(T &dst, const uint32_t *__restrict__ base, uint32_t stride, uint32_t i) {
    static_assert(NW * 4 <= sizeof(T), "prefix longer than the record");
    uint32_t w[NW];
#pragma unroll
    for (uint32_t k = 0; k < NW; ++k) w[k] = base[(size_t) k * stride + i];
    memcpy(&dst, w, NW * 4);
}
template <uint32_t NW, typename T> __device__ __forceinline__ void soa_store_prefix(const T &src, uint32_t *__restrict__ base, uint32_t stride, uint32_t i) {
    static_assert(NW * 4 <= sizeof(T), "prefix longer than the record");
    uint32_t w[NW];
    memcpy(w, &src, NW * 4);
#pragma unroll
    for (uint32_t k = 0; k < NW; ++k) base[(size_t) k * stride + i] = w[k];
}

// Lane status word: low byte = one of the codes below, bits 8… = number of vertex records written so far.
// Resident CTAs per SM requested for the stage / finish kernels (register cap = 65536 / (128·N)).  They are chains of
// dependent gathers (status → state → hit → faces → vertices → BSDF), i.e. latency-bound: at the compiler's free choice
// of 255 registers only 8 warps/SM are resident (profiles/r01_d_*).
#ifndef EPSM_STAGE_MINBLOCKS
#define EPSM_STAGE_MINBLOCKS 2
#endif
// Per-kernel choice measured on the B200 (room / glass pass, ms): stage A / B at 2,2 → 54.3 / 27.8; 3,3 → 52.3 / 26.5;
// 4,3 → 81.4 / 26.0 (stage A spills at 128 registers); 4,4 → 58.1 / 41.1.
#ifndef EPSM_STAGEB_MINBLOCKS
#define EPSM_STAGEB_MINBLOCKS 3
#endif
#ifndef EPSM_STAGEA_MINBLOCKS
#define EPSM_STAGEA_MINBLOCKS 3
#endif
#define WF_NONE 0u      // slot without a lane
#define WF_LIVE 1u      // path still being traced
#define WF_DONE 2u      // path has terminated (its remaining logged vertices are inactive)
#define WF_SCATTERED 3u // gradient chain ended at the first vertex and its only seed is already in grad_pos (k_wf_stage_a)

struct WfBuffers {
    uint32_t *status;          // [C]
    uint32_t *states;          // WfState, word-major
    uint32_t *heads;           // PathHead, word-major
    uint32_t *vlogs;           // EPSM_MAXV × VLog, word-major per vertex: base + it·WORDS_OF(VLog)·C
    Hit *hits;                 // [C] path hits, then [MAXV·C] direct hits
    uint32_t *occl;            // [C]
    QRay *q_closest /* 2C */, *q_any /* C */;
    uint32_t *counters;        // [0..15] closest-queue sizes, [16..31] closest heads, [32..47] any sizes, [48..63] any heads,
                               // [64 + it] size of the lane list of iteration it ≥ WF_LIST_FROM
    uint32_t *live0, *live1;   // [C] each: lanes still on their path at iteration it > 2 (ping-pong on it & 1), appended by stage B(it − 1)
    uint32_t *live2;           // [C] the lanes of iteration 2, kept to the end: they are the lanes of the long-chain finish kernel
    uint32_t capacity;
};
#define WF_COUNTER_WORDS 96u
#define WF_LONG_COUNT 80u      // counters[80]: number of lanes selected for the long-chain finish launch
// From this iteration on the stage kernels run over the compacted lane list instead of over all slots: by then a few
// percent of the lanes are left and a full-grid launch is a 0.6 ms floor (262 144 CTAs that read one status word each);
// in the dense early iterations slot order (coalesced records) wins (profiles/r01_f_refit_iteration.md).
#define WF_LIST_FROM 2u
#define WF_LIST_SPARSE 8u   // … and only when fewer than 1/8 of the chunk's slots are still alive

size_t epsm_wf_workspace_bytes(uint32_t capacity) {
    size_t c = capacity;
    return c * (4 * (1 + WORDS_OF(WfState) + WORDS_OF(PathHead) + EPSM_MAXV * WORDS_OF(VLog)) + sizeof(Hit) * (1 + EPSM_MAXV)
                + sizeof(uint32_t) + sizeof(QRay) * 3 + 3 * sizeof(uint32_t)) + WF_COUNTER_WORDS * sizeof(uint32_t) + 20 * 256;
}
static WfBuffers wf_carve(void *base, uint32_t capacity) {
    WfBuffers w; char *p = (char *) base; size_t c = capacity;
    auto take = [&](size_t bytes) { char *r = p; p += (bytes + 255) & ~(size_t) 255; return r; };
    w.counters = (uint32_t *) take(WF_COUNTER_WORDS * sizeof(uint32_t));
    w.status = (uint32_t *) take(c * 4);
    w.states = (uint32_t *) take(c * 4 * WORDS_OF(WfState)); w.heads = (uint32_t *) take(c * 4 * WORDS_OF(PathHead));
    w.vlogs = (uint32_t *) take(c * 4 * EPSM_MAXV * WORDS_OF(VLog));
    w.hits = (Hit *) take(c * sizeof(Hit) * (1 + EPSM_MAXV)); w.occl = (uint32_t *) take(c * sizeof(uint32_t));
    w.q_closest = (QRay *) take(2 * c * sizeof(QRay)); w.q_any = (QRay *) take(c * sizeof(QRay));
    w.live0 = (uint32_t *) take(c * 4); w.live1 = (uint32_t *) take(c * 4); w.live2 = (uint32_t *) take(c * 4);
    w.capacity = capacity; return w;
}

// Work index → lane.  Work items are enumerated tile by tile (8×4 pixels); inside a tile either by sample, then by pixel
// (32 consecutive items = neighbouring pixels at the same sample index) or by pixel, then by sample (DLaunch::pixel_major, below).
// The lane id — which alone fixes the random stream — is pixel·spp + s as in the reference wavefront.
struct WorkItem { uint32_t px, py, s; uint64_t lane; bool valid; };
__device__ __forceinline__ WorkItem wf_work_item(const DLaunch &ln, uint64_t w) {
    const uint32_t W = ln.sensor.w, H = ln.sensor.h, S = ln.s_end - ln.s_begin;
    const uint32_t tiles_x = (W + TILE_W - 1) / TILE_W;
    const uint32_t per_tile = 32u * S;
    uint32_t tile, r;
    if (w < 0x100000000ull) { const uint32_t w32 = (uint32_t) w; tile = w32 / per_tile; r = w32 - tile * per_tile; }   // 32-bit divisions: the common case
    else { tile = (uint32_t) (w / per_tile); r = (uint32_t) (w % per_tile); }
    const uint32_t ty_rel = tile / tiles_x, tx = tile - ty_rel * tiles_x, ty = ln.tile_y0 + ty_rel;
    // Inside a tile, either the tile's 32 pixels at one sample index (a warp = neighbouring pixels: the primal render, whose film
    // kernel splats a warp's samples into 32 different pixels of the shared tile), or consecutive samples of one pixel (the
    // gradient pass: a warp's camera rays are all but identical — one path through the BVH, one hit triangle for the first-hit
    // gathers, one vertex triple for the aggregated scatter; room pass 41.97 → 39.7 ms, profiles/r02_summary.md §R).  Lane ids, and
    // with them every random stream, do not depend on the order.
    WorkItem wi; uint32_t pin;
    if (ln.pixel_major) { pin = r / S; wi.s = ln.s_begin + (r - pin * S); }
    else { wi.s = ln.s_begin + r / 32; pin = r & 31; }
    wi.px = tx * TILE_W + (pin & (TILE_W - 1)); wi.py = ty * TILE_H + pin / TILE_W;
    wi.lane = ((uint64_t) wi.py * W + wi.px) * ln.spp + wi.s;
    wi.valid = wi.px < W && wi.py < H && ty < ln.tile_y1 && wi.lane >= ln.lane_begin && wi.lane < ln.lane_end;
    return wi;
}

// Lane list of iteration `it` ≥ WF_LIST_FROM (selects, not an index: indexing a kernel parameter forces a local copy of it).
__device__ __forceinline__ uint32_t *wf_lane_list(const WfBuffers &wb, uint32_t it) { return it == 2u ? wb.live2 : ((it & 1u) ? wb.live1 : wb.live0); }

// Warp-aggregated queue append; returns the entry index (undefined when !pred).
__device__ __forceinline__ uint32_t queue_push(uint32_t *count, bool pred) {
    if (!pred) return 0;
    const cooperative_groups::coalesced_group g = cooperative_groups::coalesced_threads();      // the lanes that push now
    uint32_t base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(count, g.size());
    return g.shfl(base, 0) + g.thread_rank();
}
__device__ __forceinline__ void queue_write(QRay *q, uint32_t idx, const Ray &r, uint32_t out) {
    float4 *d = reinterpret_cast<float4 *>(q + idx);
    d[0] = make_float4(r.o.x, r.o.y, r.o.z, r.maxt); d[1] = make_float4(r.d.x, r.d.y, r.d.z, __uint_as_float(out));
}

// Persistent traversal kernel.  Every lane owns one ray at a time; the warp schedules itself between three phases so
// that each phase runs with many active lanes (the first wavefront version executed "one step of whatever each lane
// needs" per iteration and ran its leaf code at 3.2/32 and its refill code at 1.9/32 active lanes: profiles/r01_b_*):
//
//   refill   when ≥ WF_REFILL_MIN lanes are idle (or none has a ray): one warp-aggregated atomic claims queue entries,
//            each idle lane loads its 32-byte ray with one 256-bit load;
//   inner    lanes whose `cur` is an inner node: one 128-byte node (3 × LDG.E.256 + 1 × LDG.E.128), 4 slab tests,
//            nearest child kept in a register, the others pushed far → near.  A lane that reaches a leaf parks it in
//            `leaf` and (WF_SPECULATE) keeps descending from its stack until it meets a second leaf;
//   leaf     run when ≥ WF_LEAF_MIN lanes hold a parked leaf or no lane can descend: ≤ 4 Möller–Trumbore tests each
//            (1 × LDG.E.256 + 1 × LDG.E.64 per triangle).
//
// Stack entries are 64-bit (ref, entry distance) words in shared memory, [entry][thread] ⇒ conflict-free, spilling to
// a thread-local array beyond WF_STACK.  The arithmetic is that of bvh_trace (epsm_device.cuh) and the result does
// not depend on the visiting order (closest t, ties → larger primitive index; every subtree whose padded box reaches
// t ≤ best is visited), so hits stay bit-identical to the oracle.
#ifndef WF_STACK
#define WF_STACK 16          // shared-memory stack entries per thread (16 KB per CTA: 8 CTAs/SM leave ≈ 100 KB of L1); A/B in profiles/r02_summary.md §O
#endif
#define WF_LSTACK 80
#define WF_REF_NONE 0x7fffffffu
#ifndef WF_REFILL_MIN
#define WF_REFILL_MIN 8
#endif
#ifndef WF_LEAF_MIN
#define WF_LEAF_MIN 12
#endif
#ifndef WF_SPECULATE
#define WF_SPECULATE 1
#endif
// GEN: the rays are the camera rays of work items [work_base, work_base + n_gen) of `ln` (iteration 0 of the gradient pass and
// of the primal render), generated here exactly as wf_begin / the stage kernels regenerate them — no init kernel, no queue.
template <bool ANY, bool COUNT, bool GEN>
__device__ __forceinline__ void wf_trace_body(const DScene &sc, const QRay *__restrict__ queue, const uint32_t *__restrict__ count, uint32_t *head,
                                              Hit *__restrict__ hits, uint32_t *__restrict__ occl, unsigned long long *counters,
                                              const DLaunch *ln, uint64_t work_base, uint32_t n_gen) {
    __shared__ uint2 sstack[WF_STACK][CTA];
    uint2 lstack[WF_LSTACK];
    const uint32_t n = GEN ? n_gen : *count; const int lane = threadIdx.x & 31;
    uint32_t c_nodes = 0, c_tris = 0, c_rays = 0;
    bool has_ray = false, exhausted = false;
    Ray ray; float ix = 0, iy = 0, iz = 0, ox = 0, oy = 0, oz = 0;
    float best_t = 0, best_u = 0, best_v = 0; uint32_t best_prim = 0xffffffffu, out = 0, cur = WF_REF_NONE, leaf = WF_REF_NONE; bool found = false; int sp = 0;
    ray.o = ray.d = mk3(0, 0, 0); ray.maxt = 0;
    for (;;) {
        // ---- refill ----
        const unsigned idle = __ballot_sync(0xffffffffu, !has_ray);
        if (idle) {
            const int n_idle = __popc(idle);
            if (!exhausted && (n_idle >= WF_REFILL_MIN || idle == 0xffffffffu)) {
                const int leader = __ffs(idle) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(head, (uint32_t) n_idle);
                base = __shfl_sync(0xffffffffu, base, leader);
                if (!has_ray) {
                    const uint32_t i = base + __popc(idle & ((1u << lane) - 1));
                    bool take = i < n;
                    if (GEN) {
                        if (take) {
                            const WorkItem wi = wf_work_item(*ln, work_base + i);
                            take = wi.valid;
                            if (take) { Pcg rng = pcg_for_lane(ln->seed, (uint32_t) wi.lane); CamRay c; camera_ray(ln->sensor, wi.px, wi.py, rng, c); ray = c.ray; out = i; }
                        }
                    } else if (take) {
                        const f8 q = ldg8(reinterpret_cast<const f4 *>(queue + i));
                        ray.o = mk3(q.a.x, q.a.y, q.a.z); ray.d = mk3(q.b.x, q.b.y, q.b.z); ray.maxt = q.a.w; out = f2u(q.b.w);
                    }
                    if (take) {
                        const float dx = fabsf(ray.d.x) < 1e-20f ? mulsign(1e-20f, ray.d.x) : ray.d.x;
                        const float dy = fabsf(ray.d.y) < 1e-20f ? mulsign(1e-20f, ray.d.y) : ray.d.y;
                        const float dz = fabsf(ray.d.z) < 1e-20f ? mulsign(1e-20f, ray.d.z) : ray.d.z;
                        ix = 1.0f / dx; iy = 1.0f / dy; iz = 1.0f / dz; ox = ray.o.x * ix; oy = ray.o.y * iy; oz = ray.o.z * iz;
                        best_t = ray.maxt; best_u = best_v = 0.f; best_prim = 0xffffffffu; found = false; sp = 0;
                        cur = sc.n_nodes ? 0u : WF_REF_NONE; leaf = WF_REF_NONE; has_ray = true;
                        if (COUNT) ++c_rays;
                    }
                }
                if (base + (uint32_t) n_idle >= n) exhausted = true;
            }
            if (__ballot_sync(0xffffffffu, has_ray) == 0u) break;      // queue drained and every lane finished
        }
        // ---- park a reached leaf, pop the next subtree, retire finished rays ----
        if (has_ray) {
            if ((int32_t) cur < 0 && leaf == WF_REF_NONE) { leaf = cur; cur = WF_REF_NONE; }
            while (cur == WF_REF_NONE && (WF_SPECULATE || leaf == WF_REF_NONE)) {
                bool got = false;
                while (sp > 0) {
                    --sp;
                    const uint2 e = sp < WF_STACK ? sstack[sp][threadIdx.x] : lstack[sp - WF_STACK];
                    if (__uint_as_float(e.y) <= best_t) { cur = e.x; got = true; break; }
                }
                if (!got) {
                    if (leaf == WF_REF_NONE) {                           // nothing left: the ray is finished
                        if (ANY) occl[out] = 0u;
                        else reinterpret_cast<uint4 *>(hits)[out] = make_uint4(f2u(found ? best_t : INFINITY), f2u(best_u), f2u(best_v), best_prim);
                        has_ray = false;
                    }
                    break;
                }
                if ((int32_t) cur < 0 && leaf == WF_REF_NONE) { leaf = cur; cur = WF_REF_NONE; }
            }
        }
        // ---- choose the phase ----
        const bool can_inner = has_ray && (int32_t) cur >= 0 && cur != WF_REF_NONE;
        const bool has_leaf = has_ray && leaf != WF_REF_NONE;
        const unsigned m_in = __ballot_sync(0xffffffffu, can_inner), m_leaf = __ballot_sync(0xffffffffu, has_leaf);
        if (m_leaf != 0u && (m_in == 0u || __popc(m_leaf) >= WF_LEAF_MIN)) {
            // ---- leaf phase ----
            if (has_leaf) {
                const uint32_t code = ~leaf, first = code >> EPSM_LEAF_BITS, cnt = (code & (EPSM_MAX_LEAF - 1u)) + 1u;
                bool done = false;
                for (uint32_t k = 0; k < cnt; ++k) {
                    const f4 *tp = sc.tris + (size_t) (first + k) * EPSM_TRI_F4;
                    const f8 ab = ldg8(tp); const f2 c = ldg2(tp + 2); const f4 a = ab.a, b = ab.b;
                    if (COUNT) ++c_tris;
                    float t, u, v;
                    if (tri_test(ray, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(b.w, c.x, c.y), best_t, t, u, v)) {
                        if (ANY) { done = true; break; }
                        const uint32_t prim = f2u(a.w);
                        if (!found || t < best_t || prim > best_prim) { best_t = t; best_u = u; best_v = v; best_prim = prim; found = true; }
                    }
                }
                leaf = WF_REF_NONE;
                if (ANY && done) { occl[out] = 1u; has_ray = false; cur = WF_REF_NONE; sp = 0; }
            }
            continue;
        }
        if (!can_inner) continue;
        // ---- inner phase ----
        const f4 *np = sc.nodes + (size_t) cur * EPSM_NODE_F4;
        const f8 n0 = ldg8(np), n1 = ldg8(np + 2), n2 = ldg8(np + 4);
        const f4 lx = n0.a, ly = n0.b, lz = n1.a, hx = n1.b, hy = n2.a, hz = n2.b, ch = ldg4(np + 6);
        if (COUNT) ++c_nodes;
        uint32_t key[4];
        const uint32_t cref[4] = {f2u(ch.x), f2u(ch.y), f2u(ch.z), f2u(ch.w)};
#define WF_SLAB(k, LX, LY, LZ, HX, HY, HZ)                                                          \
        {                                                                                           \
            const float ax = fmaf(LX, ix, -ox), bx = fmaf(HX, ix, -ox);                             \
            const float ay = fmaf(LY, iy, -oy), by = fmaf(HY, iy, -oy);                             \
            const float az = fmaf(LZ, iz, -oz), bz = fmaf(HZ, iz, -oz);                             \
            const float t0 = fmaxf(fmax3(fminf(ax, bx), fminf(ay, by), fminf(az, bz)), 0.0f);        \
            const float t1 = fminf(fmin3(fmaxf(ax, bx), fmaxf(ay, by), fmaxf(az, bz)), best_t);      \
            key[k] = (t0 <= t1 * 1.0000004f && cref[k] != 0x7fffffffu) ? ((f2u(t0) & ~3u) | (uint32_t) k) : 0xffffffffu; \
        }
        WF_SLAB(0, lx.x, ly.x, lz.x, hx.x, hy.x, hz.x)
        WF_SLAB(1, lx.y, ly.y, lz.y, hx.y, hy.y, hz.y)
        WF_SLAB(2, lx.z, ly.z, lz.z, hx.z, hy.z, hz.z)
        WF_SLAB(3, lx.w, ly.w, lz.w, hx.w, hy.w, hz.w)
#undef WF_SLAB
        cswap(key[0], key[1]); cswap(key[2], key[3]); cswap(key[0], key[2]); cswap(key[1], key[3]); cswap(key[1], key[2]);
#pragma unroll
        for (int i = 3; i >= 1; --i) {
            if (key[i] != 0xffffffffu) {
                const uint32_t k = key[i] & 3u;
                const uint32_t r = k == 0 ? cref[0] : k == 1 ? cref[1] : k == 2 ? cref[2] : cref[3];
                // the key carries the entry distance (two mantissa bits cleared: never larger, so the pop test stays conservative)
                const float t = u2f(key[i] & ~3u);
                const uint2 e = make_uint2(r, __float_as_uint(t));
                if (sp < WF_STACK) sstack[sp][threadIdx.x] = e; else if (sp < WF_STACK + WF_LSTACK) lstack[sp - WF_STACK] = e;
                if (sp < WF_STACK + WF_LSTACK) ++sp;
            }
        }
        if (key[0] != 0xffffffffu) {
            const uint32_t k = key[0] & 3u;
            cur = k == 0 ? cref[0] : k == 1 ? cref[1] : k == 2 ? cref[2] : cref[3];
        } else cur = WF_REF_NONE;
    }
    if (COUNT) {
        unsigned long long a = c_nodes, b = c_tris, c = c_rays;
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); c += __shfl_down_sync(0xffffffffu, c, o); }
        if (lane == 0) { atomicAdd(counters, a); atomicAdd(counters + 1, b); atomicAdd(counters + 2, c); }
    }
}
template <bool ANY, bool COUNT>
__global__ void __launch_bounds__(CTA) k_wf_trace(const __grid_constant__ DScene sc, const QRay *__restrict__ queue, const uint32_t *__restrict__ count, uint32_t *head,
                                                  Hit *__restrict__ hits, uint32_t *__restrict__ occl, unsigned long long *counters) {
    wf_trace_body<ANY, COUNT, false>(sc, queue, count, head, hits, occl, counters, nullptr, 0, 0);
}
// Closest hits of the camera rays of one chunk: hits[i] for work item work_base + i (slots that hold no lane are skipped).
template <bool COUNT>
__global__ void __launch_bounds__(CTA, 8) k_wf_trace_primary(const __grid_constant__ DScene sc, const __grid_constant__ DLaunch ln, uint64_t work_base, uint32_t n, uint32_t *head,
                                                             Hit *__restrict__ hits, unsigned long long *counters) {
    wf_trace_body<false, COUNT, true>(sc, nullptr, nullptr, head, hits, nullptr, counters, &ln, work_base, n);
}

// `grad_pos` != NULL enables the first-vertex fast path (scatter passes; the record probe keeps every lane).
// STRIDED (iterations ≥ WF_LIST_FROM): launched with a small fixed grid that strides over the compacted lane list when few
// lanes are left (the list size lives on the device; its order is scrambled by the concurrent appends, so it pays only when
// sparse) and over the slots otherwise.  The dense early iterations use the loop-free instantiation (one slot per thread).
template <bool STRIDED>
__global__ void __launch_bounds__(128, EPSM_STAGEA_MINBLOCKS) k_wf_stage_a(const __grid_constant__ DScene sc, const __grid_constant__ DLaunch ln, uint32_t it, uint32_t n, WfBuffers wb, float *__restrict__ grad_pos,
                                                                            uint64_t work_base, const float *__restrict__ grad_in) {
    const bool listed = STRIDED && wb.counters[64 + it] < n / WF_LIST_SPARSE;
    const uint32_t cnt = listed ? wb.counters[64 + it] : n;
    uint32_t base = blockIdx.x * 128;
    if (base >= cnt) return;
    do {
    const uint32_t j = base + threadIdx.x;
    const uint32_t i = listed ? (j < cnt ? wf_lane_list(wb, it)[j] : 0u) : j;
    bool push = false; Ray sr;
    if (!STRIDED && it == 0) {
        // Iteration 0 starts the lane: the camera ray, its random stream and the pixel-motion tangent are regenerated from
        // the work item (the traversal kernel did the same for the ray), so no init kernel and no state hand-over exist.
        if (j < cnt) {
            const WorkItem wi = wf_work_item(ln, work_base + i);
            uint32_t status = WF_NONE;
            if (wi.valid) {
                // grad_in is film-sized; samples of the border (sample_border sensors) carry no pixel-motion gradient
                const int fx = (int) wi.px - ln.sensor.border, fy = (int) wi.py - ln.sensor.border;
                const bool in_film = fx >= 0 && fy >= 0 && fx < (int) ln.sensor.fw && fy < (int) ln.sensor.fh;
                const uint64_t pixel = in_film ? (uint64_t) fy * ln.sensor.fw + fx : 0;
                WfState S; PathHead H;
                wf_begin(ln.sensor, wi.px, wi.py, ln.seed, (uint32_t) wi.lane, in_film ? grad_in[pixel * 5 + 3] : 0.f, in_film ? grad_in[pixel * 5 + 4] : 0.f, S, H);
                SI si;
                push = wf_stage_a(sc, S, H, wb.hits[i], it, ln.max_depth, ln.variant, ln.prune != 0, si);
                if (grad_pos && wf_first_vertex_final(S, it, ln.prune != 0)) {
                    // the path ends here for the gradient: seed (ii) of its first vertex is its only output
                    f3 dg; const bool sel = wf_first_vertex_seed(sc, S, H, si, dg);
#pragma unroll
                    for (int c = 0; c < 3; ++c) scatter3(grad_pos, si.fi[c], dg * si.bw[c], sel);
                    status = WF_SCATTERED; push = false;
                } else {
                    soa_store(H, wb.heads, wb.capacity, i);
                    sr = S.shadow_ray; soa_store(S, wb.states, wb.capacity, i);
                    status = WF_LIVE;
                }
            }
            wb.status[i] = status;
        }
    } else if (j < cnt && (listed || (wb.status[i] & 0xffu) == WF_LIVE)) {
        WfState S; const Hit h = wb.hits[i];
        soa_load(S, wb.states, wb.capacity, i);
        PathHead H;                                   // only written at the first hit
        push = wf_stage_a(sc, S, H, h, it, ln.max_depth, ln.variant, ln.prune != 0);
        sr = S.shadow_ray; soa_store(S, wb.states, wb.capacity, i);
    }
    const uint32_t q = queue_push(wb.counters + 32 + it, push);
    if (push) queue_write(wb.q_any, q, sr, i);
    base += gridDim.x * 128;
    } while (STRIDED && base < cnt);
}

template <bool STRIDED>
__global__ void __launch_bounds__(128, EPSM_STAGEB_MINBLOCKS) k_wf_stage_b(const __grid_constant__ DScene sc, const __grid_constant__ DLaunch ln, uint32_t it, uint32_t max_iter, uint32_t n, WfBuffers wb, bool keep_state) {
    const bool listed = STRIDED && wb.counters[64 + it] < n / WF_LIST_SPARSE;
    const uint32_t cnt = listed ? wb.counters[64 + it] : n;
    uint32_t base = blockIdx.x * 128;
    if (base >= cnt) return;
    do {
    const uint32_t j = base + threadIdx.x;
    const uint32_t i = listed ? (j < cnt ? wf_lane_list(wb, it)[j] : 0u) : j;
    bool push_direct = false, push_next = false; Ray direct, next;
    if (j < cnt && (listed || (wb.status[i] & 0xffu) == WF_LIVE)) {
        WfState S; soa_load(S, wb.states, wb.capacity, i);
        const Hit h = wb.hits[i];
        const bool occluded = S.need_shadow && wb.occl[i] != 0u;
        VLog v;
        wf_stage_b(sc, S, v, h, occluded, it, ln.rr_depth, ln.prune != 0, direct, push_direct);
        if (it < EPSM_MAXV) soa_store_prefix<VLOG_STAGE_WORDS>(v, wb.vlogs + (size_t) it * WORDS_OF(VLog) * wb.capacity, wb.capacity, i);   // the tail is the finish kernel's
        push_next = S.active && it + 1 < max_iter; next = S.ray;
        if (S.active || keep_state) soa_store(S, wb.states, wb.capacity, i);       // a finished lane's state is read by the record probe only (P.L)
        const uint32_t n_written = it + 1 < EPSM_MAXV ? it + 1 : EPSM_MAXV;
        wb.status[i] = (S.active ? WF_LIVE : WF_DONE) | (n_written << 8);
    }
    const uint32_t qd = queue_push(wb.counters + (it + 1), push_direct);
    if (push_direct) queue_write(wb.q_closest, qd, direct, wb.capacity + it * wb.capacity + i);
    const uint32_t qn = queue_push(wb.counters + (it + 1), push_next);
    if (push_next) queue_write(wb.q_closest, qn, next, i);
    if (it + 1 >= WF_LIST_FROM) {                           // the lanes of the next iteration, for its stage kernels
        const uint32_t l = queue_push(wb.counters + 64 + (it + 1), push_next);
        if (push_next) wf_lane_list(wb, it + 1)[l] = i;
    }
    base += gridDim.x * 128;
    } while (STRIDED && base < cnt);
}

__global__ void k_count_lanes(uint32_t n, WfBuffers wb, unsigned long long *counters) {
    const uint32_t i = blockIdx.x * 128 + threadIdx.x;
    const bool live = i < n && (wb.status[i] & 0xffu) != WF_NONE;
    const unsigned m = __ballot_sync(0xffffffffu, live);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(counters + 3, (unsigned long long) __popc(m));
}

// Minimal "inactive vertex" for the slot just behind the loaded prefix: calc_grad reads it as the next vertex of
// section B (active / flags / triangle) and as the loop terminator.
__device__ __forceinline__ void vlog_inactive(VLog &v) {
    v.active = v.active_em = v.direct_valid = v.direct_any = v.direct_grad = v.direct_traced = false;
    v.flags = v.ismesh = v.mesh_flags = v.has_normals = 0; v.prim = 0xffffffffu;
    v.p0 = v.p1 = v.p2 = mk3(0, 0, 0); v.b0 = v.b1 = 0.f;
}

// NV = vertex slots held per lane.  The scatter pass runs twice: NV = WF_SHORT_NV takes the lanes whose replay logged at
// most WF_SHORT_NV − 1 vertices (the gradient chain of most paths ends at their second vertex; with static slot indices
// the solve stays in registers), NV = EPSM_MAXV takes the rest (`skip_short`).  The record probe uses NV = EPSM_MAXV alone.
#define WF_SHORT_NV 3
#ifndef EPSM_FINISH_SHORT_BLOCKS
#define EPSM_FINISH_SHORT_BLOCKS 4
#endif
#ifndef EPSM_FINISH_LONG_BLOCKS
#define EPSM_FINISH_LONG_BLOCKS 2
#endif
template <bool RECORD, int NV>
__device__ __forceinline__ void finish_lane(const DScene &sc, const DLaunch &ln, uint64_t work_base, uint32_t i, const WfBuffers &wb,
                                            float *__restrict__ grad_pos, float *__restrict__ grad_nrm, float *__restrict__ records, bool skip_short) {
    const uint32_t status = wb.status[i], code = status & 0xffu;
    if (code == WF_NONE || code == WF_SCATTERED) return;
    const int n_written = (int) (status >> 8);
    if (NV < EPSM_MAXV ? n_written > NV - 1 : (skip_short && n_written <= WF_SHORT_NV - 1)) return;
    PathLogT<NV> P;
    soa_load(static_cast<PathHead &>(P), wb.heads, wb.capacity, i);
    const uint32_t max_iter = wf_iterations(ln.max_depth);
    const int n_logged = (int) (max_iter < EPSM_MAXV ? max_iter : EPSM_MAXV);      // len(path_info) − 1 (independent of pruning)
    // Only the prefix of the path that can carry gradients is fetched: the chain ends at the first diffuse vertex
    // (manifold, epsm.py:855,920) or at the second one (caustic: vertex 1 must be diffuse, epsm.py:999,1008), at a
    // non-mesh vertex or where the path left the scene; every output beyond it is masked to zero by calc_grad.
    bool chain = true; int n_diffuse = 0, n_avail = 0;
    EPSM_UNROLL_NV
    for (int k = 0; k < NV; ++k) {
        if (k < n_logged && k < n_written && (chain || RECORD)) {
            soa_load_prefix<VLOG_STAGE_WORDS>(P.v[k], wb.vlogs + (size_t) k * WORDS_OF(VLog) * wb.capacity, wb.capacity, i);
            VLog &v = P.v[k];
            v.dfi[0] = v.dfi[1] = v.dfi[2] = 0u; v.dbw[0] = v.dbw[1] = v.dbw[2] = 0.f; v.dis = 0.f; v.hf = mk3(0, 0, 0);
            if (v.active && v.direct_traced) wf_direct_hit(sc, v, wb.hits[wb.capacity + (size_t) k * wb.capacity + i]);
            const bool diffuse = (v.flags & EPSM_BF_DIFFUSE) != 0;
            n_diffuse += diffuse ? 1 : 0;
            if (!v.active || !v.ismesh) chain = false;
            if (ln.variant == EPSM_VARIANT_MANIFOLD) { if (diffuse) chain = false; }
            else { if ((k == 0 && !diffuse) || n_diffuse >= 2) chain = false; }
            n_avail = k + 1;
        } else if (RECORD) {
            // a lane that has left the path: the logged vertex is inactive
            memset(&P.v[k], 0, sizeof(VLog)); P.v[k].prim = 0xffffffffu;
        } else {
            vlog_inactive(P.v[k]);                        // the gradient-dead tail starts here; nothing behind it is read
            if (NV == EPSM_MAXV) break;                   // (the short instantiation fills its few slots: static indices)
        }
    }
    if (RECORD) { WfState S; soa_load(S, wb.states, wb.capacity, i); P.L = S.L; n_avail = NV; }
    PathGradsT<NV> G;
    calc_grad<NV>(P, ln.variant, n_logged, G, n_avail);
    if (RECORD) { if (NV == EPSM_MAXV) { const WorkItem wi = wf_work_item(ln, work_base + i); write_record(reinterpret_cast<const PathLog &>(P), reinterpret_cast<const PathGrads &>(G), n_logged, records + (wi.lane - ln.lane_begin) * EPSM_REC_FLOATS); } }
    else scatter_path<NV>(P, G, ln.variant, n_logged, n_avail, ln.max_depth, grad_pos, grad_nrm);
}
// LISTED (the long-chain scatter launch when those lanes are few): one thread per entry of `list`, the slots whose lane
// logged more than two vertices, selected in slot order by a device compaction (neighbouring pixels → similar control
// flow and neighbouring records).
struct WfLongLane {
    const uint32_t *status;
    __device__ __forceinline__ bool operator()(uint32_t i) const {
        const uint32_t st = status[i], code = st & 0xffu;
        return code != WF_NONE && code != WF_SCATTERED && (st >> 8) > (uint32_t) (WF_SHORT_NV - 1);
    }
};
template <bool RECORD, int NV, bool LISTED>
__global__ void __launch_bounds__(128, NV < EPSM_MAXV ? EPSM_FINISH_SHORT_BLOCKS : EPSM_FINISH_LONG_BLOCKS) k_wf_finish(const __grid_constant__ DScene sc, const __grid_constant__ DLaunch ln, uint64_t work_base, uint32_t n, WfBuffers wb,
                                                   float *__restrict__ grad_pos, float *__restrict__ grad_nrm, float *__restrict__ records, bool skip_short,
                                                   const uint32_t *__restrict__ list) {
    const uint32_t j = blockIdx.x * 128 + threadIdx.x;
    if (j >= n) return;                                       // LISTED: n = entries of the list
    finish_lane<RECORD, NV>(sc, ln, work_base, LISTED ? list[j] : j, wb, grad_pos, grad_nrm, records, skip_short);
}

// ------------------------------------------------------------------------------------------------
//  Primal render — wavefront (EPSMIntegrator.render / sample, epsm.py:13-82, 308-501)
//
//  [closest: camera rays generated in k_wf_trace_primary] → k_pr_stage(0) → [any ∥ closest] → k_pr_stage(1) → … → k_pr_film.  One stage per bounce
//  (pr_stage, epsm_device.cuh); both queries of a bounce are drained by the persistent traversal kernel of the gradient
//  pass.  The film kernel owns one 8×4-pixel tile per CTA — work items are enumerated tile by tile, so a tile's samples
//  are contiguous — accumulates the 5×5 Gaussian footprints in shared memory and flushes each film pixel once.
// ------------------------------------------------------------------------------------------------
#ifndef EPSM_PR_MINBLOCKS
#define EPSM_PR_MINBLOCKS 5    // measured (room primal, ms): 4 CTAs/SM (113 registers) 188.4, 5 → 186.5, 6 → 186.7, 8 → 186.9
#endif
__global__ void __launch_bounds__(128, EPSM_PR_MINBLOCKS) k_pr_stage(const __grid_constant__ DScene sc, const __grid_constant__ DLaunch ln, uint32_t it, uint32_t n, WfBuffers wb, uint64_t work_base) {
    const uint32_t i = blockIdx.x * 128 + threadIdx.x;
    bool push_shadow = false, push_next = false, live = false; Ray shadow; PrState S;
    if (i < n) {
        if (it == 0) {
            // bounce 0 starts the lane from its work item (the traversal kernel generated the same camera ray)
            const WorkItem wi = wf_work_item(ln, work_base + i);
            live = wi.valid;
            if (live) { S.rng = pcg_for_lane(ln.seed, (uint32_t) wi.lane); CamRay c; camera_ray(ln.sensor, wi.px, wi.py, S.rng, c); S.ray = c.ray; pr_state_defaults(S); }
            else wb.status[i] = WF_NONE;
        } else {
            live = wb.status[i] == WF_LIVE;
            if (live) { soa_load(S, wb.states, wb.capacity, i); pr_resolve_pending(S, wb.occl[i] != 0u); }
        }
    }
    if (live) {
        push_shadow = pr_stage(sc, S, wb.hits[i], ln.max_depth, ln.rr_depth, shadow);
        push_next = (S.bits & PR_ACTIVE) != 0 && it + 1 < ln.max_depth;
        soa_store(S, wb.states, wb.capacity, i);
        if (it == 0) wb.status[i] = push_next ? WF_LIVE : WF_DONE;
        else if (!push_next) wb.status[i] = WF_DONE;
    }
    const uint32_t slot = (it + 1) & 15u;
    const uint32_t qa = queue_push(wb.counters + 32 + slot, push_shadow);
    if (push_shadow) queue_write(wb.q_any, qa, shadow, i);
    const uint32_t qn = queue_push(wb.counters + slot, push_next);
    if (push_next) queue_write(wb.q_closest, qn, S.ray, i);
}

// ImageBlock::put, coalesced branch (imageblock.cpp:414-494) into a tile-local RGBW block with a 2-pixel apron, then
// HDRFilm::put_block (hdrfilm.cpp:286-320): pixels outside the film are the cropped border.
__global__ void __launch_bounds__(128) k_pr_film(const __grid_constant__ DScene sc, const __grid_constant__ DLaunch ln, uint64_t work_base, uint32_t n, WfBuffers wb, float *__restrict__ film,
                                                 float *__restrict__ lane_L) {
    __shared__ float ftile[FT_H][FT_W][4];
    for (int k = threadIdx.x; k < FT_H * FT_W * 4; k += 128) (&ftile[0][0][0])[k] = 0.f;
    __syncthreads();
    const uint32_t W = ln.sensor.w, S = ln.s_end - ln.s_begin;          // sampling grid (film + border)
    const int B = ln.sensor.border, FW = (int) ln.sensor.fw, FH = (int) ln.sensor.fh;
    const uint32_t tiles_x = (W + TILE_W - 1) / TILE_W;
    const uint64_t per_tile = 32ull * S;
    const uint64_t tile = work_base / per_tile + blockIdx.x;
    const uint64_t w0 = tile * per_tile > work_base ? tile * per_tile : work_base;
    const uint64_t w1 = (tile + 1) * per_tile < work_base + n ? (tile + 1) * per_tile : work_base + n;
    const uint32_t tx = (uint32_t) (tile % tiles_x), ty = ln.tile_y0 + (uint32_t) (tile / tiles_x);
    for (uint64_t w = w0 + threadIdx.x; w < w1; w += 128) {
        const uint32_t i = (uint32_t) (w - work_base);
        if (wb.status[i] == WF_NONE) continue;
        const WorkItem wi = wf_work_item(ln, w);
        // radiance: words of PrState::L (+ the NEE term still waiting for its shadow ray)
        PrState St; soa_load(St, wb.states, wb.capacity, i);
        pr_resolve_pending(St, wb.occl[i] != 0u);
        const f3 L = St.L;
        if (lane_L) { float *o = lane_L + 3 * (size_t) wi.lane; o[0] = L.x; o[1] = L.y; o[2] = L.z; }    // per-lane radiance for prb_reparam's reverse pass
        Pcg rng = pcg_for_lane(ln.seed, (uint32_t) wi.lane);
        const float jx = pcg_f32(rng), jy = pcg_f32(rng);
        const float pos_x = (float) ((int) wi.px - B) + jx, pos_y = (float) ((int) wi.py - B) + jy;     // camera_ray's film position
        const int pix = (int) floorf(pos_x) - 2, piy = (int) floorf(pos_y) - 2;
        float rel_x = (float) pix + .5f - pos_x, rel_y = (float) piy + .5f - pos_y;
        float wx[5], wy[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) { wx[k] = filter_eval(sc.filter, rel_x); wy[k] = filter_eval(sc.filter, rel_y); rel_x += 1; rel_y += 1; }
        const int lx = pix - ((int) (tx * TILE_W) - B) + 2, ly = piy - ((int) (ty * TILE_H) - B) + 2;
#pragma unroll
        for (int ys = 0; ys < 5; ++ys)
#pragma unroll
            for (int xs = 0; xs < 5; ++xs) {
                const float wgt = wy[ys] * wx[xs];
                float *d = ftile[ly + ys][lx + xs];
                atomicAdd(d + 0, L.x * wgt); atomicAdd(d + 1, L.y * wgt); atomicAdd(d + 2, L.z * wgt); atomicAdd(d + 3, 1.0f * wgt);
            }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < FT_H * FT_W; k += 128) {
        const int fy = (int) (ty * TILE_H) - B + k / FT_W - 2, fx = (int) (tx * TILE_W) - B + k % FT_W - 2;      // film coordinates
        if (fx < 0 || fy < 0 || fx >= FW || fy >= FH) continue;
        const float *sv = ftile[k / FT_W][k % FT_W];
        if (sv[3] == 0.f && sv[0] == 0.f && sv[1] == 0.f && sv[2] == 0.f) continue;
        float *d = film + ((size_t) fy * FW + fx) * 4;
        atomicAdd(d + 0, sv[0]); atomicAdd(d + 1, sv[1]); atomicAdd(d + 2, sv[2]); atomicAdd(d + 3, sv[3]);
    }
}

// ------------------------------------------------------------------------------------------------
//  Parity probes
// ------------------------------------------------------------------------------------------------
__global__ void k_sampler_stream(uint32_t seed, uint64_t lane_begin, uint64_t n_lanes, uint32_t n_floats, float *__restrict__ out) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lanes) return;
    Pcg r = pcg_for_lane(seed, (uint32_t) (lane_begin + i));
    for (uint32_t k = 0; k < n_floats; ++k) out[i * n_floats + k] = pcg_f32(r);
}
__global__ void k_sample_rays(DSensor s, uint32_t seed, uint32_t spp, uint64_t lane_begin, uint64_t n_lanes, float *__restrict__ out) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lanes) return;
    uint64_t lane = lane_begin + i; uint32_t pixel = (uint32_t) (lane / spp);
    Pcg r = pcg_for_lane(seed, (uint32_t) lane);
    CamRay c; camera_ray(s, pixel % s.w, pixel / s.w, r, c);
    float *o = out + 14 * i;
    o[0] = c.ray.o.x; o[1] = c.ray.o.y; o[2] = c.ray.o.z; o[3] = c.ray.d.x; o[4] = c.ray.d.y; o[5] = c.ray.d.z;
    o[6] = c.d_x.x; o[7] = c.d_x.y; o[8] = c.d_x.z; o[9] = c.d_y.x; o[10] = c.d_y.y; o[11] = c.d_y.z; o[12] = c.pos_x; o[13] = c.pos_y;
}

// ------------------------------------------------------------------------------------------------
//  Mesh::recompute_vertex_normals (mesh.cpp:296-330, JIT branch): angle-weighted face normals scattered to
//  vertices, then normalised.
// ------------------------------------------------------------------------------------------------
__global__ void k_normals_accumulate(const float *__restrict__ pos, const uint32_t *__restrict__ faces, uint32_t f_off, uint32_t f_cnt,
                                     float *__restrict__ nrm) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f_cnt) return;
    const uint32_t *f = faces + 3 * (size_t) (f_off + i);
    const uint32_t fi[3] = {f[0], f[1], f[2]};
    const f3 v[3] = {ld3(pos, fi[0]), ld3(pos, fi[1]), ld3(pos, fi[2])};
    const f3 n = norm3(cross3(v[1] - v[0], v[2] - v[0]));
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const f3 d0 = norm3(v[(k + 1) % 3] - v[k]), d1 = norm3(v[(k + 2) % 3] - v[k]);
        const float c = fminf(fmaxf(dot3(d0, d1), -1.f), 1.f);
        const f3 nn = n * acosf(c);
        if (nn.x == nn.x && nn.y == nn.y && nn.z == nn.z) {
            float *d = nrm + 3 * (size_t) fi[k];
            atomicAdd(d, nn.x); atomicAdd(d + 1, nn.y); atomicAdd(d + 2, nn.z);
        }
    }
}
__global__ void k_normals_finish(float *__restrict__ nrm, uint32_t v_off, uint32_t v_cnt) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v_cnt) return;
    float *d = nrm + 3 * (size_t) (v_off + i);
    f3 n = mk3(d[0], d[1], d[2]); float l = len3(n);
    n = l != 0.f ? n * (1.0f / l) : mk3(1, 0, 0);
    d[0] = n.x; d[1] = n.y; d[2] = n.z;
}
// Adjoint of the two kernels above (the reference differentiates recompute_vertex_normals through Dr.Jit's AD when
// params.update() re-derives the normals from AD-attached positions, mesh.cpp:82-87,311-330, so the sh_frame.n·g3 seed
// of epsm.py:644-645 reaches vertex_positions).  `acc` holds the un-normalised sums Σ n·angle (k_normals_accumulate);
// k_normals_bwd_vertex turns it in place into dL/d(sum) = (g − n̂ (n̂·g)) / |sum|; k_normals_bwd_face pushes that through
// n = normalize(cross(e1, e2)) and angle_k = safe_acos(dot(normalize(a), normalize(b))) into the positions.
__global__ void k_normals_bwd_vertex(float *__restrict__ acc, const float *__restrict__ grad_nrm, uint32_t v_off, uint32_t v_cnt) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v_cnt) return;
    float *d = acc + 3 * (size_t) (v_off + i);
    const f3 s = mk3(d[0], d[1], d[2]), g = ld3(grad_nrm, v_off + i);
    const float l = len3(s);
    f3 o = mk3(0, 0, 0);
    if (l != 0.f) { const f3 n = s * (1.0f / l); o = (g - n * dot3(n, g)) * (1.0f / l); }      // the bogus (1,0,0) normal is a constant
    d[0] = o.x; d[1] = o.y; d[2] = o.z;
}
__device__ __forceinline__ void normalize_bwd(const f3 &v, const f3 &g_unit, f3 &g_v) {      // u = v/|v|: g_v = (g_u − u (u·g_u)) / |v|
    const float l = len3(v);
    if (!(l > 0.f)) { g_v = mk3(0, 0, 0); return; }
    const f3 u = v * (1.0f / l);
    g_v = (g_unit - u * dot3(u, g_unit)) * (1.0f / l);
}
__global__ void k_normals_bwd_face(const float *__restrict__ pos, const uint32_t *__restrict__ faces, uint32_t f_off, uint32_t f_cnt,
                                   const float *__restrict__ g_sum, float *__restrict__ grad_pos) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f_cnt) return;
    const uint32_t *f = faces + 3 * (size_t) (f_off + i);
    const uint32_t fi[3] = {f[0], f[1], f[2]};
    const f3 v[3] = {ld3(pos, fi[0]), ld3(pos, fi[1]), ld3(pos, fi[2])};
    const f3 e1 = v[1] - v[0], e2 = v[2] - v[0], c = cross3(e1, e2), n = norm3(c);
    f3 gv[3] = {mk3(0, 0, 0), mk3(0, 0, 0), mk3(0, 0, 0)};
    f3 g_n = mk3(0, 0, 0);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
        const f3 a = v[k1] - v[k], b = v[k2] - v[k], d0 = norm3(a), d1 = norm3(b);
        const float x = dot3(d0, d1), xc = fminf(fmaxf(x, -1.f), 1.f), ang = acosf(xc);
        const f3 nn = n * ang;
        if (!(nn.x == nn.x && nn.y == nn.y && nn.z == nn.z)) continue;               // the forward skipped this corner
        const f3 gs = ld3(g_sum, fi[k]);
        g_n = g_n + gs * ang;
        const float g_ang = dot3(n, gs);
        if (x > -1.f && x < 1.f) {                                                     // safe_acos: no gradient in the clamped range
            const float g_x = -g_ang / sqrtf(1.f - x * x);
            f3 g_a, g_b; normalize_bwd(a, d1 * g_x, g_a); normalize_bwd(b, d0 * g_x, g_b);
            gv[k1] = gv[k1] + g_a; gv[k2] = gv[k2] + g_b; gv[k] = gv[k] - g_a - g_b;
        }
    }
    f3 g_c; normalize_bwd(c, g_n, g_c);
    const f3 g_e1 = cross3(e2, g_c), g_e2 = cross3(g_c, e1);
    gv[1] = gv[1] + g_e1; gv[2] = gv[2] + g_e2; gv[0] = gv[0] - g_e1 - g_e2;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (!(gv[k].x == gv[k].x && gv[k].y == gv[k].y && gv[k].z == gv[k].z)) continue;
        float *d = grad_pos + 3 * (size_t) fi[k];
        atomicAdd(d, gv[k].x); atomicAdd(d + 1, gv[k].y); atomicAdd(d + 2, gv[k].z);
    }
}
__global__ void k_zero_range(float *__restrict__ p, uint64_t off, uint64_t n) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[off + i] = 0.f;
}

// ------------------------------------------------------------------------------------------------
//  BVH4 refit on the device (per-iteration scene update; reference: Scene::parameters_changed rebuilds the
//  Embree / OptiX structure, scene.cpp:304-336, scene_embree.inl:130-187).  Two launches:
//    k_refit_tris   one thread per triangle slot: p0 | e1 | e2 from the new positions, max |coordinate|
//    k_refit_nodes  one thread per node whose children are all leaves; a finished node stores its box into
//                   its parent's slot and bumps the parent's arrival counter, the last arrival climbs on
//  scratch: [0] max|coordinate| bits, [2..3] Σ half-area (double), [4 + i] arrival counter of node i.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_refit_tris(f4 *__restrict__ tris, uint32_t n_slots, const uint32_t *__restrict__ faces,
                                                    const float *__restrict__ pos, uint32_t *__restrict__ scratch) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.f;
    if (i < n_slots) m = refit_tri_slot(tris, faces, pos, i);
    const uint32_t wm = __reduce_max_sync(0xffffffffu, __float_as_uint(m));        // non-negative floats order like their bits
    if ((threadIdx.x & 31u) == 0 && wm != 0) atomicMax(scratch, wm);
}
__global__ void __launch_bounds__(256) k_refit_nodes(f4 *nodes, uint32_t n_nodes, const f4 *__restrict__ tris,
                                                     const uint32_t *__restrict__ faces, const float *__restrict__ pos, uint32_t *__restrict__ scratch) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    double area = 0.0;
    if (i < n_nodes && __float_as_uint(nodes[(size_t) i * EPSM_NODE_F4 + 7].y) == 0u) {
        const float pad = refit_pad(__uint_as_float(*(volatile uint32_t *) scratch));
        uint32_t *arrive = scratch + 4;
        uint32_t cur = i;
        for (;;) {
            float lo[3], hi[3];
            area += (double) refit_node(nodes, cur, tris, faces, pos, pad, lo, hi);
            const uint32_t link = __float_as_uint(nodes[(size_t) cur * EPSM_NODE_F4 + 7].x);
            if (link == 0xffffffffu) break;                                          // root
            refit_store_parent(nodes, link, lo, hi);
            __threadfence();
            const uint32_t parent = link >> 2, need = __float_as_uint(nodes[(size_t) parent * EPSM_NODE_F4 + 7].y);
            if (atomicAdd(arrive + parent, 1u) + 1u != need) break;                  // a sibling subtree is still open
            __threadfence();
            cur = parent;
        }
    }
    for (int o = 16; o > 0; o >>= 1) area += __shfl_xor_sync(0xffffffffu, area, o);
    if ((threadIdx.x & 31u) == 0 && area != 0.0) atomicAdd(reinterpret_cast<double *>(scratch + 2), area);
}

// ------------------------------------------------------------------------------------------------
//  Launchers
// ------------------------------------------------------------------------------------------------
static inline void tile_rows(DLaunch &ln) {
    const uint64_t row_lanes = (uint64_t) ln.sensor.w * ln.spp;
    uint64_t y0 = ln.lane_begin / row_lanes, y1 = (ln.lane_end + row_lanes - 1) / row_lanes;
    if (y1 > ln.sensor.h) y1 = ln.sensor.h;
    ln.tile_y0 = (uint32_t) (y0 / TILE_H); ln.tile_y1 = (uint32_t) ((y1 + TILE_H - 1) / TILE_H);
}
static inline uint32_t tile_grid(const DLaunch &ln) {
    const uint32_t tiles_x = (ln.sensor.w + TILE_W - 1) / TILE_W;
    return tiles_x * (ln.tile_y1 > ln.tile_y0 ? ln.tile_y1 - ln.tile_y0 : 0);
}

cudaError_t epsm_launch_trace(const DScene &sc, const float *rays, uint64_t n, void *hits, int any_hit, unsigned long long *counters, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const uint32_t grid = (uint32_t) ((n + CTA - 1) / CTA);
    const f4 *r = reinterpret_cast<const f4 *>(rays); uint32_t *h = reinterpret_cast<uint32_t *>(hits);
    if (any_hit) { if (counters) k_trace<true, true><<<grid, CTA, 0, s>>>(sc, r, n, h, counters); else k_trace<true, false><<<grid, CTA, 0, s>>>(sc, r, n, h, nullptr); }
    else { if (counters) k_trace<false, true><<<grid, CTA, 0, s>>>(sc, r, n, h, counters); else k_trace<false, false><<<grid, CTA, 0, s>>>(sc, r, n, h, nullptr); }
    return cudaGetLastError();
}
cudaError_t epsm_launch_render(const DScene &sc, DLaunch ln, float *film, unsigned long long *counters, cudaStream_t s) {
    ln.pixel_major = 0u; tile_rows(ln); const uint32_t grid = tile_grid(ln);
    if (grid == 0) return cudaSuccess;
    if (counters) k_render<true><<<grid, CTA, 0, s>>>(sc, ln, film, counters); else k_render<false><<<grid, CTA, 0, s>>>(sc, ln, film, nullptr);
    return cudaGetLastError();
}
// Wavefront primal render.  `host_count` is a pinned (or pageable) word used to poll the queue size once the path
// length exceeds 8 bounces (max_depth = −1 means "until every path has died", common.py:33-37).
cudaError_t epsm_launch_render_wf(const DScene &sc, DLaunch ln, float *film, unsigned long long *counters, void *workspace, uint32_t capacity,
                                  uint32_t trace_grid, uint64_t *n_launches, cudaStream_t s, float *lane_L) {
    ln.pixel_major = 0u;
    tile_rows(ln);
    const uint32_t tiles = tile_grid(ln);
    if (tiles == 0 || ln.s_end <= ln.s_begin) return cudaSuccess;
    const uint64_t per_tile = 32ull * (ln.s_end - ln.s_begin);
    const uint64_t total = (uint64_t) tiles * per_tile;
    WfBuffers wb = wf_carve(workspace, capacity);
    uint64_t launches = 0;
    for (uint64_t base = 0; base < total; base += capacity) {
        const uint32_t n = (uint32_t) ((total - base) < capacity ? (total - base) : capacity);
        const uint32_t grid = (n + 127) / 128;
        cudaMemsetAsync(wb.counters, 0, 64 * sizeof(uint32_t), s);
        auto trace = [&](bool any, uint32_t slot) {
            if (any) {
                if (counters) k_wf_trace<true, true><<<trace_grid, CTA, 0, s>>>(sc, wb.q_any, wb.counters + 32 + slot, wb.counters + 48 + slot, nullptr, wb.occl, counters);
                else k_wf_trace<true, false><<<trace_grid, CTA, 0, s>>>(sc, wb.q_any, wb.counters + 32 + slot, wb.counters + 48 + slot, nullptr, wb.occl, nullptr);
            } else {
                if (counters) k_wf_trace<false, true><<<trace_grid, CTA, 0, s>>>(sc, wb.q_closest, wb.counters + slot, wb.counters + 16 + slot, wb.hits, nullptr, counters);
                else k_wf_trace<false, false><<<trace_grid, CTA, 0, s>>>(sc, wb.q_closest, wb.counters + slot, wb.counters + 16 + slot, wb.hits, nullptr, nullptr);
            }
            ++launches;
        };
        uint32_t last_slot = 0; bool shadow_left = true;
        for (uint32_t it = 0; it < ln.max_depth; ++it) {
            const uint32_t slot = it & 15u, next = (it + 1) & 15u;
            if (it == 0) {                                // camera rays: generated inside the traversal kernel
                if (counters) k_wf_trace_primary<true><<<trace_grid, CTA, 0, s>>>(sc, ln, base, n, wb.counters + 16, wb.hits, counters);
                else k_wf_trace_primary<false><<<trace_grid, CTA, 0, s>>>(sc, ln, base, n, wb.counters + 16, wb.hits, nullptr);
                ++launches;
            } else trace(false, slot);
            if (it > 0) trace(true, slot);
            if (it + 1 >= 16) {                           // the counter slots are reused cyclically
                cudaMemsetAsync(wb.counters + next, 0, 4, s); cudaMemsetAsync(wb.counters + 16 + next, 0, 4, s);
                cudaMemsetAsync(wb.counters + 32 + next, 0, 4, s); cudaMemsetAsync(wb.counters + 48 + next, 0, 4, s);
            }
            k_pr_stage<<<grid, 128, 0, s>>>(sc, ln, it, n, wb, base); ++launches;
            last_slot = next;
            if (it + 1 >= 8 && it + 1 < ln.max_depth) {   // long paths: poll the queue sizes, stop when no path ray is left
                uint32_t left[2] = {0, 0};
                cudaMemcpyAsync(&left[0], wb.counters + next, 4, cudaMemcpyDeviceToHost, s);
                cudaMemcpyAsync(&left[1], wb.counters + 32 + next, 4, cudaMemcpyDeviceToHost, s);
                cudaError_t e = cudaStreamSynchronize(s); if (e != cudaSuccess) return e;
                if (left[0] == 0) { shadow_left = left[1] != 0; break; }
            }
        }
        if (shadow_left && ln.max_depth > 0) trace(true, last_slot);      // shadow rays pushed by the last stage
        const uint32_t n_tiles = (uint32_t) ((base + n - 1) / per_tile - base / per_tile + 1);
        k_pr_film<<<n_tiles, 128, 0, s>>>(sc, ln, base, n, wb, film, lane_L); ++launches;
        if (counters) { k_count_lanes<<<grid, 128, 0, s>>>(n, wb, counters); ++launches; }
    }
    if (n_launches) *n_launches += launches;
    return cudaGetLastError();
}
cudaError_t epsm_launch_develop(const float *film, uint64_t n_pixels, float *out, cudaStream_t s) {
    if (n_pixels == 0) return cudaSuccess;
    k_film_develop<<<(uint32_t) ((n_pixels + 255) / 256), 256, 0, s>>>(film, n_pixels, out);
    return cudaGetLastError();
}
cudaError_t epsm_launch_backward(const DScene &sc, DLaunch ln, const float *grad_in, float *grad_pos, float *grad_nrm, float *records,
                                 unsigned long long *counters, void *workspace, uint32_t capacity, uint32_t trace_grid, uint64_t *n_launches,
                                 EpsmEventList *trace_events, cudaStream_t s) {
    auto ev_begin = [&]() { if (trace_events) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, s); trace_events->push_back({a, b}); } };
    auto ev_end = [&]() { if (trace_events) cudaEventRecord(trace_events->back().second, s); };
    ln.pixel_major = 1u;
    tile_rows(ln);
    const uint32_t tiles = tile_grid(ln);
    if (tiles == 0 || ln.s_end <= ln.s_begin) return cudaSuccess;
    const uint64_t total = (uint64_t) tiles * 32ull * (ln.s_end - ln.s_begin);
    const uint32_t max_iter = wf_replay_iterations(ln.max_depth, ln.prune != 0);
    WfBuffers wb = wf_carve(workspace, capacity);
    uint64_t launches = 0;
    for (uint64_t base = 0; base < total; base += capacity) {
        const uint32_t n = (uint32_t) ((total - base) < capacity ? (total - base) : capacity);
        const uint32_t grid = (n + 127) / 128;
        cudaMemsetAsync(wb.counters, 0, WF_COUNTER_WORDS * sizeof(uint32_t), s);
        for (uint32_t it = 0; it <= max_iter; ++it) {
            // closest-hit queue `it`: path rays of iteration `it` and the direct rays of iteration `it − 1`
            ev_begin();
            if (it == 0) {                                // camera rays: generated inside the traversal kernel
                if (counters) k_wf_trace_primary<true><<<trace_grid, CTA, 0, s>>>(sc, ln, base, n, wb.counters + 16, wb.hits, counters);
                else k_wf_trace_primary<false><<<trace_grid, CTA, 0, s>>>(sc, ln, base, n, wb.counters + 16, wb.hits, nullptr);
            } else if (counters) k_wf_trace<false, true><<<trace_grid, CTA, 0, s>>>(sc, wb.q_closest, wb.counters + it, wb.counters + 16 + it, wb.hits, nullptr, counters);
            else k_wf_trace<false, false><<<trace_grid, CTA, 0, s>>>(sc, wb.q_closest, wb.counters + it, wb.counters + 16 + it, wb.hits, nullptr, nullptr);
            ev_end();
            ++launches;
            if (it == max_iter) break;
            const bool strided = it >= WF_LIST_FROM && grid > trace_grid * 4;
            const uint32_t sgrid = strided ? trace_grid * 4 : grid;
            if (strided) k_wf_stage_a<true><<<sgrid, 128, 0, s>>>(sc, ln, it, n, wb, records ? nullptr : grad_pos, base, grad_in);
            else k_wf_stage_a<false><<<sgrid, 128, 0, s>>>(sc, ln, it, n, wb, records ? nullptr : grad_pos, base, grad_in);
            ++launches;
            ev_begin();
            if (counters) k_wf_trace<true, true><<<trace_grid, CTA, 0, s>>>(sc, wb.q_any, wb.counters + 32 + it, wb.counters + 48 + it, nullptr, wb.occl, counters);
            else k_wf_trace<true, false><<<trace_grid, CTA, 0, s>>>(sc, wb.q_any, wb.counters + 32 + it, wb.counters + 48 + it, nullptr, wb.occl, nullptr);
            ev_end();
            ++launches;
            if (strided) k_wf_stage_b<true><<<sgrid, 128, 0, s>>>(sc, ln, it, max_iter, n, wb, records != nullptr);
            else k_wf_stage_b<false><<<sgrid, 128, 0, s>>>(sc, ln, it, max_iter, n, wb, records != nullptr);
            ++launches;
        }
        if (records) { k_wf_finish<true, EPSM_MAXV, false><<<grid, 128, 0, s>>>(sc, ln, base, n, wb, nullptr, nullptr, records, false, nullptr); ++launches; }
        else {
            k_wf_finish<false, WF_SHORT_NV, false><<<grid, 128, 0, s>>>(sc, ln, base, n, wb, grad_pos, grad_nrm, nullptr, false, nullptr); ++launches;
            if (max_iter > 2) {                               // lanes with more than two logged vertices exist only then
                // Ordered compaction on the device (the lane lists are free by now: live0 receives the selection, live1 is the
                // library's scratch), then one 4-byte read-back per chunk to size the launch.  Measured alternatives
                // (profiles/r02_summary.md §L): a fixed grid striding over the device-side count needs no synchronisation
                // but runs this 255-register kernel 65 % slower (6.8 against 4.1 ms); the radix sort of the appended lane
                // list that this replaces cost five library launches for the same order.
                size_t temp = 0; uint32_t n_long = 0; bool listed = false;
                thrust::counting_iterator<uint32_t> slots(0u);
                cub::DeviceSelect::If(nullptr, temp, slots, wb.live0, wb.counters + WF_LONG_COUNT, (int) n, WfLongLane{wb.status}, s);
                if (temp <= (size_t) capacity * sizeof(uint32_t) &&
                    cub::DeviceSelect::If(wb.live1, temp, slots, wb.live0, wb.counters + WF_LONG_COUNT, (int) n, WfLongLane{wb.status}, s) == cudaSuccess) {   // (the library's kernels are not counted as ours)
                    cudaMemcpyAsync(&n_long, wb.counters + WF_LONG_COUNT, sizeof(uint32_t), cudaMemcpyDeviceToHost, s);
                    cudaError_t e = cudaStreamSynchronize(s); if (e != cudaSuccess) return e;
                    listed = true;
                }
                if (listed && n_long != 0) k_wf_finish<false, EPSM_MAXV, true><<<(n_long + 127) / 128, 128, 0, s>>>(sc, ln, base, n_long, wb, grad_pos, grad_nrm, nullptr, true, wb.live0);
                else if (!listed) k_wf_finish<false, EPSM_MAXV, false><<<grid, 128, 0, s>>>(sc, ln, base, n, wb, grad_pos, grad_nrm, nullptr, true, nullptr);
                if (!listed || n_long != 0) ++launches;
            }
        }
        if (counters) { k_count_lanes<<<grid, 128, 0, s>>>(n, wb, counters); ++launches; }
    }
    if (n_launches) *n_launches += launches;
    return cudaGetLastError();
}
cudaError_t epsm_launch_sampler_stream(uint32_t seed, uint64_t lane_begin, uint64_t n_lanes, uint32_t n_floats, float *out, cudaStream_t s) {
    if (n_lanes == 0) return cudaSuccess;
    k_sampler_stream<<<(uint32_t) ((n_lanes + 127) / 128), 128, 0, s>>>(seed, lane_begin, n_lanes, n_floats, out);
    return cudaGetLastError();
}
cudaError_t epsm_launch_sample_rays(const DSensor &se, uint32_t seed, uint32_t spp, uint64_t lane_begin, uint64_t n_lanes, float *out, cudaStream_t s) {
    if (n_lanes == 0) return cudaSuccess;
    k_sample_rays<<<(uint32_t) ((n_lanes + 127) / 128), 128, 0, s>>>(se, seed, spp, lane_begin, n_lanes, out);
    return cudaGetLastError();
}
cudaError_t epsm_launch_vertex_normals(const float *pos, const uint32_t *faces, uint32_t f_off, uint32_t f_cnt, uint32_t v_off, uint32_t v_cnt,
                                       float *nrm, cudaStream_t s) {
    if (v_cnt == 0 || f_cnt == 0) return cudaSuccess;
    k_zero_range<<<(3 * v_cnt + 255) / 256, 256, 0, s>>>(nrm, 3 * (uint64_t) v_off, 3 * (uint64_t) v_cnt);
    k_normals_accumulate<<<(f_cnt + 255) / 256, 256, 0, s>>>(pos, faces, f_off, f_cnt, nrm);
    k_normals_finish<<<(v_cnt + 255) / 256, 256, 0, s>>>(nrm, v_off, v_cnt);
    return cudaGetLastError();
}

cudaError_t epsm_launch_vertex_normals_backward(const float *pos, const uint32_t *faces, uint32_t f_off, uint32_t f_cnt, uint32_t v_off, uint32_t v_cnt,
                                                const float *grad_nrm, float *grad_pos, float *scratch, cudaStream_t s) {
    if (v_cnt == 0 || f_cnt == 0) return cudaSuccess;
    k_zero_range<<<(3 * v_cnt + 255) / 256, 256, 0, s>>>(scratch, 3 * (uint64_t) v_off, 3 * (uint64_t) v_cnt);
    k_normals_accumulate<<<(f_cnt + 255) / 256, 256, 0, s>>>(pos, faces, f_off, f_cnt, scratch);
    k_normals_bwd_vertex<<<(v_cnt + 255) / 256, 256, 0, s>>>(scratch, grad_nrm, v_off, v_cnt);
    k_normals_bwd_face<<<(f_cnt + 255) / 256, 256, 0, s>>>(pos, faces, f_off, f_cnt, scratch, grad_pos);
    return cudaGetLastError();
}

size_t epsm_refit_scratch_words(uint32_t n_nodes) { return 4 + (size_t) n_nodes; }
cudaError_t epsm_launch_refit(f4 *nodes, uint32_t n_nodes, f4 *tris, uint32_t n_tris, const uint32_t *faces, const float *pos,
                              uint32_t *scratch, cudaStream_t s) {
    if (n_nodes == 0 || n_tris == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(scratch, 0, epsm_refit_scratch_words(n_nodes) * sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
    k_refit_tris<<<(n_tris + 255) / 256, 256, 0, s>>>(tris, n_tris, faces, pos, scratch);
    k_refit_nodes<<<(n_nodes + 255) / 256, 256, 0, s>>>(nodes, n_nodes, tris, faces, pos, scratch);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
//  prb_reparam reverse pass (epsm_prb.cuh): one thread per lane.  Lanes are enumerated pixel-major over the launch's
//  sample range, so that a warp holds neighbouring samples; the lane id (pixel·spp + s) fixes both random streams.
// ------------------------------------------------------------------------------------------------
struct AtomicSink {
    float *gp, *gn;
    __device__ __forceinline__ void pos(uint32_t v, f3 g) { atomicAdd(gp + 3 * v, g.x); atomicAdd(gp + 3 * v + 1, g.y); atomicAdd(gp + 3 * v + 2, g.z); }
    __device__ __forceinline__ void nrm(uint32_t v, f3 g) { atomicAdd(gn + 3 * v, g.x); atomicAdd(gn + 3 * v + 1, g.y); atomicAdd(gn + 3 * v + 2, g.z); }
};
#ifndef EPSM_PRB_MINBLOCKS
#define EPSM_PRB_MINBLOCKS 2
#endif
__global__ void __launch_bounds__(CTA, EPSM_PRB_MINBLOCKS) k_prb_backward(const __grid_constant__ DScene sc, const __grid_constant__ DLaunch ln, const __grid_constant__ PrbConfig cfg,
                                                      const float *__restrict__ film, const float *__restrict__ grad_in, float *grad_pos, float *grad_nrm) {
    __shared__ SharedStack ss;
    TStack st; bind_stack(st, ss);
    Tracer tr(sc, st);
    const uint32_t S = ln.s_end - ln.s_begin;
    const uint64_t i = (uint64_t) blockIdx.x * CTA + threadIdx.x, n_pix = (uint64_t) ln.sensor.w * ln.sensor.h;
    if (i >= n_pix * S) return;
    const uint64_t pixel = i / S; const uint32_t s = ln.s_begin + (uint32_t) (i - pixel * S);
    const uint64_t lane = pixel * ln.spp + s;
    if (lane < ln.lane_begin || lane >= ln.lane_end) return;
    const uint32_t py = (uint32_t) (pixel / ln.sensor.w), px = (uint32_t) (pixel - (uint64_t) py * ln.sensor.w);
    AtomicSink sink{grad_pos, grad_nrm};
    MegaIo<false> io(tr);
    prb_backward_lane(sc, io, ln.sensor, cfg, px, py, ln.seed, (uint32_t) lane, film, grad_in, sink);
}
cudaError_t epsm_launch_prb_backward(const DScene &sc, DLaunch ln, const PrbConfig &cfg, const float *film, const float *grad_in,
                                     float *grad_pos, float *grad_nrm, cudaStream_t s) {
    const uint64_t total = (uint64_t) ln.sensor.w * ln.sensor.h * (ln.s_end - ln.s_begin);
    if (total == 0) return cudaSuccess;
    k_prb_backward<<<(unsigned) ((total + CTA - 1) / CTA), CTA, 0, s>>>(sc, ln, cfg, film, grad_in, grad_pos, grad_nrm);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
//  prb_reparam reverse pass as a replay wavefront.  A lane's rays come in rounds (epsm_prb.cuh, MegaIo comment): pass p of
//  k_prb_pass replays the lane from its seed — the detached walk is cheap next to its rays — reads the results of rounds
//  < p from per-lane slots, queues the rays of round p for the persistent traversal kernel (k_wf_trace) and stops; the
//  adjoint of iteration `it` runs in the one pass (2·it + 3) in which its auxiliary hits are fresh.  No path state is
//  stored between passes: only hits.  2·K + 2 passes for K iterations, each followed by one closest-hit and one any-hit
//  drain of the queues; lanes that have finished drop out through their status word.
// ------------------------------------------------------------------------------------------------
struct WfPrb {
    uint32_t *status;       // [C] 0 live, 1 finished / not a lane
    Hit *hits;              // [(K + 1 + 2n)·C] slot-major: path hits 0 … K, then the auxiliary hits
    uint32_t *occl;         // [2K·C]
    QRay *q_closest, *q_any;
    uint32_t *counters;     // [0] closest count, [1] closest head, [2] any count, [3] any head
    const float *lane_L;    // per-lane radiance of the primal pass, indexed by global lane id
    uint32_t *ckpt;         // PrbCheckpoint, word-major: the state at the top of the lane's latest iteration
    uint32_t C, K, n_aux;
};
struct WfIo {
    static constexpr bool wavefront = true;
    const WfPrb &wb; int pass; uint32_t i; uint64_t lane;
    __device__ __forceinline__ void push(QRay *q, uint32_t *count, const Ray &r, uint32_t out) {
        const uint32_t k = queue_push(count, true); queue_write(q, k, r, out);
    }
    __device__ __forceinline__ Hit closest(int round, uint32_t slot, const Ray &r) {
        if (round == pass) { push(wb.q_closest, wb.counters + 0, r, slot * wb.C + i); return Hit{0.f, 0.f, 0.f, 0xffffffffu}; }
        return wb.hits[(size_t) slot * wb.C + i];
    }
    __device__ __forceinline__ bool occluded(int round, uint32_t slot, const Ray &r) {
        if (round == pass) { push(wb.q_any, wb.counters + 2, r, slot * wb.C + i); return false; }
        return wb.occl[(size_t) slot * wb.C + i] != 0u;
    }
    __device__ __forceinline__ bool stop_after(int round) const { return pass == round; }
    __device__ __forceinline__ bool adjoint_now(int round) const { return pass == round + 1; }
    __device__ __forceinline__ bool aux_push_only(int round) const { return pass == round; }
    __device__ __forceinline__ bool aux_skip(int round) const { return pass > round + 1; }
    // n consecutive entries of the closest-hit queue for this lane (one atomic per warp for all the lanes that reserve together)
    __device__ __forceinline__ uint32_t aux_reserve(uint32_t n) {
        const cooperative_groups::coalesced_group g = cooperative_groups::coalesced_threads();
        uint32_t base = 0;
        if (g.thread_rank() == 0) base = atomicAdd(wb.counters + 0, n * g.size());
        return g.shfl(base, 0) + n * g.thread_rank();
    }
    __device__ __forceinline__ void aux_push(uint32_t j, uint32_t q, const Ray &r) { queue_write(wb.q_closest, q, r, (wb.K + 1 + j) * wb.C + i); }
    __device__ __forceinline__ Hit aux_hit(uint32_t j, const Ray &) { return wb.hits[(size_t) (wb.K + 1 + j) * wb.C + i]; }
    __device__ __forceinline__ bool lane_radiance(f3 &L) const { const float *p = wb.lane_L + 3 * lane; L = mk3(p[0], p[1], p[2]); return true; }
    __device__ __forceinline__ void finished() { wb.status[i] = 1u; }
    // passes 0 and 1 start at the camera; every later pass starts at the top of the iteration stored by pass 2·it + 1
    __device__ __forceinline__ bool resume(PrbCheckpoint &c) const { if (pass < 2) return false; soa_load(c, wb.ckpt, wb.C, i); return true; }
    __device__ __forceinline__ void checkpoint(uint32_t it, const PrbCheckpoint &c) const { if (pass == 2 * (int) it + 1) soa_store(c, wb.ckpt, wb.C, i); }
};
EPSM_HD f3 path_primal_io(const DScene &, WfIo &, LocalTrace &, Ray, Pcg &, uint32_t, uint32_t, bool &) { return mk3(0, 0, 0); }   // never reached: lane_radiance() is always available

#ifndef EPSM_PRBPASS_MINBLOCKS
#define EPSM_PRBPASS_MINBLOCKS 2
#endif
__global__ void __launch_bounds__(128, EPSM_PRBPASS_MINBLOCKS) k_prb_pass(const __grid_constant__ DScene sc, const __grid_constant__ DLaunch ln, const __grid_constant__ PrbConfig cfg, int pass,
                                                    uint64_t work_base, uint32_t n, const __grid_constant__ WfPrb wb, const float *__restrict__ film,
                                                    const float *__restrict__ grad_in, float *grad_pos, float *grad_nrm) {
    const uint32_t i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    if (pass > 0 && wb.status[i]) return;
    const uint32_t S = ln.s_end - ln.s_begin;
    const uint64_t g = work_base + i, pixel = g / S; const uint32_t smp = ln.s_begin + (uint32_t) (g - pixel * S);
    const uint64_t lane = pixel * ln.spp + smp;
    if (pass == 0) {
        const bool ok = lane >= ln.lane_begin && lane < ln.lane_end;
        wb.status[i] = ok ? 0u : 1u;
        if (!ok) return;
    }
    const uint32_t py = (uint32_t) (pixel / ln.sensor.w), px = (uint32_t) (pixel - (uint64_t) py * ln.sensor.w);
    WfIo io{wb, pass, i, lane};
    AtomicSink sink{grad_pos, grad_nrm};
    prb_backward_lane(sc, io, ln.sensor, cfg, px, py, ln.seed, (uint32_t) lane, film, grad_in, sink);
}

size_t epsm_prb_workspace_bytes(uint32_t capacity, uint32_t K, uint32_t n_aux) {
    const size_t c = capacity;
    return 1024 + c * 4 + c * 4 * WORDS_OF(PrbCheckpoint) + 256 + c * sizeof(Hit) * (K + 1 + 2 * (size_t) n_aux) + c * 4 * 2 * (size_t) K + c * sizeof(QRay) * (2 * (size_t) n_aux + 1) + c * sizeof(QRay) + 16 * 256;
}
cudaError_t epsm_launch_prb_backward_wf(const DScene &sc, DLaunch ln, const PrbConfig &cfg, const float *film, const float *lane_L, const float *grad_in,
                                        float *grad_pos, float *grad_nrm, void *workspace, uint32_t capacity, uint32_t trace_grid, uint64_t *n_launches, cudaStream_t s) {
    const uint64_t total = (uint64_t) ln.sensor.w * ln.sensor.h * (ln.s_end - ln.s_begin);
    if (total == 0) return cudaSuccess;
    const uint32_t K = cfg.max_depth, n_aux = cfg.n_rays;
    WfPrb wb; char *p = (char *) workspace; const size_t c = capacity;
    auto take = [&](size_t bytes) { char *r = p; p += (bytes + 255) & ~(size_t) 255; return r; };
    wb.counters = (uint32_t *) take(1024);
    wb.status = (uint32_t *) take(c * 4);
    wb.ckpt = (uint32_t *) take(c * 4 * WORDS_OF(PrbCheckpoint));
    wb.hits = (Hit *) take(c * sizeof(Hit) * (K + 1 + 2 * (size_t) n_aux));
    wb.occl = (uint32_t *) take(c * 4 * 2 * (size_t) K);
    wb.q_closest = (QRay *) take(c * sizeof(QRay) * (2 * (size_t) n_aux + 1));
    wb.q_any = (QRay *) take(c * sizeof(QRay));
    wb.lane_L = lane_L; wb.C = capacity; wb.K = K; wb.n_aux = n_aux;
    uint64_t launches = 0;
    for (uint64_t base = 0; base < total; base += capacity) {
        const uint32_t n = (uint32_t) ((total - base) < capacity ? (total - base) : capacity);
        for (int pass = 0; pass <= 2 * (int) K + 1; ++pass) {
            cudaMemsetAsync(wb.counters, 0, 4 * sizeof(uint32_t), s);
            k_prb_pass<<<(n + 127) / 128, 128, 0, s>>>(sc, ln, cfg, pass, base, n, wb, film, grad_in, grad_pos, grad_nrm); ++launches;
            if (pass == 2 * (int) K + 1) break;                                     // the last pass only runs the last adjoint
            k_wf_trace<false, false><<<trace_grid, CTA, 0, s>>>(sc, wb.q_closest, wb.counters + 0, wb.counters + 1, wb.hits, nullptr, nullptr); ++launches;
            k_wf_trace<true, false><<<trace_grid, CTA, 0, s>>>(sc, wb.q_any, wb.counters + 2, wb.counters + 3, nullptr, wb.occl, nullptr); ++launches;
        }
    }
    if (n_launches) *n_launches += launches;
    return cudaGetLastError();
}
