// epsm_sinkhorn.cu — the matcher that turns (rendered image, target image) into grad_in for the manifold integrators:
// Matcher.match_Sinkhorn (EPSM/utils/matcher.py:51-63) = geomloss SamplesLoss("sinkhorn", blur, scaling), p = 2,
// debias, balanced, uniform weights, differentiated w.r.t. the rendered 5-D cloud.
//
// geomloss is absent from the reference tree (README.md:11, unpinned); the algorithm restated here is geomloss 0.2.x
// `sinkhorn_loop` / `sinkhorn_cost` / `epsilon_schedule` (oracle/sinkhorn_ref.py holds the literal restatement):
//
//   ε-schedule  [d², exp(arange(2 ln d, 2 ln blur, 2 ln scaling)) …, blur²],  d = |bbox(x ∪ y)|
//   init        f_ba = softmin(ε₀, C_xy, ln b)   g_ab = softmin(ε₀, C_yx, ln a)   f_aa, g_bb likewise on (x,x), (y,y)
//   per ε       four softmins from the OLD potentials, then averaged with them (symmetric Sinkhorn)
//   last        one more un-averaged step at blur²; loss = ⟨a, f_ba − f_aa⟩ + ⟨b, g_ab − g_bb⟩
//   gradient    ∂loss/∂x_i = a_i (Σ_j w^xx_ij x_j − Σ_j w^xy_ij y_j),  w = softmax_j(h_j − C_ij/ε)   (potentials detached)
//
// with softmin(ε, C, h)_i = −ε ln Σ_j exp(h_j − C_ij/ε), C_ij = |x_i − y_j|²/2.  The reference materialises N×N matrices
// (tensorized backend) or streams them through KeOps; here one kernel evaluates all four softmins of a step as an online
// log-sum-exp over column tiles staged in shared memory — nothing of size N² ever exists.  FP32 FMA + SFU ex2: K = 5 is
// far too thin for tensor cores, and C/ε at ε = 1e-4 needs the full fp32 mantissa: the inner loop uses the dot-product form
// only after recentring both clouds on the CTA's rows (term2), which keeps the cancelling terms small; the gradient launch
// forms distances from coordinate differences.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <math.h>
#include <stdint.h>
#include "epsm_sinkhorn.h"

namespace {

constexpr int SK_THREADS = 256;          // 8 warps: each takes a column slice of every tile
constexpr int SK_ROWS = 64;              // rows per CTA: 2 per lane
constexpr int SK_TILE = 1024;            // columns per shared-memory tile (8 floats each = 32 KB)
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Packed FP32x2 arithmetic (sm_100a FFMA2 / FADD2 / FMUL2: two fp32 lanes per issue slot).
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

struct RowAcc { float m, s, a[5]; };

// Packed columns: `x` holds (x_d, x_d), the tile holds (y_d[j], y_d[j+1]); log2 of the two terms comes back packed.
struct RowAcc2 { float m; u64 s; };
// log2 of two terms in the recentred dot-product form: with c the centre of the CTA's rows, x̂ = (x − c)·log2e/ε (per row),
// ỹ = y − c and h' = h·log2e − ½·log2e/ε·|ỹ|² (per column, formed while the tile is staged),
//     (h − |x − y|²/2ε)·log2e = h' + x̂·ỹ − ½·log2e/ε·|x − c|²,
// and the last term is a per-row constant that leaves the sum and is applied once at the end.  Five packed FMAs per pair of
// columns instead of eleven operations.  Precision: the products are of size |x − c|·|y − c|/ε — for the pairs that matter
// (|x − y|² ≲ 60 ε) and rows of one 8×8-pixel block that is ≤ a few units, i.e. an error of ≤ 1e-3 in the exponent at ε = blur²,
// the same as the rounding of h' itself (|h| ~ 1/ε); the un-centred |x|² + |y|² − 2x·y would lose three more digits.
__device__ __forceinline__ u64 term2(const u64 *xh, const ulonglong2 t0, const ulonglong2 t1, const ulonglong2 t2) {
    return fma2(xh[4], t2.x, fma2(xh[3], t1.y, fma2(xh[2], t1.x, fma2(xh[1], t0.y, fma2(xh[0], t0.x, t2.y)))));
}
// Four columns (two packed pairs) into one row's running log-sum-exp: one maximum test per four terms.
__device__ __forceinline__ void fold4(RowAcc2 &r, u64 ta, u64 tb) {
    float a0, a1, b0, b1; upk(ta, a0, a1); upk(tb, b0, b1);
    const float tm = fmaxf(fmaxf(a0, a1), fmaxf(b0, b1));
    if (tm > r.m) { const float c = ex2(r.m - tm); r.s = mul2(r.s, pk(c, c)); r.m = tm; }     // rare once the maximum has settled
    const u64 nm = pk(-r.m, -r.m);
    upk(add2(ta, nm), a0, a1); upk(add2(tb, nm), b0, b1);
    r.s = add2(r.s, add2(pk(ex2(a0), ex2(a1)), pk(ex2(b0), ex2(b1))));
}

template <bool GRAD>
__device__ __forceinline__ void visit(RowAcc &r, const float *x, const float4 y0, const float4 y1, float halfk) {
    const float d0 = x[0] - y0.x, d1 = x[1] - y0.y, d2 = x[2] - y0.z, d3 = x[3] - y0.w, d4 = x[4] - y1.x;
    const float q = fmaf(d4, d4, fmaf(d3, d3, fmaf(d2, d2, fmaf(d1, d1, d0 * d0))));
    const float t = fmaf(-halfk, q, y1.y);                 // log2 of the term: (h_j − C_ij/ε)·log2 e
    if (t > r.m) {                                          // rare once the running maximum has settled
        const float c = ex2(r.m - t);
        r.s *= c;
        if (GRAD) { r.a[0] *= c; r.a[1] *= c; r.a[2] *= c; r.a[3] *= c; r.a[4] *= c; }
        r.m = t;
    }
    const float e = ex2(t - r.m);
    r.s += e;
    if (GRAD) {
        r.a[0] = fmaf(e, y0.x, r.a[0]); r.a[1] = fmaf(e, y0.y, r.a[1]); r.a[2] = fmaf(e, y0.z, r.a[2]);
        r.a[3] = fmaf(e, y0.w, r.a[3]); r.a[4] = fmaf(e, y1.x, r.a[4]);
    }
}

// One Sinkhorn step: blockIdx.y selects the softmin (job), blockIdx.x a block of SK_ROWS rows.
template <bool GRAD>
__global__ void __launch_bounds__(SK_THREADS, GRAD ? 2 : 3) k_sink_softmin(SinkStep st) {
    const SinkJob job = st.job[blockIdx.y];
    const uint32_t row0 = blockIdx.x * SK_ROWS;
    if (row0 >= job.n_rows) return;
    __shared__ float4 tile[2 * SK_TILE];
    __shared__ float red[SK_THREADS / 32][SK_ROWS][GRAD ? 7 : 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float inv_eps = 1.0f / st.eps, halfk = 0.5f * LOG2E * inv_eps;
    float x[2][5]; RowAcc acc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        uint32_t i = row0 + 32 * r + lane; if (i >= job.n_rows) i = job.n_rows - 1;
#pragma unroll
        for (int d = 0; d < 5; ++d) x[r][d] = job.rows[5 * (size_t) i + d];
        acc[r].m = -INFINITY; acc[r].s = 0.f;
#pragma unroll
        for (int d = 0; d < 5; ++d) acc[r].a[d] = 0.f;
    }
    if (GRAD) {
        for (uint32_t j0 = 0; j0 < job.n_cols; j0 += SK_TILE) {
            const uint32_t nt = job.n_cols - j0 < SK_TILE ? job.n_cols - j0 : SK_TILE;
            __syncthreads();
            for (uint32_t k = threadIdx.x; k < nt; k += SK_THREADS) {
                const float *c = job.cols + 5 * (size_t) (j0 + k);
                const float h = job.pot_cols ? fmaf(job.pot_cols[j0 + k], inv_eps, job.w_log) : job.w_log;
                tile[2 * k] = make_float4(c[0], c[1], c[2], c[3]);
                tile[2 * k + 1] = make_float4(c[4], h * LOG2E, 0.f, 0.f);
            }
            __syncthreads();
            // warp w walks columns w, w + 8, … of the tile: every lane reads the same address (broadcast)
            for (uint32_t k = warp; k < nt; k += SK_THREADS / 32) {
                const float4 y0 = tile[2 * k], y1 = tile[2 * k + 1];
                visit<GRAD>(acc[0], x[0], y0, y1, halfk);
                visit<GRAD>(acc[1], x[1], y0, y1, halfk);
            }
        }
    } else {
        u64 xp[2][5]; RowAcc2 a2[2];
        const ulonglong2 *tp = reinterpret_cast<const ulonglong2 *>(tile);
        // Far-tile test.  Every term of a tile is at most hmax − halfk·d², d = distance between the boxes of this CTA's rows and
        // of the tile's points; once that is 40 binary orders below the running maximum of every row, the tile cannot change an
        // fp32 sum and is skipped (ε-scaling makes the kernel narrow: at ε = blur² most tiles of a pixel-grid cloud are far).
        // Tiles are visited from the nearest outwards so that the maxima are set early.
        float rlo[5], rhi[5];
#pragma unroll
        for (int d = 0; d < 5; ++d) {
            rlo[d] = fminf(x[0][d], x[1][d]); rhi[d] = fmaxf(x[0][d], x[1][d]);
            for (int of = 16; of > 0; of >>= 1) { rlo[d] = fminf(rlo[d], __shfl_xor_sync(0xffffffffu, rlo[d], of)); rhi[d] = fmaxf(rhi[d], __shfl_xor_sync(0xffffffffu, rhi[d], of)); }
        }
        // frame of this CTA: the centre of its rows' box; x̂ = (x − c)·2·halfk, r = halfk·|x − c|² (removed from the maximum at the end)
        float ctr[5], rsh[2];
#pragma unroll
        for (int d = 0; d < 5; ++d) ctr[d] = 0.5f * (rlo[d] + rhi[d]);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float q = 0.f;
#pragma unroll
            for (int d = 0; d < 5; ++d) { const float xc = x[r][d] - ctr[d]; q = fmaf(xc, xc, q); const float xh = xc * (2.f * halfk); xp[r][d] = pk(xh, xh); }
            rsh[r] = halfk * q;
            a2[r].m = -INFINITY; a2[r].s = pk(0.f, 0.f);
        }
        const uint32_t n_tiles = (job.n_cols + SK_TILE - 1) / SK_TILE;
        auto tile_bound = [&](uint32_t t) {
            const float *ti = job.tile_info + 12 * (size_t) t;
            float d2 = 0.f;
#pragma unroll
            for (int d = 0; d < 5; ++d) { const float g = fmaxf(fmaxf(ti[d] - rhi[d], rlo[d] - ti[5 + d]), 0.f); d2 = fmaf(g, g, d2); }
            return fmaf(-halfk, d2, ti[10]);
        };
        uint32_t t_near = 0; float b_near = -INFINITY;
        for (uint32_t t = 0; t < n_tiles; ++t) { const float b = tile_bound(t); if (b > b_near) { b_near = b; t_near = t; } }
        __shared__ float s_mmin[2];
        uint32_t visited = 0;
        for (uint32_t step = 0; step < 2 * n_tiles; ++step) {
            // step 0: the nearest tile; step 2o − 1: o tiles further; step 2o: o tiles nearer the start (all CTA-uniform)
            const uint32_t off = (step + 1) >> 1;
            uint32_t t;
            if (step == 0u) t = t_near;
            else if (step & 1u) { if (t_near + off >= n_tiles) continue; t = t_near + off; }
            else { if (off > t_near) continue; t = t_near - off; }
            const uint32_t slot = visited++ & 1u;                               // double-buffered: a skip has no second barrier
            if (warp == 0) {
                float mm = fminf(a2[0].m - rsh[0], a2[1].m - rsh[1]);                // running maxima in absolute terms
                for (int of = 16; of > 0; of >>= 1) mm = fminf(mm, __shfl_xor_sync(0xffffffffu, mm, of));
                if (lane == 0) s_mmin[slot] = mm;
            }
            __syncthreads();
            if (tile_bound(t) < s_mmin[slot] - 40.f) continue;                  // CTA-uniform
            const uint32_t j0 = t * SK_TILE;
            const uint32_t nt = job.n_cols - j0 < SK_TILE ? job.n_cols - j0 : SK_TILE, nq = (nt + 3) / 4;
            // the columns arrive packed by k_sink_pack: group of four columns = 6 float4
            // (y0a y0b y1a y1b | y2a y2b y3a y3b | y4a y4b ha hb) × 2 pairs, h = (ln w + pot/ε)·log2 e, tail padded with h = −∞
            {
                // one pair of columns (3 float4: y0a y0b y1a y1b | y2a y2b y3a y3b | y4a y4b ha hb) per thread and turn:
                // ỹ = y − c, h' = h − halfk·|ỹ|²
                const float4 *src = job.packed + (size_t) (j0 / 4) * 6;
                for (uint32_t k = threadIdx.x; k < 2 * nq; k += SK_THREADS) {
                    float4 a = src[3 * k], b = src[3 * k + 1], c = src[3 * k + 2];
                    a.x -= ctr[0]; a.y -= ctr[0]; a.z -= ctr[1]; a.w -= ctr[1]; b.x -= ctr[2]; b.y -= ctr[2]; b.z -= ctr[3]; b.w -= ctr[3]; c.x -= ctr[4]; c.y -= ctr[4];
                    const float qa = fmaf(c.x, c.x, fmaf(b.z, b.z, fmaf(b.x, b.x, fmaf(a.z, a.z, a.x * a.x))));
                    const float qb = fmaf(c.y, c.y, fmaf(b.w, b.w, fmaf(b.y, b.y, fmaf(a.w, a.w, a.y * a.y))));
                    c.z = fmaf(-halfk, qa, c.z); c.w = fmaf(-halfk, qb, c.w);
                    tile[3 * k] = a; tile[3 * k + 1] = b; tile[3 * k + 2] = c;
                }
            }
            __syncthreads();
            // warp w walks groups w, w + 8, … of four columns: every lane reads the same 96 bytes (broadcast)
            for (uint32_t k = warp; k < nq; k += SK_THREADS / 32) {
                const ulonglong2 t0 = tp[6 * k], t1 = tp[6 * k + 1], t2 = tp[6 * k + 2], u0 = tp[6 * k + 3], u1 = tp[6 * k + 4], u2 = tp[6 * k + 5];
                const u64 a0 = term2(xp[0], t0, t1, t2), b0 = term2(xp[0], u0, u1, u2);
                const u64 a1 = term2(xp[1], t0, t1, t2), b1 = term2(xp[1], u0, u1, u2);
                fold4(a2[0], a0, b0);
                fold4(a2[1], a1, b1);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) { float sa, sb; upk(a2[r].s, sa, sb); acc[r].m = a2[r].m - rsh[r]; acc[r].s = sa + sb; }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        float *o = red[warp][32 * r + lane];
        o[0] = acc[r].m; o[1] = acc[r].s;
        if (GRAD) { for (int d = 0; d < 5; ++d) o[2 + d] = acc[r].a[d]; }
    }
    __syncthreads();
    if (threadIdx.x < SK_ROWS) {
        const uint32_t i = row0 + threadIdx.x;
        if (i < job.n_rows) {
            float M = -INFINITY;
#pragma unroll
            for (int w = 0; w < SK_THREADS / 32; ++w) M = fmaxf(M, red[w][threadIdx.x][0]);
            float S = 0.f, A[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int w = 0; w < SK_THREADS / 32; ++w) {
                const float c = ex2(red[w][threadIdx.x][0] - M);
                S = fmaf(red[w][threadIdx.x][1], c, S);
                if (GRAD) { for (int d = 0; d < 5; ++d) A[d] = fmaf(red[w][threadIdx.x][2 + d], c, A[d]); }
            }
            float val = -st.eps * LN2 * (M + lg2(S));       // softmin(ε, C, h)_i
            if (st.average && job.pot_rows_old) val = 0.5f * (job.pot_rows_old[i] + val);
            job.pot_rows_new[i] = val;
            if (GRAD && job.bary) {
                const float inv = 1.0f / S;
                for (int d = 0; d < 5; ++d) job.bary[5 * (size_t) i + d] = A[d] * inv;
            }
        }
    }
}

// Columns of every softmin of a step in the layout the packed inner loop reads: one CTA per tile of SK_TILE columns
// (blockIdx.y = job).  Also writes the tile's record for the far-tile test of k_sink_softmin: bounding box of its points
// and the largest h·log2 e among them (12 floats: lo[5], hi[5], hmax, –).
__global__ void __launch_bounds__(256) k_sink_pack(SinkStep st) {
    const SinkJob job = st.job[blockIdx.y];
    const uint32_t j0 = blockIdx.x * SK_TILE;
    if (j0 >= job.n_cols) return;
    const uint32_t padded = (job.n_cols + 3u) & ~3u;
    const float inv_eps = 1.0f / st.eps;
    float lo[5], hi[5], hmax = -INFINITY;
#pragma unroll
    for (int d = 0; d < 5; ++d) { lo[d] = INFINITY; hi[d] = -INFINITY; }
    for (uint32_t k = j0 + threadIdx.x; k < j0 + SK_TILE && k < padded; k += 256) {
        float *o = reinterpret_cast<float *>(job.packed) + 12 * (size_t) (k >> 1) + (k & 1);
        if (k < job.n_cols) {
            const float *c = job.cols + 5 * (size_t) k;
            const float h = (job.pot_cols ? fmaf(job.pot_cols[k], inv_eps, job.w_log) : job.w_log) * LOG2E;
            o[0] = c[0]; o[2] = c[1]; o[4] = c[2]; o[6] = c[3]; o[8] = c[4]; o[10] = h;
#pragma unroll
            for (int d = 0; d < 5; ++d) { lo[d] = fminf(lo[d], c[d]); hi[d] = fmaxf(hi[d], c[d]); }
            hmax = fmaxf(hmax, h);
        } else { o[0] = o[2] = o[4] = o[6] = o[8] = 0.f; o[10] = -INFINITY; }
    }
    __shared__ float red[8][11];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 0; d < 5; ++d)
        for (int of = 16; of > 0; of >>= 1) { lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], of)); hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], of)); }
    for (int of = 16; of > 0; of >>= 1) hmax = fmaxf(hmax, __shfl_xor_sync(0xffffffffu, hmax, of));
    if (lane == 0) { for (int d = 0; d < 5; ++d) { red[warp][d] = lo[d]; red[warp][5 + d] = hi[d]; } red[warp][10] = hmax; }
    __syncthreads();
    if (threadIdx.x < 11) {
        float v = red[0][threadIdx.x];
        for (int w = 1; w < 8; ++w) v = threadIdx.x < 5 ? fminf(v, red[w][threadIdx.x]) : fmaxf(v, red[w][threadIdx.x]);
        job.tile_info[12 * (size_t) blockIdx.x + threadIdx.x] = v;
    }
}

// Bounding box of x ∪ y (geomloss max_diameter): out[0..4] = min, out[5..9] = max, as ordered-int atomics.
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__global__ void k_sink_bbox(const float *x, uint32_t n, const float *y, uint32_t m, int *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const float *p = i < n ? x + 5 * (size_t) i : (i < n + m ? y + 5 * (size_t) (i - n) : nullptr);
#pragma unroll
    for (int d = 0; d < 5; ++d) {
        int lo = p ? f2ord(p[d]) : 0x7fffffff, hi = p ? lo : (int) 0x80000000;
        lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
        if ((threadIdx.x & 31) == 0) { atomicMin(out + d, lo); atomicMax(out + 5 + d, hi); }
    }
}
__global__ void k_sink_bbox_init(int *out) { if (threadIdx.x < 5) { out[threadIdx.x] = 0x7fffffff; out[5 + threadIdx.x] = (int) 0x80000000; } }

// Z-curve order of a cloud along its last two coordinates (the pixel grid of the matcher's clouds, matcher.py:14-18): consecutive
// points are then spatially compact, so that the 1024-column tiles and the 64-row blocks of k_sink_softmin have small boxes and the
// far-tile test skips most tiles once ε is small.  Any permutation leaves the result unchanged up to the order of the fp32 sums.
__device__ __forceinline__ uint32_t spread16(uint32_t v) { v &= 0xffffu; v = (v | (v << 8)) & 0x00ff00ffu; v = (v | (v << 4)) & 0x0f0f0f0fu; v = (v | (v << 2)) & 0x33333333u; return (v | (v << 1)) & 0x55555555u; }
__global__ void k_sink_keys(const float *pts, uint32_t n, float lo3, float sc3, float lo4, float sc4, uint32_t *keys, uint32_t *idx) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = (pts[5 * (size_t) i + 3] - lo3) * sc3, b = (pts[5 * (size_t) i + 4] - lo4) * sc4;
    const uint32_t qa = (uint32_t) fminf(fmaxf(a, 0.f), 65535.f), qb = (uint32_t) fminf(fmaxf(b, 0.f), 65535.f);
    keys[i] = spread16(qa) | (spread16(qb) << 1); idx[i] = i;
}
__global__ void k_sink_gather(const float *src, const uint32_t *perm, uint32_t n, float *dst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = src + 5 * (size_t) perm[i];
#pragma unroll
    for (int d = 0; d < 5; ++d) dst[5 * (size_t) i + d] = p[d];
}

// g_i = factor·(bary_xx_i − bary_xy_i);  loss = Σ_i (f_ba − f_aa)_i / n + Σ_j (g_ab − g_bb)_j / m
__global__ void k_sink_finish(const float *bary_xx, const float *bary_xy, float factor, uint32_t n, float *grad, const uint32_t *perm,
                              const float *f_ba, const float *f_aa, const float *g_ab, const float *g_bb, uint32_t m, double *loss) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && grad) {
        const size_t o = perm ? perm[i] : i;                  // back to the caller's point order
#pragma unroll
        for (int d = 0; d < 5; ++d) grad[5 * o + d] = factor * (bary_xx[5 * (size_t) i + d] - bary_xy[5 * (size_t) i + d]);
    }
    double l = 0.0;
    if (i < n) l += ((double) f_ba[i] - (double) f_aa[i]) / (double) n;
    if (i < m) l += ((double) g_ab[i] - (double) g_bb[i]) / (double) m;
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    if ((threadIdx.x & 31) == 0 && l != 0.0) atomicAdd(loss, l);
}

} // namespace

cudaError_t epsm_launch_sink_bbox(const float *x, uint32_t n, const float *y, uint32_t m, int *out10, cudaStream_t s) {
    k_sink_bbox_init<<<1, 32, 0, s>>>(out10);
    k_sink_bbox<<<(n + m + 255) / 256, 256, 0, s>>>(x, n, y, m, out10);
    return cudaGetLastError();
}
float epsm_sink_ord_to_float(int o) { int i = o >= 0 ? o : o ^ 0x7fffffff; float f; memcpy(&f, &i, 4); return f; }

cudaError_t epsm_launch_sink_step(const SinkStep &st, int n_jobs, bool grad, cudaStream_t s) {
    uint32_t max_rows = 0;
    for (int k = 0; k < n_jobs; ++k) max_rows = st.job[k].n_rows > max_rows ? st.job[k].n_rows : max_rows;
    if (max_rows == 0) return cudaSuccess;
    const dim3 grid((max_rows + SK_ROWS - 1) / SK_ROWS, n_jobs);
    if (grad) k_sink_softmin<true><<<grid, SK_THREADS, 0, s>>>(st);
    else {
        uint32_t max_cols = 0;
        for (int k = 0; k < n_jobs; ++k) max_cols = st.job[k].n_cols > max_cols ? st.job[k].n_cols : max_cols;
        k_sink_pack<<<dim3((max_cols + SK_TILE - 1) / SK_TILE, n_jobs), 256, 0, s>>>(st);
        k_sink_softmin<false><<<grid, SK_THREADS, 0, s>>>(st);
    }
    return cudaGetLastError();
}

cudaError_t epsm_launch_sink_finish(const float *bary_xx, const float *bary_xy, float factor, uint32_t n, float *grad, const uint32_t *perm,
                                    const float *f_ba, const float *f_aa, const float *g_ab, const float *g_bb, uint32_t m, double *loss,
                                    cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(double), s);
    if (e != cudaSuccess) return e;
    const uint32_t k = n > m ? n : m;
    k_sink_finish<<<(k + 255) / 256, 256, 0, s>>>(bary_xx, bary_xy, factor, n, grad, perm, f_ba, f_aa, g_ab, g_bb, m, loss);
    return cudaGetLastError();
}
size_t epsm_sink_order_temp_bytes(uint32_t n) {
    size_t t = 0; cub::DeviceRadixSort::SortPairs(nullptr, t, (const uint32_t *) nullptr, (uint32_t *) nullptr, (const uint32_t *) nullptr, (uint32_t *) nullptr, (int) n);
    return (t + 255) & ~(size_t) 255;
}
// words: 4·n uint32 scratch (keys, keys sorted, idx, perm) — perm = words + 3n is the result; sorted = the reordered cloud.
cudaError_t epsm_launch_sink_order(const float *pts, uint32_t n, float lo3, float hi3, float lo4, float hi4, uint32_t *words, void *temp, size_t temp_bytes,
                                   float *sorted, cudaStream_t s) {
    const float sc3 = hi3 > lo3 ? 65535.f / (hi3 - lo3) : 0.f, sc4 = hi4 > lo4 ? 65535.f / (hi4 - lo4) : 0.f;
    uint32_t *keys = words, *keys2 = words + n, *idx = words + 2 * (size_t) n, *perm = words + 3 * (size_t) n;
    k_sink_keys<<<(n + 255) / 256, 256, 0, s>>>(pts, n, lo3, sc3, lo4, sc4, keys, idx);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys2, idx, perm, (int) n, 0, 32, s);
    if (e != cudaSuccess) return e;
    k_sink_gather<<<(n + 255) / 256, 256, 0, s>>>(pts, perm, n, sorted);
    return cudaGetLastError();
}
