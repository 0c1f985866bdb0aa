// epsm_bvh.cpp — binned-SAH BVH build on the host, collapsed into the 128-byte BVH4 node layout.
#include "epsm_bvh.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <thread>

#ifndef EPSM_SAH_BINS
#define EPSM_SAH_BINS 32   // measured on the room: 16 → 43.4 nodes per path, 32 → 42.4 (traversal −2 %), 64 → 42.8
#endif

#ifndef EPSM_SAH_CT
#define EPSM_SAH_CT 2.0f     // cost of visiting a node relative to one triangle test; measured on the room (nodes / triangles per path, traversal ms): 0.5 → 42.3 / 7.4, 21.7; 1 → 40.9 / 10.3, 20.9; 2 → 38.7 / 16.5, 20.7; ≥ 3 saturates (every leaf full); 8-triangle leaves lose
#endif

namespace {

struct Box {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; ++a) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; } }
    void grow(const float *p) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    void grow(const Box &b) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    float half_area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return (dx < 0 || dy < 0 || dz < 0) ? 0.f : dx * dy + dy * dz + dz * dx;
    }
};
struct BNode { Box box; uint32_t left, right, first, count; };   // count > 0: leaf over prims[first, first+count)

struct Builder {
    const float *pos; const uint32_t *faces; size_t nf;
    std::vector<Box> pbox; std::vector<float> cent;            // per-primitive bounds and centroids (3·nf)
    std::vector<uint32_t> prims;
    std::vector<BNode> nodes; std::atomic<uint32_t> n_nodes{0};
    std::atomic<int> threads_left{0};

    uint32_t alloc2() { return n_nodes.fetch_add(2); }

    void build(uint32_t node, uint32_t first, uint32_t count, int depth) {
        BNode &n = nodes[node];
        Box cb; cb.reset(); n.box.reset();
        for (uint32_t i = first; i < first + count; ++i) { uint32_t p = prims[i]; n.box.grow(pbox[p]); cb.grow(&cent[3 * p]); }
        n.first = first; n.count = count; n.left = n.right = 0;
        if (count <= 1) return;
        // binned SAH over the three axes
        const int NB = EPSM_SAH_BINS;
        float best_cost = FLT_MAX; int best_axis = -1, best_bin = -1;
        for (int a = 0; a < 3; ++a) {
            float lo = cb.lo[a], ext = cb.hi[a] - lo;
            if (!(ext > 0)) continue;
            Box bb[NB]; uint32_t bc[NB];
            for (int b = 0; b < NB; ++b) { bb[b].reset(); bc[b] = 0; }
            float scale = NB / ext;
            for (uint32_t i = first; i < first + count; ++i) {
                uint32_t p = prims[i]; int b = std::min(NB - 1, std::max(0, (int) ((cent[3 * p + a] - lo) * scale)));
                bb[b].grow(pbox[p]); bc[b]++;
            }
            float ra[NB]; uint32_t rc[NB]; Box acc; acc.reset(); uint32_t c = 0;
            for (int b = NB - 1; b > 0; --b) { acc.grow(bb[b]); c += bc[b]; ra[b] = acc.half_area(); rc[b] = c; }
            acc.reset(); c = 0;
            for (int b = 0; b < NB - 1; ++b) {
                acc.grow(bb[b]); c += bc[b];
                if (c == 0 || rc[b + 1] == 0) continue;
                float cost = acc.half_area() * c + ra[b + 1] * rc[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = b; }
            }
        }
        float leaf_cost = n.box.half_area() * count;
        if (count <= EPSM_MAX_LEAF && (best_axis < 0 || leaf_cost <= best_cost + EPSM_SAH_CT * n.box.half_area())) return;   // leaf
        uint32_t mid;
        if (best_axis >= 0) {
            float lo = cb.lo[best_axis], scale = NB / (cb.hi[best_axis] - lo);
            auto it = std::partition(prims.begin() + first, prims.begin() + first + count, [&](uint32_t p) {
                int b = std::min(NB - 1, std::max(0, (int) ((cent[3 * p + best_axis] - lo) * scale))); return b <= best_bin; });
            mid = (uint32_t) (it - prims.begin());
        } else mid = first;
        if (mid == first || mid == first + count) mid = first + count / 2;       // coincident centroids: split by index
        uint32_t l = alloc2();
        nodes[node].left = l; nodes[node].right = l + 1; nodes[node].count = 0;
        uint32_t lc = mid - first, rcnt = first + count - mid;
        if (count > 32768 && threads_left.fetch_sub(1) > 0) {
            std::thread t([=] { build(l, first, lc, depth + 1); });
            build(l + 1, mid, rcnt, depth + 1);
            t.join();
        } else {
            build(l, first, lc, depth + 1);
            build(l + 1, mid, rcnt, depth + 1);
        }
    }
};

struct Collapser {
    const Builder &B; Bvh4 &out; float pad;
    uint32_t max_depth = 0;
    // Emits binary subtree `b` as a wide node; returns its index.
    // `parent` = (parent node << 2) | slot, 0xffffffff for the root (kept in the node's pad word for the device refit).
    uint32_t emit(uint32_t b, uint32_t depth, uint32_t parent = 0xffffffffu) {
        max_depth = std::max(max_depth, depth);
        uint32_t ch[4]; int nc = 0;
        ch[nc++] = B.nodes[b].left; ch[nc++] = B.nodes[b].right;
        while (nc < 4) {
            int best = -1; float ba = -1.f;
            for (int i = 0; i < nc; ++i) {
                const BNode &c = B.nodes[ch[i]];
                if (c.count == 0 && c.box.half_area() > ba) { ba = c.box.half_area(); best = i; }
            }
            if (best < 0) break;
            uint32_t c = ch[best]; ch[best] = B.nodes[c].left; ch[nc++] = B.nodes[c].right;
        }
        uint32_t idx = (uint32_t) (out.nodes.size() / EPSM_NODE_F4);
        out.nodes.resize(out.nodes.size() + EPSM_NODE_F4);
        float lo[3][4], hi[3][4]; int32_t ref[4];
        for (int i = 0; i < 4; ++i) { for (int a = 0; a < 3; ++a) { lo[a][i] = 1e30f; hi[a][i] = -1e30f; } ref[i] = 0x7fffffff; }
        for (int i = 0; i < nc; ++i) {
            const BNode &c = B.nodes[ch[i]];
            for (int a = 0; a < 3; ++a) { lo[a][i] = c.box.lo[a] - pad; hi[a][i] = c.box.hi[a] + pad; }
            if (c.count > 0) ref[i] = ~(int32_t) ((emit_leaf(c) << EPSM_LEAF_BITS) | (c.count - 1));
        }
        uint32_t n_inner = 0;
        for (int i = 0; i < nc; ++i) {
            if (B.nodes[ch[i]].count == 0) { ref[i] = (int32_t) emit(ch[i], depth + 1, (idx << 2) | (uint32_t) i); ++n_inner; }
        }
        f4 *n = &out.nodes[(size_t) idx * EPSM_NODE_F4];
        for (int a = 0; a < 3; ++a) { n[a] = f4{lo[a][0], lo[a][1], lo[a][2], lo[a][3]}; n[3 + a] = f4{hi[a][0], hi[a][1], hi[a][2], hi[a][3]}; }
        float rf[4]; memcpy(rf, ref, 16);
        n[6] = f4{rf[0], rf[1], rf[2], rf[3]};
        const uint32_t link[2] = {parent, n_inner}; float lf[2]; memcpy(lf, link, 8);
        n[7] = f4{lf[0], lf[1], 0, 0};                          // never read by the traversal
        return idx;
    }
    uint32_t emit_leaf(const BNode &c) {
        uint32_t first = (uint32_t) (out.tris.size() / EPSM_TRI_F4);
        for (uint32_t i = c.first; i < c.first + c.count; ++i) {
            uint32_t p = B.prims[i]; const uint32_t *f = B.faces + 3 * (size_t) p;
            const float *p0 = B.pos + 3 * (size_t) f[0], *p1 = B.pos + 3 * (size_t) f[1], *p2 = B.pos + 3 * (size_t) f[2];
            float e1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, e2[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
            float pf; memcpy(&pf, &p, 4);
            out.tris.push_back(f4{p0[0], p0[1], p0[2], pf});
            out.tris.push_back(f4{e1[0], e1[1], e1[2], e2[0]});
            out.tris.push_back(f4{e2[1], e2[2], 0.f, 0.f});
            out.tris.push_back(f4{0.f, 0.f, 0.f, 0.f});
        }
        return first;
    }
};

} // namespace

// Σ over nodes of the half-area of the union of the node's child boxes: the traversal-cost proxy the refit policy watches.
double epsm_bvh4_cost(const f4 *nodes, size_t n_nodes) {
    double sum = 0;
    for (size_t i = 0; i < n_nodes; ++i) {
        const float *n = &nodes[i * EPSM_NODE_F4].x;
        int32_t ref[4]; memcpy(ref, n + 24, 16);
        float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
        for (int s = 0; s < 4; ++s) {
            if (ref[s] == 0x7fffffff) continue;
            for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], n[4 * a + s]); hi[a] = std::max(hi[a], n[12 + 4 * a + s]); }
        }
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx >= 0 && dy >= 0 && dz >= 0) sum += (double) (dx * dy + dy * dz + dz * dx);
    }
    return sum;
}

void epsm_build_bvh4(const float *pos, const uint32_t *faces, size_t nf, Bvh4 &out) {
    auto t0 = std::chrono::steady_clock::now();
    out.nodes.clear(); out.tris.clear(); out.depth = 0;
    if (nf == 0) { out.build_ms = 0; return; }
    Builder B; B.pos = pos; B.faces = faces; B.nf = nf;
    B.pbox.resize(nf); B.cent.resize(3 * nf); B.prims.resize(nf);
    float max_abs = 0.f;
    for (size_t i = 0; i < nf; ++i) {
        Box b; b.reset();
        for (int k = 0; k < 3; ++k) b.grow(pos + 3 * (size_t) faces[3 * i + k]);
        B.pbox[i] = b; B.prims[i] = (uint32_t) i;
        for (int a = 0; a < 3; ++a) { B.cent[3 * i + a] = 0.5f * (b.lo[a] + b.hi[a]); max_abs = std::max(max_abs, std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a]))); }
    }
    B.nodes.resize(2 * nf + 2); B.n_nodes = 1;
    B.threads_left = (int) std::min(15u, std::max(1u, std::thread::hardware_concurrency()) - 1u);
    B.build(0, 0, (uint32_t) nf, 0);
    // Padding: Möller–Trumbore accepts hits whose true position lies a few ulp outside the exact box.
    float pad = 1e-5f * max_abs + 1e-7f;
    Collapser C{B, out, pad};
    out.nodes.reserve((size_t) (nf / 2 + 8) * EPSM_NODE_F4); out.tris.reserve(nf * EPSM_TRI_F4);
    if (B.nodes[0].count > 0) {
        // single-leaf scene: wrap it in one wide node
        out.nodes.resize(EPSM_NODE_F4);
        float lo[3][4], hi[3][4]; int32_t ref[4];
        for (int i = 0; i < 4; ++i) { for (int a = 0; a < 3; ++a) { lo[a][i] = 1e30f; hi[a][i] = -1e30f; } ref[i] = 0x7fffffff; }
        for (int a = 0; a < 3; ++a) { lo[a][0] = B.nodes[0].box.lo[a] - pad; hi[a][0] = B.nodes[0].box.hi[a] + pad; }
        ref[0] = ~(int32_t) ((C.emit_leaf(B.nodes[0]) << EPSM_LEAF_BITS) | (B.nodes[0].count - 1));
        f4 *n = &out.nodes[0];
        for (int a = 0; a < 3; ++a) { n[a] = f4{lo[a][0], lo[a][1], lo[a][2], lo[a][3]}; n[3 + a] = f4{hi[a][0], hi[a][1], hi[a][2], hi[a][3]}; }
        float rf[4]; memcpy(rf, ref, 16); n[6] = f4{rf[0], rf[1], rf[2], rf[3]};
        const uint32_t root_link = 0xffffffffu; float lf; memcpy(&lf, &root_link, 4); n[7] = f4{lf, 0, 0, 0};
    } else {
        C.emit(0, 1);
    }
    out.depth = C.max_depth;
    out.pad = pad;
    out.cost = epsm_bvh4_cost(out.nodes.data(), out.nodes.size() / EPSM_NODE_F4);
    out.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
