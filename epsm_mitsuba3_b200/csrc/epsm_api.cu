// epsm_api.cu — the C ABI of include/epsm_b200.h over the kernels in epsm_kernels.cu.
// No torch types, no exceptions across the boundary, no CPU fallback.
#include <cuda_runtime.h>
#include <algorithm>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/epsm_b200.h"
#include "epsm_types.h"
#include "epsm_host_scene.h"
#include "epsm_launch.h"
#include "epsm_sinkhorn.h"

namespace {

std::string g_create_error;

template <typename T> struct DevBuf {
    T *p = nullptr; size_t n = 0;
    cudaError_t alloc(size_t count) {
        if (count <= n && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMalloc((void **) &p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    cudaError_t upload(const std::vector<T> &h, cudaStream_t s) {
        cudaError_t e = alloc(h.size());
        if (e != cudaSuccess || h.empty()) return e;
        return cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s);
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

} // namespace

struct epsm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string error;
    uint64_t launches = 0;
    bool have_scene = false, dirty = false;
    uint32_t s_begin = 0, s_end = 0;                // sample-range shard (0,0 = all)
    bool full_replay = false;                       // gradient pass without liveness pruning (epsm_set_full_replay)
    bool primal_megakernel = false;                 // EPSM_PRIMAL_MEGAKERNEL=1: render with the single-kernel k_render instead of the wavefront
    HostScene hs;                                   // host mirror (positions are needed for BVH rebuilds)
    std::vector<uint8_t> mesh_dirty;
    std::vector<uint8_t> host_pos_stale;            // meshes whose newest positions exist only on the device (pulled before a host rebuild)
    DevBuf<uint32_t> d_refit;                       // refit scratch: max|coordinate|, Σ half-area, per-node arrival counters
    uint64_t refits_since_build = 0; double refit_ms_last = 0;
    // device
    DevBuf<f4> d_nodes, d_tris; DevBuf<float> d_pos, d_nrm, d_cdf, d_pmf, d_uv, d_texels; DevBuf<DTexture> d_textures; DevBuf<uint32_t> d_faces, d_face_mesh;
    DevBuf<DMesh> d_meshes; DevBuf<DBsdf> d_bsdfs; DevBuf<DEmitter> d_emitters;
    std::vector<DevBuf<float>> d_films;
    DevBuf<float> stage_a, stage_b, stage_c;        // device staging for host-memory calls
    DevBuf<unsigned long long> d_counters;
    DevBuf<unsigned char> d_workspace; uint32_t wf_capacity = 0, trace_grid = 0;   // wavefront workspace (one chunk of lanes)
    DScene sc;
    bool trace_profiling = false; EpsmEventList trace_events;
    DevBuf<float> d_lane_L, d_scratch_film;          // prb_reparam: per-lane radiance of the primal pass; film of the re-render when the caller owns the film
    DevBuf<unsigned char> d_prb_workspace; uint32_t prb_capacity = 0, prb_ws_depth = 0, prb_ws_rays = 0;
    bool prb_megakernel = false;                     // EPSM_PRB_MEGAKERNEL=1: the per-lane megakernel (k_prb_backward) instead of the replay wavefront
    bool sink_reorder = true;            // Z-curve order of the matcher's clouds (EPSM_SINK_REORDER=0 at context creation: A/B measurements)
    DevBuf<float> d_sink;                           // Sinkhorn matcher: clouds (host-memory calls), potentials ×2, barycentres
    DevBuf<double> d_sink_loss; DevBuf<int> d_sink_bbox;
};

namespace {

int fail(epsm_ctx *c, int code, const std::string &msg) { if (c) c->error = msg; else g_create_error = msg; return code; }
int cuda_fail(epsm_ctx *c, cudaError_t e, const char *what) {
    return fail(c, EPSM_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call, what) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_fail(ctx, e__, what); } while (0)

int upload_accel(epsm_ctx *ctx) {
    CK(ctx->d_nodes.upload(ctx->hs.bvh.nodes, ctx->stream), "upload nodes");
    CK(ctx->d_tris.upload(ctx->hs.bvh.tris, ctx->stream), "upload tris");
    CK(ctx->d_cdf.upload(ctx->hs.cdf, ctx->stream), "upload cdf");
    CK(ctx->d_pmf.upload(ctx->hs.pmf, ctx->stream), "upload pmf");
    CK(ctx->d_meshes.upload(ctx->hs.meshes, ctx->stream), "upload meshes");
    DScene &sc = ctx->sc;
    sc.nodes = ctx->d_nodes.p; sc.tris = ctx->d_tris.p; sc.pos = ctx->d_pos.p; sc.nrm = ctx->d_nrm.p;
    sc.faces = ctx->d_faces.p; sc.face_mesh = ctx->d_face_mesh.p; sc.meshes = ctx->d_meshes.p; sc.bsdfs = ctx->d_bsdfs.p;
    sc.emitters = ctx->d_emitters.p; sc.cdf = ctx->d_cdf.p; sc.pmf = ctx->d_pmf.p;
    sc.uv = ctx->d_uv.p; sc.textures = ctx->d_textures.p; sc.texels = ctx->d_texels.p;
    sc.n_emitters = (uint32_t) ctx->hs.emitters.size(); sc.n_meshes = (uint32_t) ctx->hs.meshes.size();
    sc.n_tris = (uint32_t) (ctx->hs.faces.size() / 3); sc.n_nodes = (uint32_t) (ctx->hs.bvh.nodes.size() / EPSM_NODE_F4);
    memcpy(sc.filter, ctx->hs.filter, sizeof(sc.filter));
    CK(cudaStreamSynchronize(ctx->stream), "scene sync");
    return EPSM_OK;
}

// Resolve a user pointer to a device pointer, staging host memory through `stage`.
template <typename T>
int to_device(epsm_ctx *ctx, const T *user, size_t count, int mem, DevBuf<T> &stage, const T **out) {
    if (mem == EPSM_MEM_DEVICE) { *out = user; return EPSM_OK; }
    CK(stage.alloc(count), "stage alloc");
    CK(cudaMemcpyAsync(stage.p, user, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream), "H2D");
    *out = stage.p; return EPSM_OK;
}

int fetch_counters(epsm_ctx *ctx, uint64_t *counters) {
    unsigned long long h[4];
    CK(cudaMemcpyAsync(h, ctx->d_counters.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream), "counters D2H");
    CK(cudaStreamSynchronize(ctx->stream), "counters sync");
    for (int i = 0; i < 4; ++i) counters[i] += h[i];
    return EPSM_OK;
}
int zero_counters(epsm_ctx *ctx) {
    CK(ctx->d_counters.alloc(4), "counters alloc");
    CK(cudaMemsetAsync(ctx->d_counters.p, 0, 4 * sizeof(unsigned long long), ctx->stream), "counters memset");
    return EPSM_OK;
}

// Workspace for one chunk of the wavefront.  Bigger chunks mean fewer, fuller launches (the late iterations of a chunk are
// latency floors of a few thousand rays): the room pass takes 61.1 / 57.7 / 55.4 / 54.2 ms at 2^22 / 2^23 / 2^24 / 2^25 lanes
// (profiles/r01_h_*).  Default: the largest power of two whose workspace (≈ 1.5 KB per lane) fits a third of the free device
// memory, at most 2^25 lanes (48 GB of the B200's 180 GB), at least 2^20.  EPSM_WF_CHUNK overrides.
int ensure_workspace(epsm_ctx *ctx, uint64_t lanes) {
    if (ctx->d_workspace.p && ctx->wf_capacity >= lanes) return EPSM_OK;       // the whole wavefront already fits: one chunk
    uint64_t cap = 1ull << 22;
    if (const char *e = getenv("EPSM_WF_CHUNK")) { uint64_t v = strtoull(e, nullptr, 10); if (v >= 1024) cap = v; }
    else {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            free_b += ctx->d_workspace.n;                                   // what this context already holds counts as available
            cap = 1ull << 25;
            while (cap > (1ull << 20) && epsm_wf_workspace_bytes((uint32_t) cap) > free_b / 3) cap >>= 1;
        }
    }
    if (lanes < cap) cap = (lanes + 1023) & ~1023ull;
    if (cap < 1024) cap = 1024;
    if (ctx->wf_capacity >= cap && ctx->d_workspace.p) return EPSM_OK;
    CK(cudaStreamSynchronize(ctx->stream), "workspace sync");
    CK(ctx->d_workspace.alloc(epsm_wf_workspace_bytes((uint32_t) cap)), "workspace alloc");
    ctx->wf_capacity = (uint32_t) cap;
    return EPSM_OK;
}

int make_launch(epsm_ctx *ctx, uint32_t sensor, uint32_t seed, uint32_t spp, uint32_t max_depth, uint32_t rr_depth,
                uint64_t lane_begin, uint64_t lane_end, int variant, DLaunch &ln) {
    if (!ctx->have_scene) return fail(ctx, EPSM_E_NOSCENE, "no scene uploaded");
    if (ctx->dirty) return fail(ctx, EPSM_E_INVALID, "vertex updates pending: call epsm_scene_commit first");
    if (sensor >= ctx->hs.sensors.size()) return fail(ctx, EPSM_E_INVALID, "sensor index out of range");
    if (spp == 0) return fail(ctx, EPSM_E_INVALID, "spp must be > 0");
    if (rr_depth == 0) return fail(ctx, EPSM_E_INVALID, "\"rr_depth\" must be set to a value greater than zero!");   // common.py:40-41
    ln.sensor = ctx->hs.sensors[sensor];
    const uint64_t total = (uint64_t) ln.sensor.w * ln.sensor.h * spp;
    if (total > (1ull << 32)) return fail(ctx, EPSM_E_INVALID, "wavefront exceeds 2^32 samples");                       // common.py:468-475
    if (lane_end == 0 || lane_end > total) lane_end = total;
    if (lane_begin > lane_end) return fail(ctx, EPSM_E_INVALID, "lane_begin > lane_end");
    ln.seed = seed; ln.spp = spp; ln.max_depth = max_depth; ln.rr_depth = rr_depth;
    ln.lane_begin = lane_begin; ln.lane_end = lane_end; ln.variant = variant; ln.tile_y0 = ln.tile_y1 = 0;
    ln.s_begin = 0; ln.s_end = spp; ln.prune = ctx->full_replay ? 0u : 1u; ln.pixel_major = 0u;
    if (ctx->s_end != 0) {
        if (ctx->s_begin > ctx->s_end || ctx->s_end > spp) return fail(ctx, EPSM_E_INVALID, "sample range exceeds spp");
        ln.s_begin = ctx->s_begin; ln.s_end = ctx->s_end;
    }
    return EPSM_OK;
}

} // namespace

extern "C" {

int epsm_create(int device, epsm_ctx **out) {
    if (!out) return fail(nullptr, EPSM_E_INVALID, "out is NULL");
    *out = nullptr;
    int n = 0; cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, EPSM_E_NOGPU, std::string("no usable CUDA device (") + cudaGetErrorString(e) + "); libepsm_b200 has no CPU fallback");
    if (device < 0 || device >= n) return fail(nullptr, EPSM_E_INVALID, "device index out of range");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, EPSM_E_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    epsm_ctx *c = new epsm_ctx(); c->device = device;
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return fail(nullptr, EPSM_E_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e)); }
    memset(&c->sc, 0, sizeof(c->sc));
    if (const char *e = getenv("EPSM_FULL_REPLAY")) c->full_replay = atoi(e) != 0;
    if (const char *e = getenv("EPSM_SINK_REORDER")) c->sink_reorder = atoi(e) != 0;
    if (const char *e = getenv("EPSM_PRB_MEGAKERNEL")) c->prb_megakernel = atoi(e) != 0;
    if (const char *e = getenv("EPSM_PRIMAL_MEGAKERNEL")) c->primal_megakernel = atoi(e) != 0;
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, device);
    c->trace_grid = (uint32_t) prop.multiProcessorCount * 8u;        // persistent traversal CTAs: 148 SMs × 8 resident CTAs of 128 threads (≤ 64 regs)
    *out = c; return EPSM_OK;
}

int epsm_destroy(epsm_ctx *ctx) {
    if (!ctx) return EPSM_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->d_nodes.release(); ctx->d_tris.release(); ctx->d_pos.release(); ctx->d_nrm.release(); ctx->d_cdf.release(); ctx->d_pmf.release(); ctx->d_uv.release(); ctx->d_texels.release(); ctx->d_textures.release();
    ctx->d_faces.release(); ctx->d_face_mesh.release(); ctx->d_meshes.release(); ctx->d_bsdfs.release(); ctx->d_emitters.release();
    for (auto &f : ctx->d_films) f.release();
    ctx->stage_a.release(); ctx->stage_b.release(); ctx->stage_c.release(); ctx->d_counters.release(); ctx->d_workspace.release(); ctx->d_lane_L.release(); ctx->d_scratch_film.release(); ctx->d_prb_workspace.release(); ctx->d_sink.release(); ctx->d_sink_loss.release(); ctx->d_sink_bbox.release();
    cudaStreamDestroy(ctx->stream);
    delete ctx; return EPSM_OK;
}

const char *epsm_last_error(const epsm_ctx *ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int epsm_scene_upload(epsm_ctx *ctx, const epsm_mesh_desc *meshes, uint32_t n_meshes, const epsm_bsdf_desc *bsdfs, uint32_t n_bsdfs,
                      const epsm_emitter_desc *emitters, uint32_t n_emitters, const epsm_sensor_desc *sensors, uint32_t n_sensors) {
    if (!ctx) return EPSM_E_INVALID;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    std::string err = host_scene_build(ctx->hs, meshes, n_meshes, bsdfs, n_bsdfs, emitters, n_emitters, sensors, n_sensors);
    if (!err.empty()) return fail(ctx, EPSM_E_INVALID, err);
    for (auto &f : ctx->d_films) f.release();
    ctx->d_films.assign(n_sensors, DevBuf<float>());
    for (uint32_t i = 0; i < n_sensors; ++i) {
        const DSensor &s = ctx->hs.sensors[i];
        const size_t nf = (size_t) s.fw * s.fh * 4;
        CK(ctx->d_films[i].alloc(nf), "film alloc");
        CK(cudaMemsetAsync(ctx->d_films[i].p, 0, nf * sizeof(float), ctx->stream), "film memset");
    }
    CK(ctx->d_pos.upload(ctx->hs.pos, ctx->stream), "upload pos");
    CK(ctx->d_nrm.upload(ctx->hs.nrm, ctx->stream), "upload nrm");
    CK(ctx->d_uv.upload(ctx->hs.uv, ctx->stream), "upload uv");
    CK(ctx->d_texels.upload(ctx->hs.texels, ctx->stream), "upload texels");
    CK(ctx->d_textures.upload(ctx->hs.textures, ctx->stream), "upload textures");
    CK(ctx->d_faces.upload(ctx->hs.faces, ctx->stream), "upload faces");
    CK(ctx->d_face_mesh.upload(ctx->hs.face_mesh, ctx->stream), "upload face_mesh");
    CK(ctx->d_bsdfs.upload(ctx->hs.bsdfs, ctx->stream), "upload bsdfs");
    CK(ctx->d_emitters.upload(ctx->hs.emitters, ctx->stream), "upload emitters");
    ctx->mesh_dirty.assign(n_meshes, 0); ctx->host_pos_stale.assign(n_meshes, 0); ctx->refits_since_build = 0;
    int rc = upload_accel(ctx);
    if (rc != EPSM_OK) return rc;
    ctx->have_scene = true; ctx->dirty = false;
    return EPSM_OK;
}

int epsm_scene_update_vertices(epsm_ctx *ctx, uint32_t mesh_id, const float *positions, int mem) {
    if (!ctx) return EPSM_E_INVALID;
    if (!ctx->have_scene) return fail(ctx, EPSM_E_NOSCENE, "no scene uploaded");
    if (mesh_id >= ctx->hs.meshes.size() || !positions) return fail(ctx, EPSM_E_INVALID, "bad mesh id / positions");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    const DMesh &m = ctx->hs.meshes[mesh_id];
    const size_t nb = 3 * (size_t) m.v_cnt * sizeof(float);
    float *dst = ctx->d_pos.p + 3 * (size_t) m.v_off, *hdst = ctx->hs.pos.data() + 3 * (size_t) m.v_off;
    if (mem == EPSM_MEM_DEVICE) {
        CK(cudaMemcpyAsync(dst, positions, nb, cudaMemcpyDeviceToDevice, ctx->stream), "D2D positions");
        ctx->host_pos_stale[mesh_id] = 1;            // the host mirror is refreshed only when a host rebuild needs it
    } else {
        memcpy(hdst, positions, nb);
        CK(cudaMemcpyAsync(dst, hdst, nb, cudaMemcpyHostToDevice, ctx->stream), "H2D positions");
        ctx->host_pos_stale[mesh_id] = 0;
    }
    ctx->mesh_dirty[mesh_id] = 1; ctx->dirty = true;
    return EPSM_OK;
}

int epsm_scene_update_normals(epsm_ctx *ctx, uint32_t mesh_id, const float *normals, int mem) {
    if (!ctx) return EPSM_E_INVALID;
    if (!ctx->have_scene) return fail(ctx, EPSM_E_NOSCENE, "no scene uploaded");
    if (mesh_id >= ctx->hs.meshes.size() || !normals) return fail(ctx, EPSM_E_INVALID, "bad mesh id / normals");
    const DMesh &m = ctx->hs.meshes[mesh_id];
    if (!m.has_normals)                                                  // mesh.cpp:258-260
        return fail(ctx, EPSM_E_INVALID, "Storing new normals in a Mesh that didn't have normals at construction time is not implemented yet.");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    const size_t nb = 3 * (size_t) m.v_cnt * sizeof(float);
    float *dst = ctx->d_nrm.p + 3 * (size_t) m.v_off;
    CK(cudaMemcpyAsync(dst, normals, nb, mem == EPSM_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream), "normals copy");
    CK(cudaMemcpyAsync(ctx->hs.nrm.data() + 3 * (size_t) m.v_off, dst, nb, cudaMemcpyDeviceToHost, ctx->stream), "normals mirror");
    CK(cudaStreamSynchronize(ctx->stream), "normals sync");
    return EPSM_OK;
}

int epsm_scene_set_mesh_grad(epsm_ctx *ctx, uint32_t mesh_id, int enable) {
    if (!ctx) return EPSM_E_INVALID;
    if (!ctx->have_scene) return fail(ctx, EPSM_E_NOSCENE, "no scene uploaded");
    if (mesh_id >= ctx->hs.meshes.size()) return fail(ctx, EPSM_E_INVALID, "bad mesh id");
    DMesh &m = ctx->hs.meshes[mesh_id];
    const uint32_t f = enable ? (m.flags | EPSM_MESH_GRAD) : (m.flags & ~(uint32_t) EPSM_MESH_GRAD);
    if (f == m.flags) return EPSM_OK;
    m.flags = f;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(cudaMemcpyAsync(ctx->d_meshes.p + mesh_id, &m, sizeof(DMesh), cudaMemcpyHostToDevice, ctx->stream), "mesh flags");
    CK(cudaStreamSynchronize(ctx->stream), "mesh flags sync");
    return EPSM_OK;
}

int epsm_scene_update_sensor(epsm_ctx *ctx, uint32_t sensor, const epsm_sensor_desc *desc) {
    if (!ctx) return EPSM_E_INVALID;
    if (!ctx->have_scene) return fail(ctx, EPSM_E_NOSCENE, "no scene uploaded");
    if (!desc) return fail(ctx, EPSM_E_INVALID, "null sensor descriptor");
    std::string err = host_scene_set_sensor(ctx->hs, sensor, *desc);      // sensors live in the launch descriptor: host state only
    if (!err.empty()) return fail(ctx, EPSM_E_INVALID, err.c_str());
    return EPSM_OK;
}

int epsm_scene_vertex_buffers(epsm_ctx *ctx, float **positions, float **normals, uint64_t *n_floats) {
    if (!ctx) return EPSM_E_INVALID;
    if (!ctx->have_scene) return fail(ctx, EPSM_E_NOSCENE, "no scene uploaded");
    if (positions) *positions = ctx->d_pos.p;
    if (normals) *normals = ctx->d_nrm.p;
    if (n_floats) *n_floats = ctx->hs.pos.size();
    return EPSM_OK;
}

namespace {
// Bring the host mirror of the positions up to date (device-pointer updates are mirrored lazily).
int pull_stale_positions(epsm_ctx *ctx) {
    bool any = false;
    for (size_t i = 0; i < ctx->hs.meshes.size(); ++i) {
        if (!ctx->host_pos_stale[i]) continue;
        const DMesh &m = ctx->hs.meshes[i];
        CK(cudaMemcpyAsync(ctx->hs.pos.data() + 3 * (size_t) m.v_off, ctx->d_pos.p + 3 * (size_t) m.v_off, 3 * (size_t) m.v_cnt * sizeof(float),
                           cudaMemcpyDeviceToHost, ctx->stream), "D2H positions");
        ctx->host_pos_stale[i] = 0; any = true;
    }
    if (any) CK(cudaStreamSynchronize(ctx->stream), "positions sync");
    return EPSM_OK;
}
int recompute_dirty_normals(epsm_ctx *ctx) {
    // meshes are concatenated in the vertex / face arrays: one launch triple per run of adjacent dirty meshes with normals
    const size_t nm = ctx->hs.meshes.size();
    for (size_t i = 0; i < nm;) {
        const DMesh &m = ctx->hs.meshes[i];
        if (!ctx->mesh_dirty[i] || !m.has_normals) { ++i; continue; }
        size_t j = i + 1;
        while (j < nm && ctx->mesh_dirty[j] && ctx->hs.meshes[j].has_normals) ++j;
        const DMesh &l = ctx->hs.meshes[j - 1];
        CK(epsm_launch_vertex_normals(ctx->d_pos.p, ctx->d_faces.p, m.f_off, l.f_off + l.f_cnt - m.f_off, m.v_off, l.v_off + l.v_cnt - m.v_off,
                                      ctx->d_nrm.p, ctx->stream), "vertex normals");
        ctx->launches += 3;
        i = j;
    }
    return EPSM_OK;
}
} // namespace

int epsm_scene_commit(epsm_ctx *ctx) {
    if (!ctx) return EPSM_E_INVALID;
    if (!ctx->have_scene) return fail(ctx, EPSM_E_NOSCENE, "no scene uploaded");
    if (!ctx->dirty && ctx->refits_since_build == 0) return EPSM_OK;   // the tree is already a fresh build of these positions
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    int rc = recompute_dirty_normals(ctx);
    if (rc != EPSM_OK) return rc;
    rc = pull_stale_positions(ctx);
    if (rc != EPSM_OK) return rc;
    CK(cudaStreamSynchronize(ctx->stream), "commit sync");          // host mirror of the positions is complete
    std::fill(ctx->mesh_dirty.begin(), ctx->mesh_dirty.end(), 0);
    host_scene_rebuild_accel(ctx->hs);
    rc = upload_accel(ctx);
    if (rc != EPSM_OK) return rc;
    ctx->dirty = false; ctx->refits_since_build = 0;
    return EPSM_OK;
}

int epsm_scene_refit(epsm_ctx *ctx) {
    if (!ctx) return EPSM_E_INVALID;
    if (!ctx->have_scene) return fail(ctx, EPSM_E_NOSCENE, "no scene uploaded");
    if (!ctx->dirty) return EPSM_OK;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    int rc = recompute_dirty_normals(ctx);
    if (rc != EPSM_OK) return rc;
    const uint32_t n_nodes = ctx->sc.n_nodes, n_tris = ctx->sc.n_tris;
    CK(ctx->d_refit.alloc(epsm_refit_scratch_words(n_nodes)), "refit scratch alloc");
    CK(epsm_launch_refit(ctx->d_nodes.p, n_nodes, ctx->d_tris.p, n_tris, ctx->d_faces.p, ctx->d_pos.p, ctx->d_refit.p, ctx->stream), "k_refit");
    ctx->launches += 2;
    // emitter sampling tables (Mesh::build_pmf, mesh.cpp:361-393) of moved emitter meshes: small, rebuilt on the host
    bool emitter_moved = false;
    for (size_t i = 0; i < ctx->hs.meshes.size(); ++i) emitter_moved |= ctx->mesh_dirty[i] && ctx->hs.meshes[i].emitter >= 0;
    if (emitter_moved) {
        rc = pull_stale_positions(ctx);
        if (rc != EPSM_OK) return rc;
        host_scene_rebuild_area_tables(ctx->hs);
        CK(ctx->d_cdf.upload(ctx->hs.cdf, ctx->stream), "upload cdf");
        CK(ctx->d_pmf.upload(ctx->hs.pmf, ctx->stream), "upload pmf");
        CK(ctx->d_meshes.upload(ctx->hs.meshes, ctx->stream), "upload meshes");
        ctx->sc.cdf = ctx->d_cdf.p; ctx->sc.pmf = ctx->d_pmf.p; ctx->sc.meshes = ctx->d_meshes.p;
        CK(cudaStreamSynchronize(ctx->stream), "refit sync");        // the uploads read host vectors
    }
    std::fill(ctx->mesh_dirty.begin(), ctx->mesh_dirty.end(), 0);
    ctx->dirty = false; ctx->refits_since_build += 1;
    return EPSM_OK;
}

int epsm_scene_accel_cost(epsm_ctx *ctx, double *cost_built, double *cost_now, uint64_t *refits_since_build) {
    if (!ctx) return EPSM_E_INVALID;
    if (!ctx->have_scene) return fail(ctx, EPSM_E_NOSCENE, "no scene uploaded");
    if (cost_built) *cost_built = ctx->hs.bvh.cost;
    if (refits_since_build) *refits_since_build = ctx->refits_since_build;
    if (cost_now) {
        *cost_now = ctx->hs.bvh.cost;
        if (ctx->refits_since_build > 0 && ctx->d_refit.p) {
            CK(cudaSetDevice(ctx->device), "cudaSetDevice");
            CK(cudaMemcpyAsync(cost_now, ctx->d_refit.p + 2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream), "cost D2H");
            CK(cudaStreamSynchronize(ctx->stream), "cost sync");
        }
    }
    return EPSM_OK;
}

int epsm_scene_accel_raw(epsm_ctx *ctx, float **nodes, uint64_t *n_node_floats, float **tris, uint64_t *n_tri_floats) {
    if (!ctx) return EPSM_E_INVALID;
    if (!ctx->have_scene) return fail(ctx, EPSM_E_NOSCENE, "no scene uploaded");
    if (nodes) *nodes = reinterpret_cast<float *>(ctx->d_nodes.p);
    if (n_node_floats) *n_node_floats = (uint64_t) ctx->sc.n_nodes * EPSM_NODE_F4 * 4;
    if (tris) *tris = reinterpret_cast<float *>(ctx->d_tris.p);
    if (n_tri_floats) *n_tri_floats = (uint64_t) ctx->sc.n_tris * EPSM_TRI_F4 * 4;
    return EPSM_OK;
}

int epsm_trace(epsm_ctx *ctx, const float *rays, uint64_t n, void *hits, int any_hit, uint64_t *counters, int mem) {
    if (!ctx) return EPSM_E_INVALID;
    if (!ctx->have_scene) return fail(ctx, EPSM_E_NOSCENE, "no scene uploaded");
    if (ctx->dirty) return fail(ctx, EPSM_E_INVALID, "vertex updates pending: call epsm_scene_commit first");
    if (n == 0) return EPSM_OK;
    if (!rays || !hits) return fail(ctx, EPSM_E_INVALID, "rays/hits is NULL");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    const float *d_rays; int rc = to_device(ctx, rays, (size_t) n * 8, mem, ctx->stage_a, &d_rays);
    if (rc) return rc;
    const size_t hit_words = any_hit ? (size_t) n : (size_t) n * 4;
    void *d_hits = hits;
    if (mem == EPSM_MEM_HOST) { CK(ctx->stage_b.alloc(hit_words), "stage alloc"); d_hits = ctx->stage_b.p; }
    if (counters) { rc = zero_counters(ctx); if (rc) return rc; }
    CK(epsm_launch_trace(ctx->sc, d_rays, n, d_hits, any_hit, counters ? ctx->d_counters.p : nullptr, ctx->stream), "k_trace");
    ctx->launches += 1;
    if (mem == EPSM_MEM_HOST) {
        CK(cudaMemcpyAsync(hits, d_hits, hit_words * 4, cudaMemcpyDeviceToHost, ctx->stream), "D2H hits");
        CK(cudaStreamSynchronize(ctx->stream), "trace sync");
    }
    if (counters) return fetch_counters(ctx, counters);
    return EPSM_OK;
}

int epsm_film_clear(epsm_ctx *ctx, uint32_t sensor) {
    if (!ctx) return EPSM_E_INVALID;
    if (!ctx->have_scene) return fail(ctx, EPSM_E_NOSCENE, "no scene uploaded");
    if (sensor >= ctx->hs.sensors.size()) return fail(ctx, EPSM_E_INVALID, "sensor index out of range");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    const DSensor &s = ctx->hs.sensors[sensor];
    CK(cudaMemsetAsync(ctx->d_films[sensor].p, 0, (size_t) s.fw * s.fh * 4 * sizeof(float), ctx->stream), "film memset");
    return EPSM_OK;
}

int epsm_film_raw(epsm_ctx *ctx, uint32_t sensor, float **device_ptr, uint64_t *n_floats) {
    if (!ctx) return EPSM_E_INVALID;
    if (!ctx->have_scene) return fail(ctx, EPSM_E_NOSCENE, "no scene uploaded");
    if (sensor >= ctx->hs.sensors.size() || !device_ptr || !n_floats) return fail(ctx, EPSM_E_INVALID, "bad arguments");
    const DSensor &s = ctx->hs.sensors[sensor];
    *device_ptr = ctx->d_films[sensor].p; *n_floats = (uint64_t) s.fw * s.fh * 4;
    return EPSM_OK;
}

// film == nullptr: the context's film of `sensor`.  lane_L (device, 3 floats per global lane id; wavefront render only): each lane's radiance.
static int render_accumulate_impl(epsm_ctx *ctx, uint32_t sensor, uint32_t seed, uint32_t spp, uint32_t max_depth, uint32_t rr_depth,
                                  uint64_t lane_begin, uint64_t lane_end, uint64_t *counters, float *film, float *lane_L) {
    if (!ctx) return EPSM_E_INVALID;
    DLaunch ln; int rc = make_launch(ctx, sensor, seed, spp, max_depth, rr_depth, lane_begin, lane_end, 0, ln);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    if (!film) film = ctx->d_films[sensor].p;
    if (counters) { rc = zero_counters(ctx); if (rc) return rc; }
    if (ctx->primal_megakernel && !lane_L) {
        CK(epsm_launch_render(ctx->sc, ln, film, counters ? ctx->d_counters.p : nullptr, ctx->stream), "k_render");
        ctx->launches += 1;
    } else {
        const uint64_t S = (uint64_t) (ln.s_end - ln.s_begin);
        // lanes of the whole-tile enumeration that cover [lane_begin, lane_end)
        const uint64_t row_lanes = (uint64_t) ln.sensor.w * ln.spp;
        const uint64_t y0 = ln.lane_begin / row_lanes, y1 = std::min<uint64_t>((ln.lane_end + row_lanes - 1) / row_lanes, ln.sensor.h);
        const uint64_t tile_rows = y1 > y0 ? (y1 + 3) / 4 - y0 / 4 : 0;
        rc = ensure_workspace(ctx, (uint64_t) ((ln.sensor.w + 7) / 8) * tile_rows * 32ull * S); if (rc) return rc;
        CK(epsm_launch_render_wf(ctx->sc, ln, film, counters ? ctx->d_counters.p : nullptr, ctx->d_workspace.p, ctx->wf_capacity,
                                 ctx->trace_grid, &ctx->launches, ctx->stream, lane_L), "wavefront render");
    }
    if (counters) return fetch_counters(ctx, counters);
    return EPSM_OK;
}
int epsm_render_accumulate(epsm_ctx *ctx, uint32_t sensor, uint32_t seed, uint32_t spp, uint32_t max_depth, uint32_t rr_depth,
                           uint64_t lane_begin, uint64_t lane_end, uint64_t *counters) {
    return render_accumulate_impl(ctx, sensor, seed, spp, max_depth, rr_depth, lane_begin, lane_end, counters, nullptr, nullptr);
}

int epsm_film_develop(epsm_ctx *ctx, uint32_t sensor, float *out_hw5, int mem) {
    if (!ctx) return EPSM_E_INVALID;
    if (!ctx->have_scene) return fail(ctx, EPSM_E_NOSCENE, "no scene uploaded");
    if (sensor >= ctx->hs.sensors.size() || !out_hw5) return fail(ctx, EPSM_E_INVALID, "bad arguments");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    const DSensor &s = ctx->hs.sensors[sensor]; const uint64_t np = (uint64_t) s.fw * s.fh;
    float *d_out = out_hw5;
    if (mem == EPSM_MEM_HOST) { CK(ctx->stage_b.alloc(np * 5), "stage alloc"); d_out = ctx->stage_b.p; }
    CK(epsm_launch_develop(ctx->d_films[sensor].p, np, d_out, ctx->stream), "k_film_develop");
    ctx->launches += 1;
    if (mem == EPSM_MEM_HOST) {
        CK(cudaMemcpyAsync(out_hw5, d_out, np * 5 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream), "D2H image");
        CK(cudaStreamSynchronize(ctx->stream), "develop sync");
    }
    return EPSM_OK;
}

int epsm_render(epsm_ctx *ctx, uint32_t sensor, uint32_t seed, uint32_t spp, uint32_t max_depth, uint32_t rr_depth, float *out_hw5, int mem) {
    int rc = epsm_film_clear(ctx, sensor); if (rc) return rc;
    rc = epsm_render_accumulate(ctx, sensor, seed, spp, max_depth, rr_depth, 0, 0, nullptr); if (rc) return rc;
    return epsm_film_develop(ctx, sensor, out_hw5, mem);
}

int epsm_render_backward(epsm_ctx *ctx, int variant, uint32_t sensor, uint32_t seed, uint32_t spp, uint32_t max_depth, uint32_t rr_depth,
                         const float *grad_in_hw5, float *grad_pos, float *grad_nrm, int mem,
                         uint64_t lane_begin, uint64_t lane_end, uint64_t *counters) {
    if (!ctx) return EPSM_E_INVALID;
    if (variant != EPSM_VARIANT_MANIFOLD && variant != EPSM_VARIANT_MANIFOLD_CAUSTIC) return fail(ctx, EPSM_E_INVALID, "unknown integrator variant");
    if (!grad_in_hw5 || !grad_pos || !grad_nrm) return fail(ctx, EPSM_E_INVALID, "grad_in / grad_pos / grad_nrm is NULL");
    DLaunch ln; int rc = make_launch(ctx, sensor, seed, spp, max_depth, rr_depth, lane_begin, lane_end, variant, ln);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    const size_t n_img = (size_t) ln.sensor.fw * ln.sensor.fh * 5, n_v = ctx->hs.pos.size();
    const float *d_gin; rc = to_device(ctx, grad_in_hw5, n_img, mem, ctx->stage_a, &d_gin); if (rc) return rc;
    float *d_gp = grad_pos, *d_gn = grad_nrm;
    if (mem == EPSM_MEM_HOST) {
        CK(ctx->stage_b.alloc(n_v), "stage alloc"); CK(ctx->stage_c.alloc(n_v), "stage alloc");
        d_gp = ctx->stage_b.p; d_gn = ctx->stage_c.p;
        CK(cudaMemcpyAsync(d_gp, grad_pos, n_v * sizeof(float), cudaMemcpyHostToDevice, ctx->stream), "H2D grad_pos");
        CK(cudaMemcpyAsync(d_gn, grad_nrm, n_v * sizeof(float), cudaMemcpyHostToDevice, ctx->stream), "H2D grad_nrm");
    }
    if (counters) { rc = zero_counters(ctx); if (rc) return rc; }
    {
        const uint64_t S = (uint64_t) (ln.s_end - ln.s_begin);
        rc = ensure_workspace(ctx, (uint64_t) ((ln.sensor.w + 7) / 8) * ((ln.sensor.h + 3) / 4) * 32ull * S); if (rc) return rc;
    }
    CK(epsm_launch_backward(ctx->sc, ln, d_gin, d_gp, d_gn, nullptr, counters ? ctx->d_counters.p : nullptr, ctx->d_workspace.p, ctx->wf_capacity,
                            ctx->trace_grid, &ctx->launches, ctx->trace_profiling ? &ctx->trace_events : nullptr, ctx->stream), "wavefront backward");
    if (mem == EPSM_MEM_HOST) {
        CK(cudaMemcpyAsync(grad_pos, d_gp, n_v * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream), "D2H grad_pos");
        CK(cudaMemcpyAsync(grad_nrm, d_gn, n_v * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream), "D2H grad_nrm");
        CK(cudaStreamSynchronize(ctx->stream), "backward sync");
    }
    if (counters) return fetch_counters(ctx, counters);
    return EPSM_OK;
}

int epsm_prb_reparam_backward(epsm_ctx *ctx, uint32_t sensor, uint32_t seed, uint32_t spp, const epsm_reparam_config *cfg,
                              const float *grad_in_hw3, float *grad_pos, float *grad_nrm, int mem,
                              uint64_t lane_begin, uint64_t lane_end, int film_ready) {
    if (!ctx) return EPSM_E_INVALID;
    if (!cfg || !grad_in_hw3 || !grad_pos || !grad_nrm) return fail(ctx, EPSM_E_INVALID, "cfg / grad_in / grad_pos / grad_nrm is NULL");
    if (cfg->reparam_rays == 0 || cfg->reparam_rays > EPSM_PRB_MAX_RAYS) return fail(ctx, EPSM_E_INVALID, "reparam_rays must be in [1, 64]");
    if (!(cfg->reparam_kappa > 0.f)) return fail(ctx, EPSM_E_INVALID, "reparam_kappa must be positive");
    if (cfg->rr_depth == 0) return fail(ctx, EPSM_E_INVALID, "\"rr_depth\" must be set to a value greater than zero!");
    int rc;
    DLaunch ln; rc = make_launch(ctx, sensor, seed, spp, cfg->max_depth, cfg->rr_depth, lane_begin, lane_end, 0, ln);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    // replay wavefront (k_prb_pass over the persistent traversal kernel) unless the path is unbounded or the per-lane megakernel is asked for
    const bool wavefront = cfg->max_depth <= EPSM_PRB_WF_MAX_DEPTH && !ctx->prb_megakernel;
    const uint64_t n_lanes_total = (uint64_t) ln.sensor.w * ln.sensor.h * spp;
    float *lane_L = nullptr;
    if (wavefront) { CK(ctx->d_lane_L.alloc(3 * n_lanes_total), "lane radiance alloc"); lane_L = ctx->d_lane_L.p; }
    if (!film_ready) {                                   // primal pass of render_backward (common.py:905-935): the film of ALL lanes
        rc = epsm_film_clear(ctx, sensor); if (rc) return rc;
        rc = render_accumulate_impl(ctx, sensor, seed, spp, cfg->max_depth, cfg->rr_depth, 0, 0, nullptr, nullptr, lane_L); if (rc) return rc;
    } else if (wavefront) {                              // the caller rendered (and reduced) the film: only the lanes' radiance is needed
        CK(ctx->d_scratch_film.alloc((size_t) ln.sensor.fw * ln.sensor.fh * 4), "scratch film alloc");
        CK(cudaMemsetAsync(ctx->d_scratch_film.p, 0, (size_t) ln.sensor.fw * ln.sensor.fh * 4 * sizeof(float), ctx->stream), "scratch film clear");
        rc = render_accumulate_impl(ctx, sensor, seed, spp, cfg->max_depth, cfg->rr_depth, lane_begin, lane_end, nullptr, ctx->d_scratch_film.p, lane_L); if (rc) return rc;
    }
    const size_t n_img = (size_t) ln.sensor.fw * ln.sensor.fh * 3, n_v = ctx->hs.pos.size();
    const float *d_gin; rc = to_device(ctx, grad_in_hw3, n_img, mem, ctx->stage_a, &d_gin); if (rc) return rc;
    float *d_gp = grad_pos, *d_gn = grad_nrm;
    if (mem == EPSM_MEM_HOST) {
        CK(ctx->stage_b.alloc(n_v), "stage alloc"); CK(ctx->stage_c.alloc(n_v), "stage alloc");
        d_gp = ctx->stage_b.p; d_gn = ctx->stage_c.p;
        CK(cudaMemcpyAsync(d_gp, grad_pos, n_v * sizeof(float), cudaMemcpyHostToDevice, ctx->stream), "H2D grad_pos");
        CK(cudaMemcpyAsync(d_gn, grad_nrm, n_v * sizeof(float), cudaMemcpyHostToDevice, ctx->stream), "H2D grad_nrm");
    }
    PrbConfig pc{cfg->max_depth, cfg->rr_depth, cfg->reparam_max_depth, cfg->reparam_rays, cfg->reparam_kappa, cfg->reparam_exp};
    if (wavefront) {
        const uint64_t lanes = (uint64_t) ln.sensor.w * ln.sensor.h * (ln.s_end - ln.s_begin);
        uint64_t cap = 1ull << 23;
        if (const char *e = getenv("EPSM_PRB_CHUNK")) { uint64_t v = strtoull(e, nullptr, 10); if (v >= 1024) cap = v; }
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            free_b += ctx->d_prb_workspace.n;
            while (cap > (1ull << 16) && epsm_prb_workspace_bytes((uint32_t) cap, pc.max_depth, pc.n_rays) > free_b / 2) cap >>= 1;
        }
        if (lanes < cap) cap = (lanes + 1023) & ~1023ull;
        if (cap < 1024) cap = 1024;
        if (ctx->prb_capacity < cap || ctx->prb_ws_depth < pc.max_depth || ctx->prb_ws_rays < pc.n_rays) {
            CK(cudaStreamSynchronize(ctx->stream), "prb workspace sync");
            CK(ctx->d_prb_workspace.alloc(epsm_prb_workspace_bytes((uint32_t) cap, pc.max_depth, pc.n_rays)), "prb workspace alloc");
            ctx->prb_capacity = (uint32_t) cap; ctx->prb_ws_depth = pc.max_depth; ctx->prb_ws_rays = pc.n_rays;
        }
        CK(epsm_launch_prb_backward_wf(ctx->sc, ln, pc, ctx->d_films[sensor].p, lane_L, d_gin, d_gp, d_gn, ctx->d_prb_workspace.p, ctx->prb_capacity,
                                       ctx->trace_grid, &ctx->launches, ctx->stream), "k_prb_pass");
    } else {
        CK(epsm_launch_prb_backward(ctx->sc, ln, pc, ctx->d_films[sensor].p, d_gin, d_gp, d_gn, ctx->stream), "k_prb_backward");
        ctx->launches += 1;
    }
    if (mem == EPSM_MEM_HOST) {
        CK(cudaMemcpyAsync(grad_pos, d_gp, n_v * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream), "D2H grad_pos");
        CK(cudaMemcpyAsync(grad_nrm, d_gn, n_v * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream), "D2H grad_nrm");
        CK(cudaStreamSynchronize(ctx->stream), "prb backward sync");
    }
    return EPSM_OK;
}

int epsm_vertex_normals_backward(epsm_ctx *ctx, const float *grad_nrm, float *grad_pos, int mem) {
    if (!ctx) return EPSM_E_INVALID;
    if (!ctx->have_scene) return fail(ctx, EPSM_E_NOSCENE, "no scene uploaded");
    if (!grad_nrm || !grad_pos) return fail(ctx, EPSM_E_INVALID, "grad_nrm / grad_pos is NULL");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    const size_t n_v = ctx->hs.pos.size();
    const float *d_gn; int rc = to_device(ctx, grad_nrm, n_v, mem, ctx->stage_c, &d_gn); if (rc) return rc;
    float *d_gp = grad_pos;
    if (mem == EPSM_MEM_HOST) {
        CK(ctx->stage_b.alloc(n_v), "stage alloc"); d_gp = ctx->stage_b.p;
        CK(cudaMemcpyAsync(d_gp, grad_pos, n_v * sizeof(float), cudaMemcpyHostToDevice, ctx->stream), "H2D grad_pos");
    }
    CK(ctx->stage_a.alloc(n_v), "scratch alloc");
    const size_t nm = ctx->hs.meshes.size();
    auto wants = [&](const DMesh &m) { return m.has_normals && (m.flags & EPSM_MESH_GRAD); };
    for (size_t i = 0; i < nm;) {                       // one launch group per run of adjacent differentiable meshes with normals
        const DMesh &m = ctx->hs.meshes[i];
        if (!wants(m)) { ++i; continue; }
        size_t j = i + 1;
        while (j < nm && wants(ctx->hs.meshes[j])) ++j;
        const DMesh &l = ctx->hs.meshes[j - 1];
        CK(epsm_launch_vertex_normals_backward(ctx->d_pos.p, ctx->d_faces.p, m.f_off, l.f_off + l.f_cnt - m.f_off, m.v_off, l.v_off + l.v_cnt - m.v_off,
                                               d_gn, d_gp, ctx->stage_a.p, ctx->stream), "vertex normals backward");
        ctx->launches += 4;
        i = j;
    }
    if (mem == EPSM_MEM_HOST) {
        CK(cudaMemcpyAsync(grad_pos, d_gp, n_v * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream), "D2H grad_pos");
        CK(cudaStreamSynchronize(ctx->stream), "normals backward sync");
    }
    return EPSM_OK;
}

int epsm_path_records(epsm_ctx *ctx, int variant, uint32_t sensor, uint32_t seed, uint32_t spp, uint32_t max_depth, uint32_t rr_depth,
                      const float *grad_in_hw5, uint64_t lane_begin, uint64_t lane_end, float *records, int mem) {
    if (!ctx) return EPSM_E_INVALID;
    if (variant != EPSM_VARIANT_MANIFOLD && variant != EPSM_VARIANT_MANIFOLD_CAUSTIC) return fail(ctx, EPSM_E_INVALID, "unknown integrator variant");
    if (!grad_in_hw5 || !records) return fail(ctx, EPSM_E_INVALID, "grad_in / records is NULL");
    DLaunch ln; int rc = make_launch(ctx, sensor, seed, spp, max_depth, rr_depth, lane_begin, lane_end, variant, ln);
    if (rc) return rc;
    ln.prune = 0;                                    // records are those of the full replay (every logged vertex of epsm.py:648-654)
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    const size_t n_img = (size_t) ln.sensor.fw * ln.sensor.fh * 5, n_rec = (size_t) (ln.lane_end - ln.lane_begin) * EPSM_REC_FLOATS;
    const float *d_gin; rc = to_device(ctx, grad_in_hw5, n_img, mem, ctx->stage_a, &d_gin); if (rc) return rc;
    float *d_rec = records;
    if (mem == EPSM_MEM_HOST) { CK(ctx->stage_b.alloc(n_rec), "stage alloc"); d_rec = ctx->stage_b.p; }
    CK(cudaMemsetAsync(d_rec, 0, n_rec * sizeof(float), ctx->stream), "records memset");
    {
        const uint64_t S = (uint64_t) (ln.s_end - ln.s_begin);
        rc = ensure_workspace(ctx, (uint64_t) ((ln.sensor.w + 7) / 8) * ((ln.sensor.h + 3) / 4) * 32ull * S); if (rc) return rc;
    }
    CK(epsm_launch_backward(ctx->sc, ln, d_gin, nullptr, nullptr, d_rec, nullptr, ctx->d_workspace.p, ctx->wf_capacity, ctx->trace_grid,
                            &ctx->launches, nullptr, ctx->stream), "wavefront backward (records)");
    if (mem == EPSM_MEM_HOST) {
        CK(cudaMemcpyAsync(records, d_rec, n_rec * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream), "D2H records");
        CK(cudaStreamSynchronize(ctx->stream), "records sync");
    }
    return EPSM_OK;
}

int epsm_sampler_stream(epsm_ctx *ctx, uint32_t seed, uint64_t lane_begin, uint64_t lane_end, uint32_t n_floats, float *out, int mem) {
    if (!ctx) return EPSM_E_INVALID;
    if (lane_end < lane_begin || !out) return fail(ctx, EPSM_E_INVALID, "bad arguments");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    const uint64_t n = lane_end - lane_begin; const size_t cnt = (size_t) n * n_floats;
    if (cnt == 0) return EPSM_OK;
    float *d = out;
    if (mem == EPSM_MEM_HOST) { CK(ctx->stage_b.alloc(cnt), "stage alloc"); d = ctx->stage_b.p; }
    CK(epsm_launch_sampler_stream(seed, lane_begin, n, n_floats, d, ctx->stream), "k_sampler_stream");
    ctx->launches += 1;
    if (mem == EPSM_MEM_HOST) {
        CK(cudaMemcpyAsync(out, d, cnt * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream), "D2H");
        CK(cudaStreamSynchronize(ctx->stream), "sync");
    }
    return EPSM_OK;
}

int epsm_sample_rays(epsm_ctx *ctx, uint32_t sensor, uint32_t seed, uint32_t spp, uint64_t lane_begin, uint64_t lane_end, float *out, int mem) {
    if (!ctx) return EPSM_E_INVALID;
    if (!ctx->have_scene) return fail(ctx, EPSM_E_NOSCENE, "no scene uploaded");
    if (sensor >= ctx->hs.sensors.size() || spp == 0 || lane_end < lane_begin || !out) return fail(ctx, EPSM_E_INVALID, "bad arguments");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    const uint64_t n = lane_end - lane_begin; const size_t cnt = (size_t) n * 14;
    if (cnt == 0) return EPSM_OK;
    float *d = out;
    if (mem == EPSM_MEM_HOST) { CK(ctx->stage_b.alloc(cnt), "stage alloc"); d = ctx->stage_b.p; }
    CK(epsm_launch_sample_rays(ctx->hs.sensors[sensor], seed, spp, lane_begin, n, d, ctx->stream), "k_sample_rays");
    ctx->launches += 1;
    if (mem == EPSM_MEM_HOST) {
        CK(cudaMemcpyAsync(out, d, cnt * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream), "D2H");
        CK(cudaStreamSynchronize(ctx->stream), "sync");
    }
    return EPSM_OK;
}

int epsm_scene_get_stats(epsm_ctx *ctx, epsm_scene_stats *out) {
    if (!ctx || !out) return EPSM_E_INVALID;
    if (!ctx->have_scene) return fail(ctx, EPSM_E_NOSCENE, "no scene uploaded");
    out->n_triangles = ctx->hs.faces.size() / 3; out->n_vertices = ctx->hs.pos.size() / 3;
    out->n_nodes = ctx->hs.bvh.nodes.size() / EPSM_NODE_F4;
    out->node_bytes = EPSM_NODE_F4 * 16; out->tri_bytes = EPSM_TRI_BYTES; out->bvh_width = 4; out->max_leaf = EPSM_MAX_LEAF;
    out->build_ms = ctx->hs.bvh.build_ms;
    return EPSM_OK;
}

int epsm_set_sample_range(epsm_ctx *ctx, uint32_t s_begin, uint32_t s_end) {
    if (!ctx) return EPSM_E_INVALID;
    if (s_begin > s_end) return fail(ctx, EPSM_E_INVALID, "s_begin > s_end");
    ctx->s_begin = s_begin; ctx->s_end = s_end;
    return EPSM_OK;
}

int epsm_set_full_replay(epsm_ctx *ctx, int enable) {
    if (!ctx) return EPSM_E_INVALID;
    ctx->full_replay = enable != 0;
    return EPSM_OK;
}

int epsm_sinkhorn_grad(epsm_ctx *ctx, const float *x, uint32_t n, const float *y, uint32_t m, float blur, float scaling, float scale,
                       float *grad_x, double *loss, uint32_t *n_steps, int mem) {
    if (!ctx) return EPSM_E_INVALID;
    if (!x || !y || n == 0 || m == 0) return fail(ctx, EPSM_E_INVALID, "sinkhorn: empty point cloud");
    if (!(blur > 0.f) || !(scaling > 0.f && scaling < 1.f)) return fail(ctx, EPSM_E_INVALID, "sinkhorn: blur must be > 0 and 0 < scaling < 1");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t s = ctx->stream;
    // device layout of d_sink: [x 5n | y 5m]? (host calls only) then f_ba, f_aa (n) ×2, g_ab, g_bb (m) ×2, bary_xy, bary_xx (5n), grad (5n, host calls)
    const size_t cloud = mem == EPSM_MEM_HOST ? 5 * ((size_t) n + m) : 0, pots = 4 * ((size_t) n + m), bary = 10 * (size_t) n;
    const size_t gout = mem == EPSM_MEM_HOST ? 5 * (size_t) n : 0;
    const size_t packed = 2 * (epsm_sink_packed_floats(n) + epsm_sink_packed_floats(m));       // four jobs: columns y, x, x, y (16-byte aligned: first)
    const size_t tinfo = 2 * (epsm_sink_tile_floats(n) + epsm_sink_tile_floats(m));
    // Z-curve reordering (epsm_launch_sink_order): sorted copies of both clouds, 4 words per point of sort scratch, cub's temp storage
    const bool reorder = ctx->sink_reorder;
    const size_t order_temp = reorder ? epsm_sink_order_temp_bytes(n > m ? n : m) / sizeof(float) : 0;
    const size_t order = reorder ? 5 * ((size_t) n + m) + 4 * ((size_t) n + m) + order_temp + 64 : 0;
    CK(ctx->d_sink.alloc(packed + tinfo + cloud + pots + bary + gout + order), "sinkhorn alloc");
    CK(ctx->d_sink_loss.alloc(1), "sinkhorn alloc"); CK(ctx->d_sink_bbox.alloc(10), "sinkhorn alloc");
    float4 *pk_y0 = reinterpret_cast<float4 *>(ctx->d_sink.p), *pk_x0 = pk_y0 + epsm_sink_packed_floats(m) / 4,
           *pk_x1 = pk_x0 + epsm_sink_packed_floats(n) / 4, *pk_y1 = pk_x1 + epsm_sink_packed_floats(n) / 4;
    float *ti_y0 = ctx->d_sink.p + packed, *ti_x0 = ti_y0 + epsm_sink_tile_floats(m), *ti_x1 = ti_x0 + epsm_sink_tile_floats(n),
          *ti_y1 = ti_x1 + epsm_sink_tile_floats(n);
    float *base = ctx->d_sink.p + packed + tinfo;
    const float *dx = x, *dy = y;
    if (mem == EPSM_MEM_HOST) {
        CK(cudaMemcpyAsync(base, x, 5 * (size_t) n * sizeof(float), cudaMemcpyHostToDevice, s), "H2D x");
        CK(cudaMemcpyAsync(base + 5 * (size_t) n, y, 5 * (size_t) m * sizeof(float), cudaMemcpyHostToDevice, s), "H2D y");
        dx = base; dy = base + 5 * (size_t) n;
    }
    float *p = base + cloud;
    float *f_ba[2] = {p, p + n}, *f_aa[2] = {p + 2 * (size_t) n, p + 3 * (size_t) n}; p += 4 * (size_t) n;
    float *g_ab[2] = {p, p + m}, *g_bb[2] = {p + 2 * (size_t) m, p + 3 * (size_t) m}; p += 4 * (size_t) m;
    float *bary_xy = p, *bary_xx = p + 5 * (size_t) n; p += bary;
    float *d_grad = mem == EPSM_MEM_HOST ? p : grad_x;
    p += gout;
    float *sorted_x = p, *sorted_y = p + 5 * (size_t) n;
    uint32_t *words_x = reinterpret_cast<uint32_t *>(p + 5 * ((size_t) n + m)), *words_y = words_x + 4 * (size_t) n;
    void *order_tmp = reinterpret_cast<void *>((reinterpret_cast<uintptr_t>(words_y + 4 * (size_t) m) + 255) & ~(uintptr_t) 255);
    // ε-schedule (geomloss epsilon_schedule, p = 2) from the diameter of the joint bounding box (max_diameter)
    CK(epsm_launch_sink_bbox(dx, n, dy, m, ctx->d_sink_bbox.p, s), "k_sink_bbox"); ctx->launches += 2;
    int hb[10];
    CK(cudaMemcpyAsync(hb, ctx->d_sink_bbox.p, sizeof(hb), cudaMemcpyDeviceToHost, s), "bbox D2H");
    CK(cudaStreamSynchronize(s), "bbox sync");
    float d2 = 0.f;
    for (int d = 0; d < 5; ++d) { const float e = epsm_sink_ord_to_float(hb[5 + d]) - epsm_sink_ord_to_float(hb[d]); d2 += e * e; }
    const double diameter = (double) sqrtf(d2);
    if (!(diameter > 0.0)) return fail(ctx, EPSM_E_INVALID, "sinkhorn: degenerate point clouds (zero diameter)");
    const uint32_t *perm_x = nullptr;
    if (reorder) {
        const float lo3 = epsm_sink_ord_to_float(hb[3]), hi3 = epsm_sink_ord_to_float(hb[8]), lo4 = epsm_sink_ord_to_float(hb[4]), hi4 = epsm_sink_ord_to_float(hb[9]);
        CK(epsm_launch_sink_order(dx, n, lo3, hi3, lo4, hi4, words_x, order_tmp, order_temp * sizeof(float), sorted_x, s), "k_sink_order");
        CK(epsm_launch_sink_order(dy, m, lo3, hi3, lo4, hi4, words_y, order_tmp, order_temp * sizeof(float), sorted_y, s), "k_sink_order");
        ctx->launches += 4;
        dx = sorted_x; dy = sorted_y; perm_x = words_x + 3 * (size_t) n;
    }
    std::vector<double> eps_list; eps_list.push_back(diameter * diameter);
    {   // numpy.arange(2 ln d, 2 ln blur, 2 ln scaling): ceil((stop − start)/step) values start + k·step
        const double start = 2.0 * log(diameter), stop = 2.0 * log((double) blur), step = 2.0 * log((double) scaling);
        const long cnt = (long) ceil((stop - start) / step);
        for (long k = 0; k < cnt; ++k) eps_list.push_back(exp(start + (double) k * step));
    }
    eps_list.push_back((double) blur * (double) blur);
    const float la = -logf((float) n), lb = -logf((float) m);
    auto make = [&](int cur, int nxt, bool init) {
        SinkStep st; memset(&st, 0, sizeof(st));
        // f_ba: rows x, cols y, carries g_ab;  g_ab: rows y, cols x, carries f_ba;  f_aa: x,x carries f_aa;  g_bb: y,y carries g_bb
        st.job[0] = SinkJob{dx, dy, init ? nullptr : g_ab[cur], f_ba[cur], f_ba[nxt], nullptr, n, m, lb, 0, pk_y0, ti_y0};
        st.job[1] = SinkJob{dy, dx, init ? nullptr : f_ba[cur], g_ab[cur], g_ab[nxt], nullptr, m, n, la, 0, pk_x0, ti_x0};
        st.job[2] = SinkJob{dx, dx, init ? nullptr : f_aa[cur], f_aa[cur], f_aa[nxt], nullptr, n, n, la, 0, pk_x1, ti_x1};
        st.job[3] = SinkJob{dy, dy, init ? nullptr : g_bb[cur], g_bb[cur], g_bb[nxt], nullptr, m, m, lb, 0, pk_y1, ti_y1};
        return st;
    };
    int cur = 0;
    {   // initialisation at ε₀: potentials ← softmin(ε₀, C, ln weight)
        SinkStep st = make(1, 0, true); st.eps = (float) eps_list[0]; st.average = 0;
        CK(epsm_launch_sink_step(st, 4, false, s), "k_sink_softmin"); ctx->launches += 2;
    }
    for (double eps : eps_list) {
        SinkStep st = make(cur, cur ^ 1, false); st.eps = (float) eps; st.average = 1;
        CK(epsm_launch_sink_step(st, 4, false, s), "k_sink_softmin"); ctx->launches += 2;
        cur ^= 1;
    }
    {   // last extrapolation (un-averaged) + the softmax-weighted barycentres the gradient needs
        SinkStep st = make(cur, cur ^ 1, false); st.eps = (float) eps_list.back(); st.average = 0;
        st.job[0].bary = bary_xy; st.job[2].bary = bary_xx;
        CK(epsm_launch_sink_step(st, 4, true, s), "k_sink_softmin"); ctx->launches += 1;
        cur ^= 1;
    }
    CK(epsm_launch_sink_finish(bary_xx, bary_xy, scale / (float) n, n, grad_x ? d_grad : nullptr, perm_x, f_ba[cur], f_aa[cur], g_ab[cur], g_bb[cur], m,
                               ctx->d_sink_loss.p, s), "k_sink_finish");
    ctx->launches += 1;
    if (n_steps) *n_steps = (uint32_t) eps_list.size();
    if (mem == EPSM_MEM_HOST && grad_x) CK(cudaMemcpyAsync(grad_x, d_grad, 5 * (size_t) n * sizeof(float), cudaMemcpyDeviceToHost, s), "D2H grad");
    if (loss) CK(cudaMemcpyAsync(loss, ctx->d_sink_loss.p, sizeof(double), cudaMemcpyDeviceToHost, s), "D2H loss");
    if (loss || mem == EPSM_MEM_HOST) CK(cudaStreamSynchronize(s), "sinkhorn sync");
    return EPSM_OK;
}

int epsm_set_trace_profiling(epsm_ctx *ctx, int enable) {
    if (!ctx) return EPSM_E_INVALID;
    ctx->trace_profiling = enable != 0;
    return EPSM_OK;
}

int epsm_get_trace_profile(epsm_ctx *ctx, double *ms, uint64_t *launches) {
    if (!ctx || !ms || !launches) return EPSM_E_INVALID;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(cudaStreamSynchronize(ctx->stream), "sync");
    double total = 0;
    for (auto &e : ctx->trace_events) {
        float t = 0; cudaEventElapsedTime(&t, e.first, e.second); total += t;
        cudaEventDestroy(e.first); cudaEventDestroy(e.second);
    }
    *ms = total; *launches = ctx->trace_events.size();
    ctx->trace_events.clear();
    return EPSM_OK;
}

uint64_t epsm_launch_count(const epsm_ctx *ctx) { return ctx ? ctx->launches : 0; }
void *epsm_stream(epsm_ctx *ctx) { return ctx ? (void *) ctx->stream : nullptr; }
int epsm_synchronize(epsm_ctx *ctx) {
    if (!ctx) return EPSM_E_INVALID;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(cudaStreamSynchronize(ctx->stream), "sync");
    return EPSM_OK;
}

} // extern "C"
