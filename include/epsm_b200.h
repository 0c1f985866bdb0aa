/* epsm_b200.h — C ABI of libepsm_b200.so, the B200-native EPSM manifold-integrator path.
 *
 * Drop-in boundary for the hot path of jkxing/EPSM_Mitsuba3.  Every entry
 * point cites the reference interface it replaces (paths relative to the reference root).
 * Plain pointers and sizes only; no C++ or torch types cross this boundary.  All calls
 * return 0 on success and a negative EPSM_E_* code on failure; epsm_last_error() gives the
 * message.  Nothing throws across the ABI.  One CUDA stream per context; calls on one
 * context are not thread-safe (matches the reference: film guarded by a mutex,
 * src/films/hdrfilm.cpp:247-252; integrators driven single-threaded from Python).
 *
 * Memory spaces: pointer arguments with a companion `mem` argument live on the host
 * (EPSM_MEM_HOST; the library stages through pinned memory and the copies are part of the
 * call) or on the context's device (EPSM_MEM_DEVICE, e.g. a torch tensor's data_ptr()).
 */
#ifndef EPSM_B200_H
#define EPSM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EPSM_MEM_HOST   0
#define EPSM_MEM_DEVICE 1

#define EPSM_OK            0
#define EPSM_E_INVALID    -1   /* bad argument                         */
#define EPSM_E_CUDA       -2   /* CUDA runtime error                   */
#define EPSM_E_NOSCENE    -3   /* call needs an uploaded scene         */
#define EPSM_E_NOGPU      -4   /* no usable CUDA device (no fallback)  */

/* BSDF plugin kinds on the path. */
#define EPSM_BSDF_DIFFUSE         0   /* src/bsdfs/diffuse.cpp        */
#define EPSM_BSDF_DIELECTRIC      1   /* src/bsdfs/dielectric.cpp     */
#define EPSM_BSDF_CONDUCTOR       2   /* src/bsdfs/conductor.cpp      */
#define EPSM_BSDF_ROUGHCONDUCTOR  3   /* src/bsdfs/roughconductor.cpp */

/* epsm_bsdf_desc.flags */
#define EPSM_BSDF_TWOSIDED        1u  /* wrapped in src/bsdfs/twosided.cpp (same BRDF both sides) */
#define EPSM_BSDF_GGX             2u  /* microfacet distribution: GGX instead of Beckmann          */
#define EPSM_BSDF_SAMPLE_VISIBLE  4u  /* roughconductor sample_visible (default true)              */

/* Mitsuba BSDFFlags bits that the path inspects (include/mitsuba/render/bsdf.h:37-110). */
#define EPSM_BF_NULL               0x00001u
#define EPSM_BF_DIFFUSE_REFLECTION 0x00002u
#define EPSM_BF_GLOSSY_REFLECTION  0x00008u
#define EPSM_BF_DELTA_REFLECTION   0x00020u
#define EPSM_BF_DELTA_TRANSMISSION 0x00040u
#define EPSM_BF_DIFFUSE            0x00006u
#define EPSM_BF_SMOOTH             0x0001eu
#define EPSM_BF_DELTA              0x00061u

/* epsm_mesh_desc.flags */
#define EPSM_MESH_ISMESH        1u  /* a triangle mesh in the reference (si.ismesh = 1, mesh.cpp:720);
                                       clear for tessellated analytic shapes (rectangle, sphere …),
                                       which the reference excludes from EPSM gradients (epsm.py:793-796) */
#define EPSM_MESH_GRAD          2u  /* vertex_positions has gradient tracking enabled                 */
#define EPSM_MESH_FLIP_NORMALS  4u  /* Mesh::m_flip_normals                                           */
#define EPSM_MESH_RECTANGLE     8u  /* tessellation of a Mitsuba `rectangle`: vertices 0..3 are the corners
                                       (-1,-1) (1,-1) (1,1) (-1,1) under to_world; emitter sampling then follows
                                       src/shapes/rectangle.cpp:152-171 (uniform in the parameterisation)    */

typedef struct epsm_mesh_desc {
    uint32_t        vertex_count;
    uint32_t        face_count;
    const float    *positions;   /* 3*vertex_count, host                                     */
    const uint32_t *faces;       /* 3*face_count, host, indices local to this mesh           */
    const float    *normals;     /* 3*vertex_count or NULL (flat shading), host              */
    int32_t         bsdf_id;     /* index into the bsdf table                                */
    int32_t         emitter_id;  /* index into the emitter table, or -1                      */
    uint32_t        flags;       /* EPSM_MESH_*                                              */
    const float    *texcoords;   /* 2*vertex_count or NULL, host: Mesh::vertex_texcoord (mesh.cpp:737-763): si.uv and the
                                    dp_du / dp_dv that orient the shading frame follow them                               */
} epsm_mesh_desc;

/* src/textures/bitmap.cpp as a diffuse reflectance: RGB texels already converted to linear floats (raw = false applies the
   inverse sRGB gamma at load time), looked up at to_uv * si.uv with Dr.Jit's Texture (ext/drjit/include/drjit/texture.h). */
#define EPSM_TEX_NEAREST  0u
#define EPSM_TEX_BILINEAR 1u
#define EPSM_TEX_REPEAT   0u
#define EPSM_TEX_CLAMP    1u
#define EPSM_TEX_MIRROR   2u
typedef struct epsm_texture_desc {
    const float *texels;         /* 3*width*height, row-major, row 0 first (as Bitmap stores them), host */
    uint32_t     width, height;
    uint32_t     filter;         /* EPSM_TEX_NEAREST | EPSM_TEX_BILINEAR                                  */
    uint32_t     wrap;           /* EPSM_TEX_REPEAT | EPSM_TEX_CLAMP | EPSM_TEX_MIRROR                    */
    float        to_uv[6];       /* rows 0 and 1 of the 3x3 uv transform (identity: 1 0 0 / 0 1 0)         */
} epsm_texture_desc;

typedef struct epsm_bsdf_desc {
    uint32_t kind;               /* EPSM_BSDF_*                                              */
    uint32_t flags;              /* EPSM_BSDF_TWOSIDED | …                                   */
    float    reflectance[3];     /* diffuse reflectance / specular_reflectance               */
    float    transmittance[3];   /* dielectric specular_transmittance                        */
    float    eta[3];             /* conductor eta (RGB); dielectric: eta[0] = int_ior/ext_ior */
    float    k[3];               /* conductor k (RGB)                                        */
    float    alpha_u, alpha_v;   /* roughconductor roughness                                 */
    const epsm_texture_desc *reflectance_texture;   /* diffuse only: replaces `reflectance` where non-NULL (src/bsdfs/diffuse.cpp:107-140) */
} epsm_bsdf_desc;

typedef struct epsm_emitter_desc {   /* src/emitters/area.cpp, uniform radiance */
    float   radiance[3];
    int32_t mesh_id;
} epsm_emitter_desc;

/* epsm_sensor_desc.flags */
#define EPSM_SENSOR_SAMPLE_BORDER 1u /* hdrfilm `sample_border` (src/render/film.cpp; common.py:318-343): the wavefront also covers the
                                        reconstruction filter's border (2 pixels for the Gaussian) around the film, lane → pixel
                                        uses the enlarged grid, splats outside the film are cropped.  The EPSM drivers set it
                                        on sensor 0 (PRB / reference renders), not on sensors 1 and 2 (EPSM/exp/shadow.py:98-153) */

typedef struct epsm_sensor_desc {    /* src/sensors/perspective.cpp + hdrfilm + gaussian rfilter */
    float    to_world[16];       /* row-major camera-to-world, no scale                      */
    float    fov_x;              /* degrees, horizontal                                      */
    float    near_clip, far_clip;
    uint32_t width, height;      /* film size == crop size (no crop window)                  */
    uint32_t flags;              /* EPSM_SENSOR_*                                            */
} epsm_sensor_desc;

typedef struct epsm_ctx epsm_ctx;

/* Integrator variants registered by the reference (epsm.py:948, :1202). */
#define EPSM_VARIANT_MANIFOLD          0
#define EPSM_VARIANT_MANIFOLD_CAUSTIC  1

/* Context on CUDA device `device`.  Fails with EPSM_E_NOGPU when no device is usable:
 * there is no CPU fallback.  (Replaces mi.set_variant('cuda_ad_rgb'), EPSM/optim.py:18.) */
int epsm_create(int device, epsm_ctx **out);
int epsm_destroy(epsm_ctx *ctx);
const char *epsm_last_error(const epsm_ctx *ctx);   /* ctx may be NULL: last creation error */

/* Scene construction: replaces mi.load_dict / Scene::Scene + accel init
 * (src/render/scene.cpp:22-100, src/render/scene_embree.inl:83-187, scene_optix.inl:236-440).
 * Copies all host arrays, builds the wide BVH, area-sampling tables and sensor transforms. */
int epsm_scene_upload(epsm_ctx *ctx,
                      const epsm_mesh_desc *meshes, uint32_t n_meshes,
                      const epsm_bsdf_desc *bsdfs, uint32_t n_bsdfs,
                      const epsm_emitter_desc *emitters, uint32_t n_emitters,
                      const epsm_sensor_desc *sensors, uint32_t n_sensors);

/* params['<mesh>.vertex_positions'] = …; params.update()  (src/python/python/util.py:181-228 →
 * Mesh::parameters_changed, src/render/mesh.cpp:82-102 → Scene::parameters_changed, scene.cpp:304-336).
 * `positions` holds 3*vertex_count floats in `mem` space.  Vertex normals of meshes that have them
 * are recomputed (mesh.cpp:257-330), area tables and the BVH are rebuilt on the host on commit.  epsm_scene_commit
 * after epsm_scene_refit (no new updates) replaces the refitted tree by a fresh build of the current positions. */
int epsm_scene_update_vertices(epsm_ctx *ctx, uint32_t mesh_id, const float *positions, int mem);
int epsm_scene_commit(epsm_ctx *ctx);
/* dr.enable_grad(params['<mesh>.vertex_positions']) at the boundary: sets / clears EPSM_MESH_GRAD of a mesh after upload, so
 * that the gradient passes scatter into exactly the meshes whose buffers are attached (util.py:352-403 hands `params` to
 * render_backward; only attached entries receive gradients). */
int epsm_scene_set_mesh_grad(epsm_ctx *ctx, uint32_t mesh_id, int enable);
/* params['<sensor>.to_world'] = …; params.update() (src/sensors/perspective.cpp:134-171 parameters_changed →
 * update_camera_transforms): replaces the pose / fov / clip planes of sensor `sensor`; the film size must not change. */
int epsm_scene_update_sensor(epsm_ctx *ctx, uint32_t sensor, const epsm_sensor_desc *desc);
/* params['<mesh>.vertex_normals'] = …; params.update() (mesh.cpp:71-75: the key is exposed as differentiable; a change of
 * the normals alone triggers no recomputation, mesh.cpp:82-83).  3*vertex_count floats in `mem` space.  Fails for meshes
 * created without vertex normals, as the reference does (mesh.cpp:258-260).  The next epsm_scene_update_vertices of the
 * mesh re-derives the normals from the positions again. */
int epsm_scene_update_normals(epsm_ctx *ctx, uint32_t mesh_id, const float *normals, int mem);
/* Library-owned device pointers to the concatenated vertex positions / normals (3·total vertex count floats each;
 * meshes without normals hold zeros) — the current values of the `vertex_positions` / `vertex_normals` parameters. */
int epsm_scene_vertex_buffers(epsm_ctx *ctx, float **positions, float **normals, uint64_t *n_floats);

/* The same update with the acceleration structure REFITTED on the device instead of rebuilt on the host: the topology
 * of the last build is kept, triangle records and every box are recomputed from the new positions by two kernels
 * (k_refit_tris, k_refit_nodes), vertex normals as in commit; nothing is copied to the host unless an emitter mesh
 * moved (its area tables, Mesh::build_pmf mesh.cpp:361-393, are small and rebuilt on the host).  The reference rebuilds
 * its Embree / OptiX structure on every parameters_changed (scene.cpp:304-336, scene_embree.inl:130-187,
 * scene_optix.inl:236-440); hit records do not depend on the structure (closest t, ties to the larger primitive index),
 * so refit and rebuild return identical results and differ only in traversal cost.  A refit with unchanged positions
 * reproduces the host build bit for bit. */
int epsm_scene_refit(epsm_ctx *ctx);
/* Traversal-cost proxy for the refit / rebuild decision: Σ over nodes of the half-area of the node's box, for the tree
 * as last BUILT and as it is NOW (after refits; reads one double back).  Any pointer may be NULL. */
int epsm_scene_accel_cost(epsm_ctx *ctx, double *cost_built, double *cost_now, uint64_t *refits_since_build);
/* Library-owned device pointers to the BVH4 nodes (32 floats each) and triangle slots (16 floats each), for tests. */
int epsm_scene_accel_raw(epsm_ctx *ctx, float **nodes, uint64_t *n_node_floats, float **tris, uint64_t *n_tri_floats);

/* Scene::ray_intersect_preliminary / Scene::ray_test on a ray batch
 * (src/render/scene_embree.inl:202-426, scene_optix.inl:516-646, kdtree.h:2188-2290).
 * rays:  n × 8 floats (o.xyz, maxt, d.xyz, pad).  mint = 0.
 * hits (closest): n × 4 32-bit words (t, prim_u, prim_v as float; global prim_index as uint32,
 *       0xffffffff on a miss; t = +inf on a miss).
 * hits (any_hit != 0): n × uint32, 1 if occluded.
 * counters (optional, HOST pointer, 4 × uint64, accumulated into): BVH nodes fetched, triangles
 * tested, rays traced, lanes processed — the algorithmic-byte accounting of SURVEY.md §8d.  Pass NULL
 * to run the uninstrumented kernel (the instrumented one has identical control flow). */
int epsm_trace(epsm_ctx *ctx, const float *rays, uint64_t n, void *hits, int any_hit,
               uint64_t *counters, int mem);

/* Integrator::render for `manifold*` (EPSMIntegrator.render, epsm.py:13-82;
 * include/mitsuba/render/integrator.h:74-93).  Traces lanes [lane_begin, lane_end) of the
 * W·H·spp wavefront (lane_end = 0 means all) and ACCUMULATES into the context's RGBW film.
 * max_depth / rr_depth are the integrator properties (common.py:28-43). */
int epsm_render_accumulate(epsm_ctx *ctx, uint32_t sensor, uint32_t seed, uint32_t spp,
                           uint32_t max_depth, uint32_t rr_depth,
                           uint64_t lane_begin, uint64_t lane_end, uint64_t *counters);
int epsm_film_clear(epsm_ctx *ctx, uint32_t sensor);
/* Raw RGBW film (H·W·4 floats) — the buffer a multi-GPU host allreduces before develop. */
int epsm_film_raw(epsm_ctx *ctx, uint32_t sensor, float **device_ptr, uint64_t *n_floats);
/* HDRFilm::develop + the two zero channels appended by epsm.py:76-82: out is H·W·5 floats. */
int epsm_film_develop(epsm_ctx *ctx, uint32_t sensor, float *out_hw5, int mem);
/* Convenience: clear + accumulate(all lanes) + develop. */
int epsm_render(epsm_ctx *ctx, uint32_t sensor, uint32_t seed, uint32_t spp,
                uint32_t max_depth, uint32_t rr_depth, float *out_hw5, int mem);

/* Integrator::render_backward for `manifold*` (EPSMIntegrator.render_backward, epsm.py:84-306,
 * with sample_path :503-742 and calc_grad :745-946 / :952-1200 fused per path).
 * grad_in: H·W·5 floats for `sensor` (the reference hard-codes sensors()[2] and spp = 8,
 * epsm.py:142-145 — the host mirror passes those).  grad_pos / grad_nrm: 3·(total vertex count)
 * floats, concatenated over meshes in upload order, ACCUMULATED into (dr.grad semantics).
 * Only meshes flagged EPSM_MESH_GRAD receive gradients.  Lanes as in epsm_render_accumulate. */
int epsm_render_backward(epsm_ctx *ctx, int variant, uint32_t sensor, uint32_t seed, uint32_t spp,
                         uint32_t max_depth, uint32_t rr_depth,
                         const float *grad_in_hw5, float *grad_pos, float *grad_nrm, int mem,
                         uint64_t lane_begin, uint64_t lane_end, uint64_t *counters);

/* Integrator::render_backward for `prb_reparam` — hybrid phase 2 of the EPSM drivers (EPSM/optim.py:89-92,113-119):
 * RBIntegrator.render_backward (src/python/python/ad/integrators/common.py:803-975) with PRBReparamIntegrator.sample
 * (prb_reparam.py:277-607) and reparameterize_ray (ad/reparam.py:10-463).  Properties as in prb_reparam.py:225-247 (antithetic
 * sampling and loop unrolling are evaluation strategies of the reference's AD graph and do not exist here; antithetic = false).
 * grad_in: H·W·3 floats for the film of `sensor`; grad_pos / grad_nrm as in epsm_render_backward (accumulated into).
 * film_ready = 0: the call first renders the primal film of all lanes of (sensor, seed, spp) into the context's film — the
 * weights the image gradient is normalised with.  film_ready = 1: the caller has done that (epsm_film_clear +
 * epsm_render_accumulate with the same seed / spp / depths, summed over ranks when the samples are sharded). */
typedef struct epsm_reparam_config {
    uint32_t max_depth, rr_depth;        /* ADIntegrator properties (common.py:28-43); max_depth −1 → 0xffffffff   */
    uint32_t reparam_max_depth;          /* default = max_depth (prb_reparam.py:232)                                */
    uint32_t reparam_rays;               /* auxiliary rays per reparameterisation, default 16, at most 64           */
    float    reparam_kappa, reparam_exp; /* von Mises–Fisher concentration (1e5), harmonic-weight exponent (3)      */
} epsm_reparam_config;
int epsm_prb_reparam_backward(epsm_ctx *ctx, uint32_t sensor, uint32_t seed, uint32_t spp, const epsm_reparam_config *cfg,
                              const float *grad_in_hw3, float *grad_pos, float *grad_nrm, int mem,
                              uint64_t lane_begin, uint64_t lane_end, int film_ready);

/* Adjoint of Mesh::recompute_vertex_normals (src/render/mesh.cpp:311-330).  In the reference params.update() re-derives the
 * vertex normals of a mesh from its (AD-attached) positions inside Mesh::parameters_changed (mesh.cpp:82-87), so the
 * `si_follow.sh_frame.n · g3` seed of epsm.py:644-645 continues from vertex_normals into vertex_positions.  This entry
 * point does that step for every EPSM_MESH_GRAD mesh that has vertex normals: grad_pos += Jᵀ · grad_nrm, both
 * 3·(total vertex count) floats in `mem` space, J evaluated at the scene's current positions.  grad_nrm is not modified
 * (it stays dr.grad of `<mesh>.vertex_normals` for callers that set the normals directly). */
int epsm_vertex_normals_backward(epsm_ctx *ctx, const float *grad_nrm, float *grad_pos, int mem);

/* Per-path records for parity tests (the data the reference logs in epsm.py:547,648-654 and the
 * calc_grad outputs).  One record per lane in [lane_begin, lane_end); layout EPSM_REC_* below.
 * Runs the same kernel as epsm_render_backward with gradient scatter disabled. */
#define EPSM_REC_MAXV      5                  /* logged vertices per path (epsm.py:648)        */
#define EPSM_REC_VSTRIDE   16                 /* floats per vertex record                      */
/* per vertex k: [0] active, [1] prim (uint32 bits), [2..4] p, [5..6] b0,b1, [7] active_em,
 *               [8..10] light point, [11] bsdf flags (uint32 bits), [12] eta, [13..15] reserved */
#define EPSM_REC_GRAD_OFF  (EPSM_REC_MAXV * EPSM_REC_VSTRIDE)
/* then per vertex k: path_grad[5k..5k+4] (15 floats), light_grad[k] (3), diffuse_grad[k] (3) */
#define EPSM_REC_GSTRIDE   21
#define EPSM_REC_HEAD_OFF  (EPSM_REC_GRAD_OFF + EPSM_REC_MAXV * EPSM_REC_GSTRIDE)
/* then: ray o(3) d(3), dlduv(2), dldp1(3), L(3) = 14 floats */
#define EPSM_REC_FLOATS    (EPSM_REC_HEAD_OFF + 14)
int epsm_path_records(epsm_ctx *ctx, int variant, uint32_t sensor, uint32_t seed, uint32_t spp,
                      uint32_t max_depth, uint32_t rr_depth, const float *grad_in_hw5,
                      uint64_t lane_begin, uint64_t lane_end, float *records, int mem);

/* PCG32Sampler::seed + next_float32 (src/render/sampler.cpp:115-134, drjit/random.h:55-132,
 * include/mitsuba/core/random.h:76-90): first `n_floats` of lanes [lane_begin, lane_end). */
int epsm_sampler_stream(epsm_ctx *ctx, uint32_t seed, uint64_t lane_begin, uint64_t lane_end,
                        uint32_t n_floats, float *out, int mem);

/* ADIntegrator.sample_rays + PerspectiveCamera::sample_ray_differential (common.py:291-422,
 * perspective.cpp:238-279): per lane 14 floats o(3) d(3) d_x(3) d_y(3) pos(2). */
int epsm_sample_rays(epsm_ctx *ctx, uint32_t sensor, uint32_t seed, uint32_t spp,
                     uint64_t lane_begin, uint64_t lane_end, float *out, int mem);

/* Multi-GPU sharding by sample index: restrict render / backward launches of this context to the samples
 * s ∈ [s_begin, s_end) of every pixel (lane = pixel·spp + s, common.py:326-343).  Every rank then sees every
 * pixel, so the load is balanced, and the union over ranks is exactly the single-GPU wavefront because a
 * lane's random stream depends only on (seed, lane) (sampler.cpp:122-130).  (0, 0) = all samples (default). */
int epsm_set_sample_range(epsm_ctx *ctx, uint32_t s_begin, uint32_t s_end);

/* Gradient liveness pruning.  By default epsm_render_backward stops the replay of each path at the vertex where its
 * gradient chain ends: calc_grad masks every output behind the first diffuse / non-mesh / inactive vertex (`manifold`,
 * epsm.py:793-796,855,862,920,928) or behind the second diffuse one (`manifold_caustic`, epsm.py:999,1008,1169-1182), and
 * render_backward consumes nothing else of the path (its radiance δL is never propagated, epsm.py:728-732).  The
 * accumulated gradients are those of the full replay (tests/ hold the two to each other and to the oracle).
 * enable != 0 runs the full min(max_depth, 6)-iteration replay of sample_path (epsm.py:503-742) for every lane, as the
 * reference does; epsm_path_records always does.  Environment override at context creation: EPSM_FULL_REPLAY=1. */
int epsm_set_full_replay(epsm_ctx *ctx, int enable);

/* The matcher that produces grad_in: Matcher.match_Sinkhorn (EPSM/utils/matcher.py:51-63) =
 * geomloss.SamplesLoss("sinkhorn", blur, scaling) (p = 2, debias, balanced, uniform weights) between the point clouds
 * x (n × 5) and y (m × 5), differentiated w.r.t. x:  grad_x = scale · ∂loss/∂x  (the reference passes scale = res²,
 * matcher.py:60).  geomloss is an unpinned third-party dependency absent from the reference tree; the algorithm is the
 * published sinkhorn_loop of geomloss 0.2.x (ε-scaling from the clouds' diameter down to blur², symmetric averaged
 * updates, one un-averaged extrapolation, debiased cost) — see oracle/sinkhorn_ref.py.  No scene is needed.
 * grad_x: n × 5 floats in `mem` space (may be NULL); loss (HOST pointer, may be NULL): the Sinkhorn divergence;
 * n_steps (HOST, may be NULL): length of the ε-schedule.  With device pointers and loss == NULL the call does not
 * wait for the result (one host read-back of the 10-float bounding box happens inside). */
int epsm_sinkhorn_grad(epsm_ctx *ctx, const float *x, uint32_t n, const float *y, uint32_t m,
                       float blur, float scaling, float scale, float *grad_x, double *loss, uint32_t *n_steps, int mem);

/* Scene statistics for the roofline accounting: bytes per BVH node / triangle record, counts. */
typedef struct epsm_scene_stats {
    uint64_t n_triangles, n_vertices, n_nodes;
    uint32_t node_bytes, tri_bytes;
    uint32_t bvh_width, max_leaf;
    double   build_ms;
} epsm_scene_stats;
int epsm_scene_get_stats(epsm_ctx *ctx, epsm_scene_stats *out);

/* Device-side timing of the traversal kernel, for the roofline line of bench.py: when enabled, CUDA events on the
 * context's stream bracket every traversal launch of the following gradient passes.  epsm_get_trace_profile
 * synchronises, returns the accumulated milliseconds / launch count since the last call and resets them. */
int epsm_set_trace_profiling(epsm_ctx *ctx, int enable);
int epsm_get_trace_profile(epsm_ctx *ctx, double *ms, uint64_t *launches);

/* Number of kernels this library launched on the context since creation (bench `gpu_launches`). */
uint64_t epsm_launch_count(const epsm_ctx *ctx);
/* The context's stream as a cudaStream_t, for event timing on the launching stream. */
void *epsm_stream(epsm_ctx *ctx);
int epsm_synchronize(epsm_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* EPSM_B200_H */
