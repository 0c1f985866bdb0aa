// epsm_launch.h — host-side launch wrappers implemented in epsm_kernels.cu (internal, not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include "epsm_types.h"

cudaError_t epsm_launch_trace(const DScene &sc, const float *rays, uint64_t n, void *hits, int any_hit, unsigned long long *counters, cudaStream_t s);
cudaError_t epsm_launch_render(const DScene &sc, DLaunch ln, float *film, unsigned long long *counters, cudaStream_t s);
cudaError_t epsm_launch_develop(const float *film, uint64_t n_pixels, float *out, cudaStream_t s);
// Wavefront gradient pass.  `workspace` holds epsm_wf_workspace_bytes(capacity) bytes of device memory for one chunk of
// `capacity` lanes; `trace_grid` = CTAs of the persistent traversal kernel (SM count × resident CTAs).
#include <vector>
#include <utility>
typedef std::vector<std::pair<cudaEvent_t, cudaEvent_t>> EpsmEventList;   // optional timing of the traversal launches
size_t epsm_wf_workspace_bytes(uint32_t capacity);
cudaError_t epsm_launch_backward(const DScene &sc, DLaunch ln, const float *grad_in, float *grad_pos, float *grad_nrm, float *records,
                                 unsigned long long *counters, void *workspace, uint32_t capacity, uint32_t trace_grid, uint64_t *n_launches,
                                 EpsmEventList *trace_events, cudaStream_t s);
// Wavefront primal render (same workspace); k_render (epsm_launch_render) is the single-kernel variant kept for comparison.
cudaError_t epsm_launch_render_wf(const DScene &sc, DLaunch ln, float *film, unsigned long long *counters, void *workspace, uint32_t capacity,
                                  uint32_t trace_grid, uint64_t *n_launches, cudaStream_t s, float *lane_L = nullptr);
cudaError_t epsm_launch_sampler_stream(uint32_t seed, uint64_t lane_begin, uint64_t n_lanes, uint32_t n_floats, float *out, cudaStream_t s);
cudaError_t epsm_launch_sample_rays(const DSensor &se, uint32_t seed, uint32_t spp, uint64_t lane_begin, uint64_t n_lanes, float *out, cudaStream_t s);
cudaError_t epsm_launch_vertex_normals(const float *pos, const uint32_t *faces, uint32_t f_off, uint32_t f_cnt, uint32_t v_off, uint32_t v_cnt,
                                       float *nrm, cudaStream_t s);
// grad_pos += Jᵀ·grad_nrm for the vertex range: adjoint of epsm_launch_vertex_normals.  `scratch`: 3·(total vertices) floats.
cudaError_t epsm_launch_vertex_normals_backward(const float *pos, const uint32_t *faces, uint32_t f_off, uint32_t f_cnt, uint32_t v_off, uint32_t v_cnt,
                                                const float *grad_nrm, float *grad_pos, float *scratch, cudaStream_t s);
// Device refit of the BVH4 (topology of the last host build kept): rewrites triangle records and all boxes from `pos`.
// `scratch` holds epsm_refit_scratch_words(n_nodes) uint32; after the launch scratch[0] = bits of max|coordinate| and the
// double at scratch + 2 = epsm_bvh4_cost of the refitted tree.
size_t epsm_refit_scratch_words(uint32_t n_nodes);
cudaError_t epsm_launch_refit(f4 *nodes, uint32_t n_nodes, f4 *tris, uint32_t n_tris, const uint32_t *faces, const float *pos,
                              uint32_t *scratch, cudaStream_t s);
// prb_reparam reverse pass (k_prb_backward): `film` = RGBW film of the primal pass over all lanes, grad_in (H,W,3).
cudaError_t epsm_launch_prb_backward(const DScene &sc, DLaunch ln, const PrbConfig &cfg, const float *film, const float *grad_in,
                                     float *grad_pos, float *grad_nrm, cudaStream_t s);
// The same pass as a replay wavefront over the persistent traversal kernel (max_depth ≤ EPSM_PRB_WF_MAX_DEPTH).  lane_L: per-lane radiance of
// the primal pass (epsm_launch_render_wf's lane_L output), indexed by global lane id.
#define EPSM_PRB_WF_MAX_DEPTH 16
size_t epsm_prb_workspace_bytes(uint32_t capacity, uint32_t K, uint32_t n_aux);
cudaError_t epsm_launch_prb_backward_wf(const DScene &sc, DLaunch ln, const PrbConfig &cfg, const float *film, const float *lane_L, const float *grad_in,
                                        float *grad_pos, float *grad_nrm, void *workspace, uint32_t capacity, uint32_t trace_grid, uint64_t *n_launches, cudaStream_t s);
