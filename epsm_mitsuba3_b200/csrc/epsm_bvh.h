// epsm_bvh.h — host builder of the 4-wide BVH consumed by bvh_trace (epsm_device.cuh).
// Replaces the reference's Embree build (src/render/scene_embree.inl:83-187) and OptiX GAS build
// (include/mitsuba/render/optix/shapes.h:75-170): B200 has no RT cores, so the structure is ours.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>
#include "epsm_types.h"

struct Bvh4 {
    std::vector<f4> nodes;   // EPSM_NODE_F4 per node, root = node 0
    std::vector<f4> tris;    // EPSM_TRI_F4 per triangle, leaf order
    uint32_t depth = 0;
    double build_ms = 0;
    float pad = 0;           // box padding used by the build (1e-5·max|coordinate| + 1e-7)
    double cost = 0;         // epsm_bvh4_cost of the built tree
};
// Node pad word (float4 [7], never read by the traversal): .x bits = (parent node << 2) | slot (0xffffffff: root),
// .y bits = number of inner children — the links the device refit (k_refit_nodes, epsm_kernels.cu) climbs.
double epsm_bvh4_cost(const f4 *nodes, size_t n_nodes);
// pos: 3·V floats; faces: 3·F global vertex ids.  Binned-SAH binary build (multi-threaded), collapsed to 4-wide;
// child boxes are padded so that every triangle the exact Möller–Trumbore test accepts is reached.
void epsm_build_bvh4(const float *pos, const uint32_t *faces, size_t n_faces, Bvh4 &out);
