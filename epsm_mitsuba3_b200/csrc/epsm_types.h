// epsm_types.h — device-resident scene layout shared by the host builder, the kernels and the C ABI.
//
// HBM layout:
//   nodes      BVH4, 128 B per node, SoA inside the node: 8 × float4
//                [0] lo.x[4] [1] lo.y[4] [2] lo.z[4] [3] hi.x[4] [4] hi.y[4] [5] hi.z[4] [6] child[4] [7] pad
//              child >= 0: inner node index; child < 0: leaf, ~child = (first_tri << EPSM_LEAF_BITS) | (count-1);
//              empty slots have child == 0x7fffffff and are skipped.  One node = one 128-byte line, fetched
//              as three 256-bit loads (sm_100a LDG.E.256) for the boxes + one 128-bit load for the children.
//   tris       64-byte slots (32-byte aligned), 48 B of payload per triangle in BVH leaf order: 4 × float4
//                [0] p0.xyz, prim_index (bits)  [1] e1.xyz, e2.x  [2] e2.yz, 0, 0  [3] alignment pad (never read)
//              fetched as one 256-bit load (LDG.E.256: [0],[1]) + one 64-bit load (e2.yz)
//              e1 = p1 - p0 and e2 = p2 - p0 are the fp32 differences the reference's Möller–Trumbore
//              computes per query (include/mitsuba/render/mesh.h:347), so results are bit-identical.
//   pos/nrm    float[3·V] concatenated over meshes (vertex gathers on a hit: 84 B with faces)
//   faces      uint32[3·F] GLOBAL vertex ids;  face_mesh uint32[F]
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define EPSM_HD __host__ __device__ __forceinline__
#define EPSM_HDN __host__ __device__
#else
#define EPSM_HD inline
#define EPSM_HDN
#endif

struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };

#define EPSM_NODE_F4   8      // float4 per BVH node
#define EPSM_TRI_F4    4      // float4 per triangle slot (3 used: EPSM_TRI_BYTES of payload)
#define EPSM_TRI_BYTES 48     // algorithmic bytes per triangle test
#ifndef EPSM_LEAF_BITS
#define EPSM_LEAF_BITS 2      // bits of a leaf reference that hold (count − 1)
#endif
#define EPSM_MAX_LEAF  (1 << EPSM_LEAF_BITS)
#define EPSM_MAXV      5      // logged vertices per path (epsm.py:648)
#define EPSM_MAX_ITER  6      // sample_path iteration cap (epsm.py:549)

struct DMesh {
    uint32_t v_off, f_off, v_cnt, f_cnt;
    int32_t  bsdf, emitter;
    uint32_t flags, has_normals;
    uint32_t cdf_off;            // into DScene::cdf / pmf (only meaningful for emitter meshes)
    uint32_t valid_lo, valid_hi;
    float    area_sum, area_norm;
    uint32_t has_uv;             // vertex texture coordinates present (DScene::uv)
    float    _pad[2];
};

struct DBsdf {
    uint32_t kind, flags, bflags; int32_t tex;   // bflags = Mitsuba BSDFFlags word; tex: reflectance texture (DScene::textures) or -1
    float reflectance[3], transmittance[3], eta[3], k[3];
    float alpha_u, alpha_v;
    float _pad2[2];
};

struct DEmitter { float radiance[3]; int32_t mesh; };

// Bitmap reflectance (src/textures/bitmap.cpp over Dr.Jit's Texture): linear RGB texels at texels[off + 3·(y·w + x)].
struct DTexture { uint32_t off, w, h, filter, wrap; float to_uv[6]; uint32_t _pad; };

struct DSensor {
    float s2c[16], to_world[16];
    float c2s[16], w2c[16];      // camera_to_sample and to_world⁻¹ (Sensor::sample_direction, perspective.cpp:282-306): prb_reparam's film motion
    f3 dx, dy;
    float near_clip, far_clip;
    uint32_t w, h;               // sampling grid = film + 2·border: what lane → pixel enumerates (common.py:318-343)
    uint32_t fw, fh;             // film (= crop) size
    int32_t  border;             // rfilter border when hdrfilm.sample_border is set (2 for the Gaussian), else 0
};

struct DScene {
    const f4 *nodes; const f4 *tris;
    const float *pos; const float *nrm;
    const uint32_t *faces; const uint32_t *face_mesh;
    const DMesh *meshes; const DBsdf *bsdfs; const DEmitter *emitters;
    const float *cdf; const float *pmf;
    const float *uv;             // float[2·V] vertex texture coordinates (zeros for meshes without)
    const DTexture *textures; const float *texels;
    uint32_t n_emitters, n_meshes, n_tris, n_nodes;
    float filter[10];            // scaled Remez coefficients of the Gaussian rfilter (gaussian.cpp:57-89)
};

// Work description of one render / backward launch.
struct DLaunch {
    DSensor sensor;
    uint32_t seed, spp, max_depth, rr_depth;
    uint64_t lane_begin, lane_end;
    uint32_t tile_y0, tile_y1;   // tile rows covered by this launch
    uint32_t s_begin, s_end;     // sample slots of each pixel handled by this launch (multi-GPU sharding)
    int variant;
    uint32_t prune;              // gradient pass: stop each path where its gradient chain ends (epsm_device.cuh: wf_chain_vertex)
    uint32_t pixel_major;        // work-item order inside a tile: 1 = consecutive samples of one pixel (gradient pass), 0 = the tile's 32 pixels at one sample index (primal render)
};

// prb_reparam properties (prb_reparam.py:225-247) as the kernels take them.
#define EPSM_PRB_MAX_RAYS 64
struct PrbConfig { uint32_t max_depth, rr_depth, reparam_max_depth, n_rays; float kappa, exponent; };
