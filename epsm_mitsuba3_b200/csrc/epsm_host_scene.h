// epsm_host_scene.h — host-side scene preparation shared by the C ABI (epsm_api.cu) and the test-only
// host harness (tests/hostsim): concatenated mesh arrays, BSDF flag words, sensor transforms, the
// reconstruction-filter polynomial, emitter area tables and the BVH4.
#pragma once
#include <string>
#include <vector>
#include "../../include/epsm_b200.h"
#include "epsm_types.h"
#include "epsm_bvh.h"

struct HostScene {
    std::vector<float> pos, nrm, cdf, pmf, uv, texels;
    std::vector<DTexture> textures;
    std::vector<uint32_t> faces, face_mesh;
    std::vector<DMesh> meshes; std::vector<DBsdf> bsdfs; std::vector<DEmitter> emitters; std::vector<DSensor> sensors;
    Bvh4 bvh;
    float filter[10];
};
// Returns an empty string on success, else the error message.
std::string host_scene_build(HostScene &hs, const epsm_mesh_desc *meshes, uint32_t n_meshes, const epsm_bsdf_desc *bsdfs, uint32_t n_bsdfs,
                             const epsm_emitter_desc *emitters, uint32_t n_emitters, const epsm_sensor_desc *sensors, uint32_t n_sensors);
// (Re)builds the BVH4 and the emitter area tables from hs.pos.
void host_scene_rebuild_accel(HostScene &hs);
// Only the emitter area tables / per-mesh area fields (Mesh::build_pmf) from hs.pos.
void host_scene_rebuild_area_tables(HostScene &hs);
// params['<sensor>.to_world'] = …; params.update(): recomputes the camera transforms of sensor `index` (same film size).
std::string host_scene_set_sensor(HostScene &hs, uint32_t index, const epsm_sensor_desc &d);
// DScene over the HOST arrays (hostsim only; the product fills a DScene with device pointers).
void host_scene_view(const HostScene &hs, DScene &sc);
