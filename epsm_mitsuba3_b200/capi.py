"""ctypes mirror of include/epsm_b200.h (struct layouts and constants) and the loader of
libepsm_b200.so.  Interface definitions only: no arithmetic lives here.

The product path fails loudly when the CUDA library is missing or no GPU is usable — there is
no CPU fallback (north_star)."""
from __future__ import annotations

import ctypes as C
import os

MEM_HOST, MEM_DEVICE = 0, 1
BSDF_DIFFUSE, BSDF_DIELECTRIC, BSDF_CONDUCTOR, BSDF_ROUGHCONDUCTOR = 0, 1, 2, 3
BSDF_TWOSIDED, BSDF_GGX, BSDF_SAMPLE_VISIBLE = 1, 2, 4
MESH_ISMESH, MESH_GRAD, MESH_FLIP_NORMALS, MESH_RECTANGLE = 1, 2, 4, 8
VARIANT_MANIFOLD, VARIANT_MANIFOLD_CAUSTIC = 0, 1
BF_NULL, BF_DIFFUSE, BF_SMOOTH, BF_DELTA = 0x1, 0x6, 0x1e, 0x61

REC_MAXV, REC_VSTRIDE, REC_GSTRIDE = 5, 16, 21
REC_GRAD_OFF = REC_MAXV * REC_VSTRIDE
REC_HEAD_OFF = REC_GRAD_OFF + REC_MAXV * REC_GSTRIDE
REC_FLOATS = REC_HEAD_OFF + 14


class MeshDesc(C.Structure):
    _fields_ = [("vertex_count", C.c_uint32), ("face_count", C.c_uint32),
                ("positions", C.POINTER(C.c_float)), ("faces", C.POINTER(C.c_uint32)),
                ("normals", C.POINTER(C.c_float)),
                ("bsdf_id", C.c_int32), ("emitter_id", C.c_int32), ("flags", C.c_uint32),
                ("texcoords", C.POINTER(C.c_float))]


TEX_NEAREST, TEX_BILINEAR = 0, 1
TEX_REPEAT, TEX_CLAMP, TEX_MIRROR = 0, 1, 2


class TextureDesc(C.Structure):
    _fields_ = [("texels", C.POINTER(C.c_float)), ("width", C.c_uint32), ("height", C.c_uint32),
                ("filter", C.c_uint32), ("wrap", C.c_uint32), ("to_uv", C.c_float * 6)]


class BsdfDesc(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("flags", C.c_uint32),
                ("reflectance", C.c_float * 3), ("transmittance", C.c_float * 3),
                ("eta", C.c_float * 3), ("k", C.c_float * 3),
                ("alpha_u", C.c_float), ("alpha_v", C.c_float),
                ("reflectance_texture", C.POINTER(TextureDesc))]


class EmitterDesc(C.Structure):
    _fields_ = [("radiance", C.c_float * 3), ("mesh_id", C.c_int32)]


class SensorDesc(C.Structure):
    _fields_ = [("to_world", C.c_float * 16), ("fov_x", C.c_float),
                ("near_clip", C.c_float), ("far_clip", C.c_float),
                ("width", C.c_uint32), ("height", C.c_uint32), ("flags", C.c_uint32)]


class ReparamConfig(C.Structure):
    _fields_ = [("max_depth", C.c_uint32), ("rr_depth", C.c_uint32), ("reparam_max_depth", C.c_uint32), ("reparam_rays", C.c_uint32),
                ("reparam_kappa", C.c_float), ("reparam_exp", C.c_float)]


class SceneStats(C.Structure):
    _fields_ = [("n_triangles", C.c_uint64), ("n_vertices", C.c_uint64), ("n_nodes", C.c_uint64),
                ("node_bytes", C.c_uint32), ("tri_bytes", C.c_uint32),
                ("bvh_width", C.c_uint32), ("max_leaf", C.c_uint32), ("build_ms", C.c_double)]


# Every symbol include/epsm_b200.h declares (tests/test_capi.py checks the library exports all of them).
SENSOR_SAMPLE_BORDER = 1

EXPORTS = [
    "epsm_create", "epsm_destroy", "epsm_last_error", "epsm_scene_upload",
    "epsm_scene_update_vertices", "epsm_scene_update_normals", "epsm_scene_update_sensor", "epsm_scene_set_mesh_grad", "epsm_prb_reparam_backward", "epsm_scene_vertex_buffers", "epsm_scene_commit", "epsm_scene_refit", "epsm_scene_accel_cost", "epsm_scene_accel_raw", "epsm_trace",
    "epsm_render_accumulate", "epsm_film_clear", "epsm_film_raw", "epsm_film_develop", "epsm_render",
    "epsm_render_backward", "epsm_vertex_normals_backward", "epsm_path_records", "epsm_sampler_stream", "epsm_sample_rays",
    "epsm_set_sample_range", "epsm_set_full_replay", "epsm_set_trace_profiling", "epsm_get_trace_profile", "epsm_scene_get_stats", "epsm_launch_count", "epsm_stream", "epsm_synchronize",
    "epsm_sinkhorn_grad",
]

_LIB = None


def library_path() -> str:
    if os.environ.get("EPSM_B200_LIB"):                       # tuning variants of the same library (experiments only)
        return os.environ["EPSM_B200_LIB"]
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libepsm_b200.so")


def load_library() -> C.CDLL:
    """dlopen libepsm_b200.so and declare the prototypes.  Raises if the library has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  epsm_mitsuba3_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    fp = C.c_void_p  # float* in either memory space: passed as raw addresses
    lib.epsm_create.argtypes = [i32, C.POINTER(vp)]
    lib.epsm_destroy.argtypes = [vp]
    lib.epsm_last_error.argtypes = [vp]
    lib.epsm_last_error.restype = C.c_char_p
    lib.epsm_scene_upload.argtypes = [vp, C.POINTER(MeshDesc), u32, C.POINTER(BsdfDesc), u32,
                                      C.POINTER(EmitterDesc), u32, C.POINTER(SensorDesc), u32]
    lib.epsm_scene_update_vertices.argtypes = [vp, u32, fp, i32]
    lib.epsm_scene_update_normals.argtypes = [vp, u32, fp, i32]
    lib.epsm_prb_reparam_backward.argtypes = [vp, u32, u32, u32, C.POINTER(ReparamConfig), fp, fp, fp, i32, u64, u64, i32]
    lib.epsm_scene_set_mesh_grad.argtypes = [vp, u32, i32]
    lib.epsm_scene_update_sensor.argtypes = [vp, u32, C.POINTER(SensorDesc)]
    lib.epsm_scene_vertex_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)]
    lib.epsm_scene_commit.argtypes = [vp]
    lib.epsm_scene_refit.argtypes = [vp]
    lib.epsm_scene_accel_cost.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(u64)]
    lib.epsm_scene_accel_raw.argtypes = [vp, C.POINTER(vp), C.POINTER(u64), C.POINTER(vp), C.POINTER(u64)]
    lib.epsm_trace.argtypes = [vp, fp, u64, vp, i32, vp, i32]
    lib.epsm_render_accumulate.argtypes = [vp, u32, u32, u32, u32, u32, u64, u64, vp]
    lib.epsm_film_clear.argtypes = [vp, u32]
    lib.epsm_film_raw.argtypes = [vp, u32, C.POINTER(vp), C.POINTER(u64)]
    lib.epsm_film_develop.argtypes = [vp, u32, fp, i32]
    lib.epsm_render.argtypes = [vp, u32, u32, u32, u32, u32, fp, i32]
    lib.epsm_render_backward.argtypes = [vp, i32, u32, u32, u32, u32, u32, fp, fp, fp, i32, u64, u64, vp]
    lib.epsm_vertex_normals_backward.argtypes = [vp, fp, fp, i32]
    lib.epsm_path_records.argtypes = [vp, i32, u32, u32, u32, u32, u32, fp, u64, u64, fp, i32]
    lib.epsm_sampler_stream.argtypes = [vp, u32, u64, u64, u32, fp, i32]
    lib.epsm_sample_rays.argtypes = [vp, u32, u32, u32, u64, u64, fp, i32]
    lib.epsm_set_sample_range.argtypes = [vp, u32, u32]
    lib.epsm_set_full_replay.argtypes = [vp, i32]
    lib.epsm_set_trace_profiling.argtypes = [vp, i32]
    lib.epsm_get_trace_profile.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(u64)]
    lib.epsm_scene_get_stats.argtypes = [vp, C.POINTER(SceneStats)]
    lib.epsm_launch_count.argtypes = [vp]
    lib.epsm_launch_count.restype = u64
    lib.epsm_stream.argtypes = [vp]
    lib.epsm_stream.restype = vp
    lib.epsm_synchronize.argtypes = [vp]
    lib.epsm_sinkhorn_grad.argtypes = [vp, fp, u32, fp, u32, C.c_float, C.c_float, C.c_float, fp, C.POINTER(C.c_double), C.POINTER(u32), i32]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("epsm_last_error", "epsm_launch_count", "epsm_stream"):
            fn.restype = C.c_int
    _LIB = lib
    return lib
