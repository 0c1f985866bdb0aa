"""Python front-end of the CPU oracle (libepsm_oracle.so) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  It restates the reference's algorithm (see epsm_oracle.cpp for the file:line map) and
is the checker the CUDA path is compared against; it is never part of the product path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from epsm_mitsuba3_b200 import capi            # struct layouts of include/epsm_b200.h (interface only)
from epsm_mitsuba3_b200.scenes import SceneData

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Vertex(C.Structure):
    _fields_ = [("active", C.c_float), ("ismesh", C.c_float), ("active_em", C.c_float), ("eta", C.c_float),
                ("bsdf_flags", C.c_uint32), ("prim", C.c_uint32), ("fi", C.c_uint32 * 3),
                ("light", C.c_float * 3), ("p0", C.c_float * 3), ("p1", C.c_float * 3), ("p2", C.c_float * 3),
                ("p", C.c_float * 3), ("b0", C.c_float), ("b1", C.c_float), ("sh_n", C.c_float * 3),
                ("n0", C.c_float * 3), ("n1", C.c_float * 3), ("n2", C.c_float * 3), ("hf", C.c_float * 3),
                ("bw", C.c_float * 3), ("lr_dir_sum", C.c_float),
                ("direct_valid", C.c_uint32), ("direct_fi", C.c_uint32 * 3), ("direct_bw", C.c_float * 3),
                ("direct_dis", C.c_float), ("direct_any", C.c_uint32),
                ("mesh_flags", C.c_uint32), ("has_normals", C.c_uint32)]


class Path(C.Structure):
    _fields_ = [("cam", C.c_float * 3), ("d", C.c_float * 3), ("dlduv", C.c_float * 2), ("dldp1", C.c_float * 3),
                ("L", C.c_float * 3), ("n_vertices", C.c_uint32), ("v", Vertex * 5)]


def build(force: bool = False) -> str:
    """Compile libepsm_oracle.so with the committed Makefile (g++ -ffp-contract=off)."""
    so = os.path.join(_HERE, "libepsm_oracle.so")
    src = os.path.join(_HERE, "epsm_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libepsm_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_scene_create.restype = C.c_void_p
        L.orc_scene_create.argtypes = [C.POINTER(capi.MeshDesc), C.c_uint32, C.POINTER(capi.BsdfDesc), C.c_uint32,
                                       C.POINTER(capi.EmitterDesc), C.c_uint32, C.POINTER(capi.SensorDesc), C.c_uint32]
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_scene_update_vertices.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_scene_set_accel.argtypes = [C.c_void_p, C.c_int]
        L.orc_tea32.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
        L.orc_tea_float32.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
        L.orc_tea_float32.restype = C.c_float
        L.orc_pcg32_stream.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
        L.orc_sampler_stream.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
        L.orc_sincos.argtypes = [C.c_float, C.c_void_p, C.c_void_p]
        L.orc_filter_eval.argtypes = [C.c_void_p, C.c_float]
        L.orc_filter_eval.restype = C.c_float
        L.orc_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
        L.orc_surface_interaction.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_surface_uv.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_texture_eval.argtypes = [C.POINTER(capi.TextureDesc), C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_bsdf_sample.argtypes = [C.POINTER(capi.BsdfDesc), C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.orc_bsdf_eval_pdf.argtypes = [C.POINTER(capi.BsdfDesc), C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_sample_emitter.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
        L.orc_bsdf_flags.argtypes = [C.POINTER(capi.BsdfDesc)]; L.orc_bsdf_flags.restype = C.c_uint32
        L.orc_pdf_emitter.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_sample_rays.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_camera_ray.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_void_p]
        L.orc_render_accumulate.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                            C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.orc_film_develop.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_film_put.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_void_p]
        L.orc_path_log.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                   C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_sizeof_path.restype = C.c_uint32
        L.orc_sizeof_vertex.restype = C.c_uint32
        assert L.orc_sizeof_path() == C.sizeof(Path), (L.orc_sizeof_path(), C.sizeof(Path))
        _LIB = L
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def tea32(v0, v1, rounds=4):
    out = np.zeros(2, np.uint32); lib().orc_tea32(v0, v1, rounds, _p(out)); return int(out[0]), int(out[1])


def tea_float32(v0, v1, rounds=4):
    return float(lib().orc_tea_float32(v0, v1, rounds))


def pcg32_stream(initstate, initseq, n):
    out = np.zeros(n, np.float32); lib().orc_pcg32_stream(initstate, initseq, n, _p(out)); return out


def sampler_stream(seed, lane_begin, lane_end, n):
    out = np.zeros((lane_end - lane_begin, n), np.float32)
    lib().orc_sampler_stream(seed, lane_begin, lane_end, n, _p(out)); return out


def sincos(x):
    s, c = C.c_float(), C.c_float(); lib().orc_sincos(x, C.byref(s), C.byref(c)); return s.value, c.value


def bsdf_sample(bsdf, wi, s1, s2):
    d = _bsdf_desc(bsdf); wi = np.asarray(wi, np.float32); out = np.zeros(12, np.float32)
    lib().orc_bsdf_sample(C.byref(d), _p(wi), s1, s2[0], s2[1], _p(out))
    return dict(wo=out[0:3], pdf=out[3], eta=out[4], type=int(out[5]), weight=out[6:9], hf=out[9:12])


def bsdf_eval_pdf(bsdf, wi, wo):
    d = _bsdf_desc(bsdf); wi = np.asarray(wi, np.float32); wo = np.asarray(wo, np.float32); out = np.zeros(4, np.float32)
    lib().orc_bsdf_eval_pdf(C.byref(d), _p(wi), _p(wo), _p(out)); return out[:3], out[3]


def texture_eval(tex, uv):
    """Oracle lookup of a scenes.Texture at (n,2) uv coordinates."""
    t = tex.to_c(); uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2); out = np.zeros((uv.shape[0], 3), np.float32)
    lib().orc_texture_eval(C.byref(t), _p(uv), uv.shape[0], _p(out)); return out


def bsdf_flags(bsdf):
    d = _bsdf_desc(bsdf); return int(lib().orc_bsdf_flags(C.byref(d)))


def _bsdf_desc(b):
    d = capi.BsdfDesc(); d.kind, d.flags = b.kind, b.flags_word()
    d.reflectance[:] = [float(x) for x in b.reflectance]; d.transmittance[:] = [float(x) for x in b.transmittance]
    d.eta[:] = [float(x) for x in b.eta]; d.k[:] = [float(x) for x in b.k]; d.alpha_u, d.alpha_v = b.alpha_u, b.alpha_v
    return d


class OracleScene:
    """CPU restatement of Scene + EPSMIntegrator on a SceneData."""

    def __init__(self, sd: SceneData, accel: bool = False):
        self.sd = sd
        self._c = sd.to_c()
        nm, nb, ne, ns = self._c.counts
        self.h = lib().orc_scene_create(self._c.meshes, nm, self._c.bsdfs, nb, self._c.emitters, ne, self._c.sensors, ns)
        if accel:
            lib().orc_scene_set_accel(self.h, 1)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().orc_scene_destroy(self.h); self.h = None
        except Exception:       # interpreter shutdown
            pass

    def set_accel(self, on: bool):
        lib().orc_scene_set_accel(self.h, 1 if on else 0)

    def update_vertices(self, mesh_id, positions):
        p = np.ascontiguousarray(positions, np.float32)
        self.sd.meshes[mesh_id].positions = p.reshape(-1, 3)
        lib().orc_scene_update_vertices(self.h, mesh_id, _p(p))

    def trace(self, rays: np.ndarray, any_hit=False):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        if any_hit:
            out = np.zeros(n, np.uint32); lib().orc_trace(self.h, _p(rays), n, _p(out), 1); return out
        out = np.zeros((n, 4), np.float32); lib().orc_trace(self.h, _p(rays), n, _p(out), 0)
        return out

    def surface_interaction(self, o, d, maxt=np.float32(3.4e38)):
        r = np.array([o[0], o[1], o[2], maxt, d[0], d[1], d[2], 0], np.float32); out = np.zeros(40, np.float32)
        lib().orc_surface_interaction(self.h, _p(r), _p(out))
        return dict(t=out[0], p=out[1:4], n=out[4:7], sh_n=out[7:10], wi=out[10:13], b0=out[13], b1=out[14],
                    ismesh=int(out[15]), p0=out[16:19], p1=out[19:22], p2=out[22:25], n0=out[25:28], n1=out[28:31],
                    n2=out[31:34], prim=int(out[34]) if np.isfinite(out[0]) else -1, bsdf_flags=int(out[35]), uv=out[36:38])

    def surface_uv(self, o, d, maxt=np.float32(3.4e38)):
        r = np.array([o[0], o[1], o[2], maxt, d[0], d[1], d[2], 0], np.float32); out = np.zeros(17, np.float32)
        lib().orc_surface_uv(self.h, _p(r), _p(out))
        return dict(uv=out[0:2], sh_s=out[2:5], sh_t=out[5:8], reflectance=out[8:11], dp_du=out[11:14], dp_dv=out[14:17])

    def sample_emitter(self, o, d, sample):
        r = np.array([o[0], o[1], o[2], 3.4e38, d[0], d[1], d[2], 0], np.float32); out = np.zeros(14, np.float32)
        lib().orc_sample_emitter(self.h, _p(r), sample[0], sample[1], _p(out))
        return dict(p=out[0:3], n=out[3:6], d=out[6:9], pdf=out[9], dist=out[10], weight=out[11:14])

    def pdf_emitter(self, ref_p, d):
        """(pdf_emitter_direction, Emitter::eval) for the surface seen from ref_p along d."""
        r = np.array([ref_p[0], ref_p[1], ref_p[2], 3.4e38, d[0], d[1], d[2], 0], np.float32); out = np.zeros(4, np.float32)
        lib().orc_pdf_emitter(self.h, _p(r), _p(out))
        return float(out[0]), out[1:4]

    def camera_ray(self, sensor, u, v):
        out = np.zeros(13, np.float32); lib().orc_camera_ray(self.h, sensor, u, v, _p(out))
        return dict(o=out[0:3], d=out[3:6], d_x=out[6:9], d_y=out[9:12], maxt=out[12])

    def sample_rays(self, sensor, seed, spp, lane_begin=0, lane_end=None):
        s = self.sd.sensors[sensor]; lane_end = s.wavefront(spp) if lane_end is None else lane_end
        out = np.zeros((lane_end - lane_begin, 14), np.float32)
        lib().orc_sample_rays(self.h, sensor, seed, spp, lane_begin, lane_end, _p(out)); return out

    def filter_eval(self, x):
        return float(lib().orc_filter_eval(self.h, x))

    def film_put(self, film, pos, rgb):
        H, W = film.shape[:2]; rgb = np.asarray(rgb, np.float32)
        lib().orc_film_put(self.h, _p(film), W, H, pos[0], pos[1], _p(rgb))

    def render_raw(self, sensor, seed, spp, max_depth=6, rr_depth=5, lane_begin=0, lane_end=None, per_lane=False):
        s = self.sd.sensors[sensor]; lane_end = s.wavefront(spp) if lane_end is None else lane_end
        film = np.zeros((s.height, s.width, 4), np.float32)
        L = np.zeros((lane_end - lane_begin, 3), np.float32) if per_lane else None
        lib().orc_render_accumulate(self.h, sensor, seed, spp, max_depth, rr_depth, lane_begin, lane_end, _p(film),
                                    _p(L) if per_lane else None)
        return (film, L) if per_lane else film

    @staticmethod
    def develop(film):
        H, W = film.shape[:2]; out = np.zeros((H, W, 5), np.float32)
        lib().orc_film_develop(_p(np.ascontiguousarray(film)), H * W, _p(out)); return out

    def render(self, sensor, seed, spp, max_depth=6, rr_depth=5):
        """EPSMIntegrator.render (epsm.py:13-82): (H,W,5) float32."""
        return self.develop(self.render_raw(sensor, seed, spp, max_depth, rr_depth))

    def path_log(self, sensor, seed, spp, max_depth, rr_depth, grad_in, lane_begin=0, lane_end=None):
        s = self.sd.sensors[sensor]; lane_end = s.wavefront(spp) if lane_end is None else lane_end
        g = np.ascontiguousarray(grad_in, np.float32)
        assert g.shape == (s.height, s.width, 5)
        n = lane_end - lane_begin
        buf = (Path * n)()
        lib().orc_path_log(self.h, sensor, seed, spp, max_depth, rr_depth, _p(g), lane_begin, lane_end, C.byref(buf))
        arr = np.frombuffer(buf, dtype=path_dtype()).copy()
        return arr


def path_dtype():
    f3, u3 = (np.float32, 3), (np.uint32, 3)
    v = np.dtype([("active", "f4"), ("ismesh", "f4"), ("active_em", "f4"), ("eta", "f4"), ("bsdf_flags", "u4"),
                  ("prim", "u4"), ("fi", *u3), ("light", *f3), ("p0", *f3), ("p1", *f3), ("p2", *f3), ("p", *f3),
                  ("b0", "f4"), ("b1", "f4"), ("sh_n", *f3), ("n0", *f3), ("n1", *f3), ("n2", *f3), ("hf", *f3),
                  ("bw", *f3), ("lr_dir_sum", "f4"), ("direct_valid", "u4"), ("direct_fi", *u3), ("direct_bw", *f3),
                  ("direct_dis", "f4"), ("direct_any", "u4"), ("mesh_flags", "u4"), ("has_normals", "u4")])
    p = np.dtype([("cam", *f3), ("d", *f3), ("dlduv", "f4", 2), ("dldp1", *f3), ("L", *f3), ("n_vertices", "u4"),
                  ("v", v, 5)])
    assert p.itemsize == C.sizeof(Path), (p.itemsize, C.sizeof(Path))
    return p
