"""ctypes front-end of the TEST-ONLY host harness (libepsm_hostsim.so): the product's per-lane device code
compiled for the CPU.  Presents the same low-level calls as epsm_mitsuba3_b200.engine.Scene so that the
parity tests can run on the harness here (no GPU) and on the real library on the B200 (`-m gpu`)."""
import ctypes as C
import os

import numpy as np

from epsm_mitsuba3_b200 import capi
from epsm_mitsuba3_b200.scenes import SceneData

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        import __graft_entry__ as ge
        L = C.CDLL(ge.build_hostsim())
        vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
        L.sim_scene_create.restype = vp
        L.sim_scene_create.argtypes = [C.POINTER(capi.MeshDesc), u32, C.POINTER(capi.BsdfDesc), u32,
                                       C.POINTER(capi.EmitterDesc), u32, C.POINTER(capi.SensorDesc), u32]
        L.sim_scene_error.restype = C.c_char_p; L.sim_scene_error.argtypes = [vp]
        L.sim_scene_destroy.argtypes = [vp]
        L.sim_scene_update_vertices.argtypes = [vp, u32, vp]
        L.sim_scene_stats.argtypes = [vp, vp]
        L.sim_scene_refit.argtypes = [vp, u32, vp]; L.sim_scene_refit.restype = C.c_double
        L.sim_scene_cost.argtypes = [vp]; L.sim_scene_cost.restype = C.c_double
        L.sim_scene_copy_accel.argtypes = [vp, vp, vp]
        L.sim_surface_uv.argtypes = [vp, vp, vp]
        L.sim_bsdf_sample.argtypes = [vp, u32, vp, C.c_float, C.c_float, C.c_float, vp]
        L.sim_bsdf_eval_pdf.argtypes = [vp, u32, vp, vp, vp]
        L.sim_bsdf_grad.argtypes = [vp, u32, vp, vp, vp, vp]
        L.sim_trace.argtypes = [vp, vp, u64, vp, i32, vp]
        L.sim_sampler_stream.argtypes = [u32, u64, u64, u32, vp]
        L.sim_sample_rays.argtypes = [vp, u32, u32, u32, u64, u64, vp]
        L.sim_render_accumulate.argtypes = [vp, u32, u32, u32, u32, u32, u64, u64, vp, i32]
        L.sim_prb_backward.argtypes = [vp, u32, u32, u32, C.POINTER(capi.ReparamConfig), vp, vp, u64, u64, vp, vp, i32]
        L.sim_backward.argtypes = [vp, i32, u32, u32, u32, u32, u32, vp, u64, u64, vp, vp, vp, i32]
        _LIB = L
    return _LIB


def _vid(v):
    return {"manifold": 0, "manifold_caustic": 1}.get(v, v)


class SimScene:
    def __init__(self, sd: SceneData):
        self.data = sd; self._c = sd.to_c(); nm, nb, ne, ns = self._c.counts
        self.h = lib().sim_scene_create(self._c.meshes, nm, self._c.bsdfs, nb, self._c.emitters, ne, self._c.sensors, ns)
        err = lib().sim_scene_error(self.h).decode()
        if err:
            raise RuntimeError(err)
        self.n_vertices = sd.n_vertices

    def __del__(self):
        if getattr(self, "h", None):
            lib().sim_scene_destroy(self.h); self.h = None

    def stats(self):
        out = np.zeros(3, np.uint64); lib().sim_scene_stats(self.h, out.ctypes.data)
        return dict(n_triangles=int(out[0]), n_nodes=int(out[1]), depth=int(out[2]))

    def update_vertices(self, mesh_id, positions):
        p = np.ascontiguousarray(positions, np.float32); lib().sim_scene_update_vertices(self.h, mesh_id, p.ctypes.data)

    def refit_vertices(self, mesh_id, positions):
        """The device refit's per-node code, run serially on the host; returns the refitted tree's cost."""
        p = np.ascontiguousarray(positions, np.float32); return lib().sim_scene_refit(self.h, mesh_id, p.ctypes.data)

    def accel_cost(self):
        return lib().sim_scene_cost(self.h)

    def accel(self):
        st = self.stats()
        nodes = np.zeros((st["n_nodes"], 32), np.float32); tris = np.zeros((st["n_triangles"], 16), np.float32)
        lib().sim_scene_copy_accel(self.h, nodes.ctypes.data, tris.ctypes.data)
        return nodes, tris

    def surface_uv(self, o, d, maxt=np.float32(3.4e38)):
        r = np.array([o[0], o[1], o[2], maxt, d[0], d[1], d[2], 0], np.float32); out = np.zeros(11, np.float32)
        lib().sim_surface_uv(self.h, r.ctypes.data, out.ctypes.data)
        return dict(uv=out[0:2], sh_s=out[2:5], sh_t=out[5:8], reflectance=out[8:11])

    def bsdf_sample(self, bsdf_id, wi, s1, s2):
        wi = np.asarray(wi, np.float32); out = np.zeros(12, np.float32)
        lib().sim_bsdf_sample(self.h, bsdf_id, wi.ctypes.data, s1, s2[0], s2[1], out.ctypes.data)
        return dict(wo=out[0:3], pdf=out[3], eta=out[4], type=int(out[5]), weight=out[6:9], hf=out[9:12])

    def bsdf_eval_pdf(self, bsdf_id, wi, wo):
        wi = np.asarray(wi, np.float32); wo = np.asarray(wo, np.float32); out = np.zeros(4, np.float32)
        lib().sim_bsdf_eval_pdf(self.h, bsdf_id, wi.ctypes.data, wo.ctypes.data, out.ctypes.data)
        return out[:3], out[3]

    def bsdf_grad(self, bsdf_id, wi, wo, a):
        wi, wo, a = (np.ascontiguousarray(x, np.float32) for x in (wi, wo, a)); out = np.zeros(9, np.float32)
        lib().sim_bsdf_grad(self.h, bsdf_id, wi.ctypes.data, wo.ctypes.data, a.ctypes.data, out.ctypes.data)
        return out[0:3], out[3:6], out[6:9]

    def trace(self, rays, any_hit=False, counters=None):
        rays = np.ascontiguousarray(rays, np.float32); n = rays.shape[0]
        hits = np.zeros(n if any_hit else (n, 4), np.uint32)
        lib().sim_trace(self.h, rays.ctypes.data, n, hits.ctypes.data, 1 if any_hit else 0,
                        counters.ctypes.data if counters is not None else None)
        return hits

    def sampler_stream(self, seed, lb, le, n):
        out = np.zeros((le - lb, n), np.float32); lib().sim_sampler_stream(seed, lb, le, n, out.ctypes.data); return out

    def sample_rays(self, sensor, seed, spp, lb, le):
        out = np.zeros((le - lb, 14), np.float32); lib().sim_sample_rays(self.h, sensor, seed, spp, lb, le, out.ctypes.data); return out

    def render_raw(self, sensor, seed, spp, max_depth=6, rr_depth=5, lb=0, le=None, staged=False):
        s = self.data.sensors[sensor]; le = s.wavefront(spp) if le is None else le
        film = np.zeros((s.height, s.width, 4), np.float32)
        lib().sim_render_accumulate(self.h, sensor, seed, spp, max_depth, rr_depth, lb, le, film.ctypes.data, 1 if staged else 0)
        return film

    def path_records(self, variant, sensor, seed, spp, max_depth, rr_depth, grad_in, lb, le):
        g = np.ascontiguousarray(grad_in, np.float32); out = np.zeros((le - lb, capi.REC_FLOATS), np.float32)
        lib().sim_backward(self.h, _vid(variant), sensor, seed, spp, max_depth, rr_depth, g.ctypes.data, lb, le, None, None, out.ctypes.data, 0)
        return out

    def backward(self, variant, sensor, seed, spp, max_depth, rr_depth, grad_in, lb=0, le=None, prune=True):
        s = self.data.sensors[sensor]; le = s.wavefront(spp) if le is None else le
        g = np.ascontiguousarray(grad_in, np.float32)
        gp = np.zeros((self.n_vertices, 3), np.float32); gn = np.zeros((self.n_vertices, 3), np.float32)
        lib().sim_backward(self.h, _vid(variant), sensor, seed, spp, max_depth, rr_depth, g.ctypes.data, lb, le,
                           gp.ctypes.data, gn.ctypes.data, None, int(prune))     # 2 = production schedule (fast path + chain prefix)
        return gp, gn

    def prb_backward(self, sensor, seed, spp, cfg: dict, film, grad_in, lb=0, le=None, wavefront=False):
        """k_prb_backward's lane code (epsm_prb.cuh) on the host.  film: (H,W,4) primal film of all lanes; grad_in (H,W,3)."""
        s = self.data.sensors[sensor]; le = s.wavefront(spp) if le is None else le
        c = capi.ReparamConfig(**cfg)
        film = np.ascontiguousarray(film, np.float32); g = np.ascontiguousarray(grad_in, np.float32)
        assert film.shape == (s.height, s.width, 4) and g.shape == (s.height, s.width, 3)
        gp = np.zeros((self.n_vertices, 3), np.float32); gn = np.zeros((self.n_vertices, 3), np.float32)
        lib().sim_prb_backward(self.h, sensor, seed, spp, C.byref(c), film.ctypes.data, g.ctypes.data, lb, le, gp.ctypes.data, gn.ctypes.data, 1 if wavefront else 0)
        return gp, gn
