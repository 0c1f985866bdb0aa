"""Host side of the B200 EPSM path: a thin Python mirror of the reference's integrator-plugin interface
over the C ABI of libepsm_b200.so.

Reference surface mirrored here (paths relative to the reference root):
  * `mi.load_dict({'type': 'manifold' | 'manifold_caustic', 'max_depth': …, 'rr_depth': …})`
    (EPSM/optim.py:97-100; properties src/python/python/ad/integrators/common.py:28-43)
  * `mi.register_integrator(name, factory)` (epsm.py:948,1202)
  * `Integrator.render(scene, sensor, seed, spp)` → (H,W,5) tensor (epsm.py:13-82)
  * `Integrator.render_backward(scene, params, grad_in, sensor, seed, spp)` (epsm.py:84-306)
  * `mi.traverse(scene)` / `params[key] = …; params.update()` (src/python/python/util.py:181-228)
  * `mi.render(scene, params, sensor, integrator, seed, spp)` as a differentiable op (util.py:352-519)

PyTorch is used for device memory, streams and torch.distributed only; all arithmetic of the path runs
in the hand-written CUDA kernels.  There is no CPU fallback: constructing a Scene without a CUDA device
or without the built library raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Dict, Optional

import numpy as np
import torch

from . import capi
from .scenes import SceneData

TILE_H = 4   # pixel-tile height of the kernels; multi-GPU bands are aligned to it


def sample_tea_32(v0: int, v1: int, rounds: int = 4):
    """include/mitsuba/core/random.h:76-90 (host integer restatement; used for seed_grad, util.py:510-515)."""
    m = 0xFFFFFFFF
    s = 0
    for _ in range(rounds):
        s = (s + 0x9e3779b9) & m
        v0 = (v0 + ((((v1 << 4) & m) + 0xa341316c) ^ ((v1 + s) & m) ^ ((v1 >> 5) + 0xc8013ea4))) & m
        v1 = (v1 + ((((v0 << 4) & m) + 0xad90777d) ^ ((v0 + s) & m) ^ ((v0 >> 5) + 0x7e95761e))) & m
    return v0, v1


class EpsmError(RuntimeError):
    pass


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def _sync_torch(device):
    """Order torch's stream before the library's stream (the library runs on its own CUDA stream)."""
    if torch.device(device).type == "cuda":
        torch.cuda.current_stream(device).synchronize()


def sample_shard(spp: int, rank: int, world: int):
    """Sample-index shard of `rank`: samples [s0, s1) of every pixel (epsm_set_sample_range).  Balanced by
    construction (every rank sees every pixel) and exact: the union over ranks is the single-GPU wavefront."""
    return (spp * rank) // world, (spp * (rank + 1)) // world


def band_lanes(height: int, width: int, spp: int, rank: int, world: int):
    """Scanline band of `rank`: whole tile rows, so that band g owns rows [y0, y1).
    Each lane's random stream depends only on (seed, global lane index), so sharding is exact."""
    tiles = (height + TILE_H - 1) // TILE_H
    t0, t1 = (tiles * rank) // world, (tiles * (rank + 1)) // world
    y0, y1 = min(height, t0 * TILE_H), min(height, t1 * TILE_H)
    return y0 * width * spp, y1 * width * spp


class Scene:
    """Device-resident scene (replaces mi.load_dict(scene) + accel init)."""

    def __init__(self, data: SceneData, device: int = 0):
        if not torch.cuda.is_available():
            raise EpsmError("epsm_mitsuba3_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = capi.load_library()
        self.data = data
        self.device = torch.device("cuda", device)
        ctx = C.c_void_p()
        rc = self.lib.epsm_create(device, C.byref(ctx))
        if rc != 0:
            raise EpsmError(self.lib.epsm_last_error(None).decode())
        self.ctx = ctx
        c = data.to_c()
        nm, nb, ne, ns = c.counts
        self._check(self.lib.epsm_scene_upload(ctx, c.meshes, nm, c.bsdfs, nb, c.emitters, ne, c.sensors, ns))
        self.voff = data.vertex_offsets()
        self.n_vertices = int(self.voff[-1])
        # how params.update() refreshes the BVH (apply_updates): 'auto' | 'refit' | 'rebuild'
        self.accel_update = os.environ.get("EPSM_ACCEL_UPDATE", "auto")
        if self.accel_update not in ("auto", "refit", "rebuild"):
            raise EpsmError(f"EPSM_ACCEL_UPDATE must be auto, refit or rebuild, not {self.accel_update!r}")
        self.rebuild_cost_ratio = 1.5
        self.rebuild_every = 0

    # ---- plumbing --------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise EpsmError(self.lib.epsm_last_error(self.ctx).decode())

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.epsm_destroy(self.ctx); self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ptr(self, t):
        """(address, mem) of a contiguous float32/uint32 torch tensor or numpy array."""
        if isinstance(t, np.ndarray):
            assert t.flags["C_CONTIGUOUS"]
            return t.ctypes.data, capi.MEM_HOST
        assert t.is_contiguous()
        return t.data_ptr(), (capi.MEM_DEVICE if t.is_cuda else capi.MEM_HOST)

    def synchronize(self):
        self._check(self.lib.epsm_synchronize(self.ctx))

    def stream(self) -> torch.cuda.ExternalStream:
        return torch.cuda.ExternalStream(self.lib.epsm_stream(self.ctx), device=self.device)

    def launch_count(self) -> int:
        return int(self.lib.epsm_launch_count(self.ctx))

    def stats(self) -> dict:
        s = capi.SceneStats(); self._check(self.lib.epsm_scene_get_stats(self.ctx, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_}

    def sensors(self):
        return self.data.sensors

    def set_trace_profiling(self, enable: bool):
        self._check(self.lib.epsm_set_trace_profiling(self.ctx, 1 if enable else 0))

    def trace_profile(self):
        """(milliseconds, launches) spent in the traversal kernel since the last call (needs set_trace_profiling(True))."""
        ms, n = C.c_double(), C.c_uint64()
        self._check(self.lib.epsm_get_trace_profile(self.ctx, C.byref(ms), C.byref(n)))
        return ms.value, int(n.value)

    def set_full_replay(self, enable: bool):
        """Gradient pass without liveness pruning (epsm_set_full_replay): every lane replays min(max_depth, 6) iterations."""
        self._check(self.lib.epsm_set_full_replay(self.ctx, 1 if enable else 0))

    def set_sample_range(self, s_begin: int = 0, s_end: int = 0):
        self._check(self.lib.epsm_set_sample_range(self.ctx, s_begin, s_end))

    # ---- scene updates ---------------------------------------------------------------------
    def update_vertices(self, mesh_id: int, positions, commit: bool = True, synced: bool = False):
        """params['<mesh>.vertex_positions'] = positions; params.update().  `commit=True` applies `self.accel_update`.
        `synced=True`: the caller has already ordered torch's stream before the library's (one sync for many meshes)."""
        if isinstance(positions, torch.Tensor):
            positions = positions.detach().to(torch.float32).contiguous()
            if positions.is_cuda and not synced:
                torch.cuda.current_stream(self.device).synchronize()
        else:
            positions = np.ascontiguousarray(positions, np.float32)
        n = positions.size if isinstance(positions, np.ndarray) else positions.numel()
        if n != 3 * self.data.meshes[mesh_id].positions.shape[0]:
            raise EpsmError(f"mesh {mesh_id}: vertex_positions has {n} values, expected {3 * self.data.meshes[mesh_id].positions.shape[0]}")
        addr, mem = self._ptr(positions)
        self._check(self.lib.epsm_scene_update_vertices(self.ctx, mesh_id, addr, mem))
        if commit:
            self.apply_updates()

    def update_normals(self, mesh_id: int, normals):
        """params['<mesh>.vertex_normals'] = normals; params.update() (epsm_scene_update_normals)."""
        if isinstance(normals, torch.Tensor):
            normals = normals.detach().to(torch.float32).contiguous()
            if normals.is_cuda:
                torch.cuda.current_stream(self.device).synchronize()
        else:
            normals = np.ascontiguousarray(normals, np.float32)
        n = normals.size if isinstance(normals, np.ndarray) else normals.numel()
        if n != 3 * self.data.meshes[mesh_id].positions.shape[0]:
            raise EpsmError(f"mesh {mesh_id}: vertex_normals has {n} values, expected {3 * self.data.meshes[mesh_id].positions.shape[0]}")
        addr, mem = self._ptr(normals)
        self._check(self.lib.epsm_scene_update_normals(self.ctx, mesh_id, addr, mem))

    def set_mesh_grad(self, mesh_id: int, enable: bool = True):
        """dr.enable_grad on a mesh's vertex buffers (epsm_scene_set_mesh_grad): the gradient passes scatter into it."""
        if bool(self.data.meshes[mesh_id].grad) != bool(enable):
            self._check(self.lib.epsm_scene_set_mesh_grad(self.ctx, mesh_id, 1 if enable else 0))
            self.data.meshes[mesh_id].grad = bool(enable)

    def update_sensors(self):
        """params['<sensor>.to_world'] = …; params.update(): push the poses in `self.data.sensors` (epsm_scene_update_sensor)."""
        c = self.data.to_c()
        for i in range(len(self.data.sensors)):
            self._check(self.lib.epsm_scene_update_sensor(self.ctx, i, C.byref(c.sensors[i])))

    def vertex_buffers(self):
        """Zero-copy (V,3) device views of the scene's current vertex positions and normals."""
        pp, pn, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        self._check(self.lib.epsm_scene_vertex_buffers(self.ctx, C.byref(pp), C.byref(pn), C.byref(n)))
        shape = (n.value // 3, 3)
        return _wrap_device(pp.value, shape, self.device), _wrap_device(pn.value, shape, self.device)

    def commit(self):
        """Host rebuild of the BVH (what the reference does on every parameters_changed, scene.cpp:304-336)."""
        self._check(self.lib.epsm_scene_commit(self.ctx))

    def refit(self):
        """Device refit of the BVH: topology of the last build kept, boxes and triangle records recomputed."""
        self._check(self.lib.epsm_scene_refit(self.ctx))

    def accel_cost(self):
        """(cost of the tree as built, cost now, refits since the last build) — Σ node half-areas."""
        b, n, r = C.c_double(), C.c_double(), C.c_uint64()
        self._check(self.lib.epsm_scene_accel_cost(self.ctx, C.byref(b), C.byref(n), C.byref(r)))
        return b.value, n.value, r.value

    def accel_raw(self):
        """Device views of the BVH4 nodes (n,32) and triangle slots (n,16), for tests."""
        pn, nn, pt, nt = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_uint64()
        self._check(self.lib.epsm_scene_accel_raw(self.ctx, C.byref(pn), C.byref(nn), C.byref(pt), C.byref(nt)))
        self.synchronize()        # the views are read on torch's stream: a refit still running on the library's stream must be done
        return _wrap_device(pn.value, (nn.value // 32, 32), self.device), _wrap_device(pt.value, (nt.value // 16, 16), self.device)

    def apply_updates(self):
        """Make pending vertex updates visible to the next render.  Policy `self.accel_update`:
        'rebuild' — host SAH rebuild every time (the reference's behaviour);
        'refit'   — device refit every time;
        'auto'    — device refit, host rebuild once the refitted tree's cost exceeds `self.rebuild_cost_ratio` × the
                    cost of the tree as built (default 1.5) or after `self.rebuild_every` refits (default 0 = never)."""
        mode = self.accel_update
        if mode == "rebuild":
            return self.commit()
        self.refit()
        if mode == "auto":
            built, now, refits = self.accel_cost()
            if now > self.rebuild_cost_ratio * built or (self.rebuild_every and refits >= self.rebuild_every):
                self.commit()                                  # commit after a refit = fresh build of the current positions

    # ---- low-level entry points (numpy or torch in, same kind out) --------------------------
    def trace(self, rays, any_hit: bool = False, counters: Optional[np.ndarray] = None):
        n = rays.shape[0]
        addr, mem = self._ptr(rays)
        if isinstance(rays, np.ndarray):
            hits = np.zeros(n if any_hit else (n, 4), np.uint32)
        else:
            hits = torch.zeros(n if any_hit else (n, 4), dtype=torch.int32, device=rays.device)
        cp = counters.ctypes.data if counters is not None else None
        self._check(self.lib.epsm_trace(self.ctx, addr, n, self._ptr(hits)[0], 1 if any_hit else 0, cp, mem))
        return hits

    def sampler_stream(self, seed, lane_begin, lane_end, n_floats):
        out = np.zeros((lane_end - lane_begin, n_floats), np.float32)
        self._check(self.lib.epsm_sampler_stream(self.ctx, seed, lane_begin, lane_end, n_floats, out.ctypes.data, capi.MEM_HOST))
        return out

    def sample_rays(self, sensor, seed, spp, lane_begin, lane_end):
        out = np.zeros((lane_end - lane_begin, 14), np.float32)
        self._check(self.lib.epsm_sample_rays(self.ctx, sensor, seed, spp, lane_begin, lane_end, out.ctypes.data, capi.MEM_HOST))
        return out

    def film_clear(self, sensor):
        self._check(self.lib.epsm_film_clear(self.ctx, sensor))

    def film_raw(self, sensor) -> torch.Tensor:
        """The context's RGBW accumulation buffer as a (H,W,4) torch view (no copy)."""
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self.lib.epsm_film_raw(self.ctx, sensor, C.byref(p), C.byref(n)))
        s = self.data.sensors[sensor]
        return _wrap_device(p.value, (s.height, s.width, 4), self.device)

    def render_accumulate(self, sensor, seed, spp, max_depth, rr_depth, lane_begin=0, lane_end=0, counters=None):
        cp = counters.ctypes.data if counters is not None else None
        self._check(self.lib.epsm_render_accumulate(self.ctx, sensor, seed, spp, max_depth, rr_depth, lane_begin, lane_end, cp))

    def film_develop(self, sensor, out=None):
        s = self.data.sensors[sensor]
        if out is None:
            out = torch.empty((s.height, s.width, 5), dtype=torch.float32, device=self.device)
        addr, mem = self._ptr(out)
        self._check(self.lib.epsm_film_develop(self.ctx, sensor, addr, mem))
        return out

    def render_backward_raw(self, variant, sensor, seed, spp, max_depth, rr_depth, grad_in,
                            grad_pos=None, grad_nrm=None, lane_begin=0, lane_end=0, counters=None):
        """epsm_render_backward: accumulates into (and returns) grad_pos / grad_nrm, (V,3) each.
        Buffers follow the memory space of grad_in (host numpy / pinned tensor ↔ device tensor)."""
        vid = _variant_id(variant)
        s = self.data.sensors[sensor]
        host = isinstance(grad_in, np.ndarray) or not grad_in.is_cuda
        if isinstance(grad_in, np.ndarray):
            grad_in = np.ascontiguousarray(grad_in, np.float32)
        else:
            grad_in = grad_in.detach().to(torch.float32).contiguous()
        assert tuple(grad_in.shape) == (s.height, s.width, 5), "grad_in must be (H,W,5) for the gradient sensor"
        if grad_pos is None:
            if isinstance(grad_in, np.ndarray):
                grad_pos = np.zeros((self.n_vertices, 3), np.float32); grad_nrm = np.zeros((self.n_vertices, 3), np.float32)
            else:
                kw = dict(dtype=torch.float32, device=grad_in.device, pin_memory=host)
                grad_pos = torch.zeros((self.n_vertices, 3), **kw); grad_nrm = torch.zeros((self.n_vertices, 3), **kw)
        a_in, mem = self._ptr(grad_in)
        cp = counters.ctypes.data if counters is not None else None
        self._check(self.lib.epsm_render_backward(self.ctx, vid, sensor, seed, spp, max_depth, rr_depth, a_in,
                                                  self._ptr(grad_pos)[0], self._ptr(grad_nrm)[0], mem, lane_begin, lane_end, cp))
        return grad_pos, grad_nrm

    def vertex_normals_backward(self, grad_nrm, grad_pos):
        """grad_pos += Jᵀ·grad_nrm: the adjoint of Mesh::recompute_vertex_normals (mesh.cpp:311-330) for every
        differentiable mesh with vertex normals (epsm_vertex_normals_backward).  Same memory space for both."""
        a, mem = self._ptr(grad_nrm); b, mem2 = self._ptr(grad_pos)
        if mem != mem2:
            raise EpsmError("grad_nrm and grad_pos must live in the same memory space")
        self._check(self.lib.epsm_vertex_normals_backward(self.ctx, a, b, mem))
        return grad_pos

    def prb_reparam_backward_raw(self, sensor, seed, spp, integ, grad_in, grad_pos, grad_nrm, lane_begin=0, lane_end=0, film_ready=False):
        """epsm_prb_reparam_backward: accumulates into grad_pos / grad_nrm ((V,3) each, memory space of grad_in)."""
        if isinstance(grad_in, np.ndarray):
            grad_in = np.ascontiguousarray(grad_in, np.float32)
        else:
            grad_in = grad_in.detach().to(torch.float32).contiguous()
        a_in, mem = self._ptr(grad_in)
        cfg = capi.ReparamConfig(max_depth=integ.max_depth, rr_depth=integ.rr_depth, reparam_max_depth=min(integ.reparam_max_depth, 0xffffffff),
                                 reparam_rays=integ.reparam_rays, reparam_kappa=integ.reparam_kappa, reparam_exp=integ.reparam_exp)
        self._check(self.lib.epsm_prb_reparam_backward(self.ctx, sensor, seed, spp, C.byref(cfg), a_in, self._ptr(grad_pos)[0],
                                                       self._ptr(grad_nrm)[0], mem, lane_begin, lane_end, 1 if film_ready else 0))
        return grad_pos, grad_nrm

    def path_records(self, variant, sensor, seed, spp, max_depth, rr_depth, grad_in, lane_begin, lane_end):
        grad_in = np.ascontiguousarray(grad_in, np.float32)
        out = np.zeros((lane_end - lane_begin, capi.REC_FLOATS), np.float32)
        self._check(self.lib.epsm_path_records(self.ctx, _variant_id(variant), sensor, seed, spp, max_depth, rr_depth,
                                               grad_in.ctypes.data, lane_begin, lane_end, out.ctypes.data, capi.MEM_HOST))
        return out


def _variant_id(v):
    if v in (capi.VARIANT_MANIFOLD, "manifold"):
        return capi.VARIANT_MANIFOLD
    if v in (capi.VARIANT_MANIFOLD_CAUSTIC, "manifold_caustic"):
        return capi.VARIANT_MANIFOLD_CAUSTIC
    raise EpsmError(f"unknown integrator variant {v!r}")


def _wrap_device(ptr: int, shape, device) -> torch.Tensor:
    """Zero-copy torch view of library-owned device memory (via __cuda_array_interface__)."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}
    return torch.as_tensor(h, device=device)


# ------------------------------------------------------------------------------------------------
#  Integrator plugins
# ------------------------------------------------------------------------------------------------
class Integrator:
    """Property handling of ADIntegrator (common.py:28-43) and the primal render every integrator of this path shares:
    the PRB loop of EPSMIntegrator.sample (epsm.py:308-501) is the estimator of `path` (src/integrators/path.cpp:1-324)
    and of `prb` / `prb_reparam` in primal mode (NEE + BSDF sampling, MIS, Russian roulette from rr_depth)."""
    channels: int = 3                                        # RGB; the EPSM integrators append two zero channels (epsm.py:76-82)
    type_name: str = None

    def __init__(self, props: dict):
        max_depth = int(props.get("max_depth", self.default_max_depth))
        if max_depth < 0 and max_depth != -1:
            raise Exception("\"max_depth\" must be set to -1 (infinite) or a value >= 0")
        self.max_depth = max_depth if max_depth != -1 else 0xffffffff
        self.rr_depth = int(props.get("rr_depth", 5))
        if self.rr_depth <= 0:
            raise Exception("\"rr_depth\" must be set to a value greater than zero!")

    default_max_depth = 6

    def render(self, scene: Scene, sensor: int = 0, seed: int = 0, spp: int = 0, develop: bool = True, evaluate: bool = True) -> torch.Tensor:
        if not develop:
            raise Exception("develop=True must be specified when invoking AD integrators")
        s = scene.sensors()[sensor]
        spp = spp if spp != 0 else s.spp
        scene.film_clear(sensor)
        # Mitsuba renders wavefronts of at most 2^32 samples and splits larger jobs into passes with successive seeds
        # (common.py:468-475 / integrator.cpp); the 8192-spp target renders of optim.py:56-58 stay below the limit at 512².
        passes = 1
        while s.wavefront(spp) // passes > (1 << 32) or spp % passes:
            passes += 1
        dist = _dist()
        for p in range(passes):
            spp_p = spp // passes
            if dist is None:
                scene.render_accumulate(sensor, seed + p, spp_p, self.max_depth, self.rr_depth)
            else:
                s0, s1 = sample_shard(spp_p, dist.get_rank(), dist.get_world_size())
                if s1 > s0:
                    scene.set_sample_range(s0, s1)
                    try:
                        scene.render_accumulate(sensor, seed + p, spp_p, self.max_depth, self.rr_depth)
                    finally:
                        scene.set_sample_range(0, 0)      # never leave the context sharded for later calls
        if dist is not None:
            film = scene.film_raw(sensor)
            scene.synchronize()
            dist.all_reduce(film)                     # every rank splatted its samples of every pixel: sum the RGBW film
        # the library's stream writes into a fresh torch allocation: order torch's stream first (the caching allocator
        # may hand back a block with pending torch-stream work)
        _sync_torch(scene.device)
        img = scene.film_develop(sensor)
        scene.synchronize()
        self.primal_image = img if self.channels == 5 else img[..., :self.channels].contiguous()
        return self.primal_image

    def render_backward(self, scene: Scene, params, grad_in: torch.Tensor, sensor: int = 0, seed: int = 0, spp: int = 0) -> None:
        raise NotImplementedError(f"{type(self).__name__}.render_backward")          # Integrator::render_backward (integrator.h) for non-AD plugins

    def vertex_gradients(self, scene: Scene, grad_out: torch.Tensor, sensor: int, seed: int, spp: int):
        """(grad_pos, grad_nrm), (V,3) each: gradient of Σ image·grad_out w.r.t. the concatenated vertex positions (normals
        held fixed) and vertex normals (what _RenderOp.backward routes)."""
        raise NotImplementedError(f"{type(self).__name__} is not differentiable")

    def position_gradient(self, scene: Scene, grad_out: torch.Tensor, sensor: int, seed: int, spp: int) -> torch.Tensor:
        """Total (V,3) gradient w.r.t. the vertex positions, incl. the normals' dependence on them (mesh.cpp:82-87)."""
        gp, gn = self.vertex_gradients(scene, grad_out, sensor, seed, spp)
        return total_position_gradient(scene, gp, gn)


class PathIntegrator(Integrator):
    """`path` (src/integrators/path.cpp:1-324): the primal path tracer the drivers use for target / reference images
    (optim.py:56-58,80,160).  max_depth defaults to -1 (path.cpp / integrator.cpp MonteCarloIntegrator)."""
    type_name = "path"
    default_max_depth = -1


class EPSMIntegrator(Integrator):
    """EPSMIntegrator (epsm.py:12)."""
    variant: str = None
    channels = 5

    # epsm.py:84-306
    def render_backward(self, scene: Scene, params, grad_in: torch.Tensor, sensor: int = 0, seed: int = 0, spp: int = 0) -> None:
        """Accumulates into `.grad` of every `<mesh>.vertex_positions` / `<mesh>.vertex_normals` entry of `params`
        that requires grad.  As in the reference, the sensor and spp arguments are ignored: the gradient pass
        always uses sensors()[2] at 8 spp (epsm.py:142-145)."""
        sensor, spp = 2, 8
        s = scene.sensors()[sensor]
        if grad_in.shape[-1] == 3:
            # colour-only gradients reach no scene parameter in the reference (δL is never consumed, epsm.py:728-732)
            return
        res = s.width                                   # epsm.py:238-240: crop to the gradient sensor's resolution
        g = grad_in[:res, :res, :].to(device=scene.device, dtype=torch.float32).contiguous()
        _sync_torch(scene.device)
        gp, gn = self.backward_buffers(scene, g, sensor, seed, spp)
        params.accumulate_gradients(gp, gn)

    def vertex_gradients(self, scene: Scene, grad_out: torch.Tensor, sensor: int, seed: int, spp: int):
        s2 = scene.sensors()[2]                         # sensor / spp ignored: sensors()[2] at 8 spp (epsm.py:142-145)
        if grad_out.shape[-1] == 3:
            z = torch.zeros((2, scene.n_vertices, 3), dtype=torch.float32, device=scene.device)
            return z[0], z[1]
        g = grad_out[:s2.width, :s2.width, :].to(torch.float32).contiguous()
        return self.backward_buffers(scene, g, 2, seed, 8)

    def backward_buffers(self, scene: Scene, g: torch.Tensor, sensor: int, seed: int, spp: int):
        s = scene.sensors()[sensor]
        dist = _dist()
        s0, s1 = (0, spp) if dist is None else sample_shard(spp, dist.get_rank(), dist.get_world_size())
        buf = torch.zeros((2, scene.n_vertices, 3), dtype=torch.float32, device=scene.device)
        _sync_torch(scene.device)
        if s1 > s0:
            scene.set_sample_range(s0, s1)
            try:
                scene.render_backward_raw(self.variant, sensor, seed, spp, self.max_depth, self.rr_depth, g, buf[0], buf[1])
            finally:
                scene.set_sample_range(0, 0)
        scene.synchronize()
        if dist is not None:
            dist.all_reduce(buf)                        # one NCCL allreduce of the mesh-gradient buffer per step
            _sync_torch(scene.device)
        return buf[0], buf[1]


class ManifoldIntegrator(EPSMIntegrator):              # epsm.py:744
    variant = type_name = "manifold"


class ManifoldCausticIntegrator(EPSMIntegrator):       # epsm.py:951
    variant = type_name = "manifold_caustic"


class PRBReparamIntegrator(Integrator):
    """`prb_reparam` (src/python/python/ad/integrators/prb_reparam.py:8-609 over RBIntegrator, common.py:574-975): hybrid
    phase 2 of the EPSM drivers (optim.py:89-92,113-119), sensor 0, L2 loss.  Properties of prb_reparam.py:225-247."""
    type_name = "prb_reparam"

    def __init__(self, props: dict):
        super().__init__(props)
        self.reparam_max_depth = int(props.get("reparam_max_depth", self.max_depth))
        self.reparam_rays = int(props.get("reparam_rays", 16))
        self.reparam_kappa = float(props.get("reparam_kappa", 1e5))
        self.reparam_exp = float(props.get("reparam_exp", 3.0))
        self.reparam_antithetic = bool(props.get("reparam_antithetic", False))
        self.reparam_unroll = bool(props.get("reparam_unroll", False))
        if self.reparam_antithetic:
            raise NotImplementedError("prb_reparam: reparam_antithetic=True is not built (default False; no EPSM driver sets it)")

    def vertex_gradients(self, scene: Scene, grad_out: torch.Tensor, sensor: int, seed: int, spp: int):
        """RBIntegrator.render_backward (common.py:803-975): grad_out is (H,W,3) for the film of `sensor`."""
        s = scene.sensors()[sensor]
        spp = spp if spp != 0 else s.spp
        g = grad_out[..., :3].to(device=scene.device, dtype=torch.float32).contiguous()
        if tuple(g.shape) != (s.height, s.width, 3):
            raise EpsmError(f"prb_reparam: grad_in has shape {tuple(g.shape)}, the film of sensor {sensor} is {(s.height, s.width, 3)}")
        dist = _dist()
        s0, s1 = (0, spp) if dist is None else sample_shard(spp, dist.get_rank(), dist.get_world_size())
        buf = torch.zeros((2, scene.n_vertices, 3), dtype=torch.float32, device=scene.device)
        _sync_torch(scene.device)
        if s1 > s0:
            scene.set_sample_range(s0, s1)
            try:
                scene.prb_reparam_backward_raw(sensor, seed, spp, self, g, buf[0], buf[1])
            finally:
                scene.set_sample_range(0, 0)
        scene.synchronize()
        if dist is not None:
            dist.all_reduce(buf)
            _sync_torch(scene.device)
        return buf[0], buf[1]

    def render_backward(self, scene: Scene, params, grad_in: torch.Tensor, sensor: int = 0, seed: int = 0, spp: int = 0) -> None:
        gp, gn = self.vertex_gradients(scene, grad_in, sensor, seed, spp)
        params.accumulate_gradients(gp, gn)


_REGISTRY: Dict[str, Callable[[dict], EPSMIntegrator]] = {}


def register_integrator(name: str, factory: Callable[[dict], EPSMIntegrator]) -> None:
    """mi.register_integrator (include/mitsuba/python/python.h:57-90)."""
    _REGISTRY[name] = factory


register_integrator("path", lambda props: PathIntegrator(props))                         # src/integrators/path.cpp
register_integrator("prb_reparam", lambda props: PRBReparamIntegrator(props))           # prb_reparam.py:609
register_integrator("manifold", lambda props: ManifoldIntegrator(props))                 # epsm.py:948
register_integrator("manifold_caustic", lambda props: ManifoldCausticIntegrator(props))  # epsm.py:1202


def load_dict(d: dict):
    """mi.load_dict for the integrator plugins of this path."""
    t = d.get("type")
    if t not in _REGISTRY:
        raise EpsmError(f'Plugin "{t}" not found')           # PluginManager::create_object, src/core/plugin.cpp:174-211
    return _REGISTRY[t]({k: v for k, v in d.items() if k != "type"})


# ------------------------------------------------------------------------------------------------
#  SceneParameters / mi.render
# ------------------------------------------------------------------------------------------------
class SceneParameters(dict):
    """mi.traverse(scene) restricted to mesh vertex data (util.py:24-228): `<mesh>.vertex_positions` for every mesh and
    `<mesh>.vertex_normals` for the meshes that have vertex normals (mesh.cpp:71-75), flat float32 tensors like the
    reference's buffers."""

    def __init__(self, scene: Scene):
        super().__init__()
        self.scene = scene
        self._dirty = set()
        self._mesh_of = {}          # vertex_positions key → mesh id
        self._nrm_of = {}           # vertex_normals key   → mesh id
        seen = set()
        for i, m in enumerate(scene.data.meshes):
            name = m.name if m.name not in seen else f"{m.name}_{i}"
            seen.add(name)
            key = f"{name}.vertex_positions"
            super().__setitem__(key, torch.from_numpy(m.positions.copy()).to(scene.device))
            self._mesh_of[key] = i
            if m.normals is not None:
                nkey = f"{name}.vertex_normals"
                super().__setitem__(nkey, torch.from_numpy(m.normals.copy()).to(scene.device))
                self._nrm_of[nkey] = i

    def __setitem__(self, key, value):
        if key not in self:
            raise KeyError(key)
        super().__setitem__(key, value)
        self._dirty.add(key)

    def update(self, values: Optional[dict] = None):
        if values:
            for k, v in values.items():
                self[k] = v
        stage = lambda key: torch.as_tensor(self[key]).detach().to(torch.float32).reshape(-1, 3).contiguous()
        staged = sorted(((self._mesh_of[k], stage(k)) for k in self._dirty if k in self._mesh_of), key=lambda t: t[0])
        staged_n = sorted(((self._nrm_of[k], stage(k)) for k in self._dirty if k in self._nrm_of), key=lambda t: t[0])
        if any(t.is_cuda for _, t in staged + staged_n):
            torch.cuda.current_stream(self.scene.device).synchronize()      # once: the tensors above are complete
        for mesh_id, t in staged:
            self.scene.update_vertices(mesh_id, t, commit=False, synced=True)
        if staged:
            self.scene.apply_updates()          # positions changed: normals re-derived (mesh.cpp:82-87), BVH refreshed
            _, nrm = self.scene.vertex_buffers()
            voff = self.scene.voff
            moved = {i for i, _ in staged}
            for nkey, i in self._nrm_of.items():
                if i in moved and nkey not in self._dirty:
                    dict.__setitem__(self, nkey, nrm[int(voff[i]):int(voff[i + 1])].clone())
            self.scene.synchronize()
        for mesh_id, t in staged_n:             # explicit normals win over the recomputed ones (keys = [vertex_normals])
            self.scene.update_normals(mesh_id, t)
        self._dirty.clear()

    def accumulate_gradients(self, grad_pos: torch.Tensor, grad_nrm: torch.Tensor):
        """dr.grad of the parameters after render_backward: `vertex_normals` entries that require grad receive grad_nrm;
        `vertex_positions` entries receive grad_pos plus grad_nrm pulled back through recompute_vertex_normals — the
        normals are a differentiable function of the positions after params.update() (mesh.cpp:82-87,311-330)."""
        voff = self.scene.voff
        for key, i in self._nrm_of.items():
            t = self[key]
            if isinstance(t, torch.Tensor) and t.requires_grad:
                g = grad_nrm[int(voff[i]):int(voff[i + 1])].reshape(t.shape)
                t.grad = g.clone() if t.grad is None else t.grad + g
        total = total_position_gradient(self.scene, grad_pos, grad_nrm)
        for key, i in self._mesh_of.items():
            t = self[key]
            if not (isinstance(t, torch.Tensor) and t.requires_grad):
                continue
            g = total[int(voff[i]):int(voff[i + 1])].reshape(t.shape)
            t.grad = g.clone() if t.grad is None else t.grad + g


def total_position_gradient(scene: Scene, grad_pos: torch.Tensor, grad_nrm: torch.Tensor) -> torch.Tensor:
    """grad_pos + Jᵀ·grad_nrm (adjoint of Mesh::recompute_vertex_normals, epsm_vertex_normals_backward)."""
    total = grad_pos.clone()
    if any(m.normals is not None and m.grad for m in scene.data.meshes):
        _sync_torch(scene.device)
        scene.vertex_normals_backward(grad_nrm.contiguous(), total)
        scene.synchronize()
    return total


def traverse(scene: Scene) -> SceneParameters:
    return SceneParameters(scene)


trace_hook: Optional[Callable[[dict], None]] = None      # set by the stand-in `mitsuba` module when EPSM_SHIM_TRACE is given (tests)


class _RenderOp(torch.autograd.Function):
    """mi.render's custom op (util.py:352-403): forward = integrator.render, backward = render_backward with
    seed_grad = sample_tea_32(seed, 1)[0] unless given, spp_grad = spp unless given (util.py:505-515).
    `keys` lists (kind, mesh id) per differentiable input tensor: 'p' vertex_positions, 'n' vertex_normals."""

    @staticmethod
    def forward(ctx, scene, integrator, sensor, seed, spp, seed_grad, spp_grad, keys, *tensors):
        ctx.scene, ctx.integrator, ctx.sensor, ctx.keys = scene, integrator, sensor, keys
        ctx.seed_grad = seed_grad if seed_grad != 0 else sample_tea_32(seed, 1)[0]
        ctx.spp_grad = spp_grad if spp_grad != 0 else spp
        ctx.shapes = [t.shape for t in tensors]
        if trace_hook:
            trace_hook({"op": "render", "integrator": integrator.type_name, "sensor": sensor, "seed": seed, "spp": spp})
        return integrator.render(scene, sensor=sensor, seed=seed, spp=spp).clone()

    @staticmethod
    def backward(ctx, grad_out):
        scene, integ = ctx.scene, ctx.integrator
        gp, gn = integ.vertex_gradients(scene, grad_out, ctx.sensor, ctx.seed_grad, ctx.spp_grad)
        if trace_hook:
            trace_hook({"op": "backward", "integrator": integ.type_name, "sensor": ctx.sensor, "seed_grad": ctx.seed_grad,
                        "spp_grad": ctx.spp_grad, "grad_abs_sum": float(gp.abs().sum() + gn.abs().sum())})
        # the normals of a mesh whose vertex_normals are not an input are a function of its positions (mesh.cpp:82-87)
        explicit = {m for kind, m in ctx.keys if kind == "n"}
        if explicit:
            gn_implicit = gn.clone()
            for m in explicit:
                gn_implicit[int(scene.voff[m]):int(scene.voff[m + 1])] = 0
        else:
            gn_implicit = gn
        total = total_position_gradient(scene, gp, gn_implicit)
        voff = scene.voff
        outs = []
        for (kind, mesh_id), shape in zip(ctx.keys, ctx.shapes):
            src = total if kind == "p" else gn
            outs.append(src[int(voff[mesh_id]):int(voff[mesh_id + 1])].reshape(shape).clone())
        return (None,) * 8 + tuple(outs)


def render(scene: Scene, params: SceneParameters, sensor: int = 0, integrator: Integrator = None, seed: int = 0, spp: int = 0,
           seed_grad: int = 0, spp_grad: int = 0) -> torch.Tensor:
    """mi.render(scene, params, sensor, integrator, seed, seed_grad, spp, spp_grad): differentiable w.r.t. the
    vertex_positions / vertex_normals tensors in `params` that require grad."""
    keys, tensors = [], []
    for kind, table in (("p", params._mesh_of), ("n", params._nrm_of)):
        for key, mesh_id in table.items():
            t = params[key]
            if isinstance(t, torch.Tensor) and t.requires_grad:
                keys.append((kind, mesh_id)); tensors.append(t)
                scene.set_mesh_grad(mesh_id, True)          # attached buffers receive gradients (dr.enable_grad semantics)
    return _RenderOp.apply(scene, integrator, sensor, seed, spp, seed_grad, spp_grad, keys, *tensors)
