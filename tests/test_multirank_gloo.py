"""The N > 1 host path on CPU: two ranks over `gloo`, each running the engine's own render / render_backward host logic
(sample_shard, set_sample_range, one all_reduce of the film / of the [grad_pos | grad_nrm] buffer) on top of the test-only
host harness (tests/hostsim: the product's per-lane device code compiled for the CPU) instead of the CUDA library.
The reduced results must equal the single-rank pass (sum order differs ⇒ tolerance)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class SimBackedScene:
    """Duck-types the slice of engine.Scene that EPSMIntegrator.render / backward_buffers use."""

    def __init__(self, sd):
        from tests.hostsim.hostsim import SimScene
        self.data, self.sim = sd, SimScene(sd)
        self.device = torch.device("cpu")
        self.voff = sd.vertex_offsets(); self.n_vertices = int(self.voff[-1])
        self.range = (0, 0); self.films = {}

    def sensors(self): return self.data.sensors
    def synchronize(self): pass
    def set_sample_range(self, s0=0, s1=0): self.range = (s0, s1)

    def _lane_ranges(self, sensor, spp):
        s = self.data.sensors[sensor]; s0, s1 = self.range if self.range != (0, 0) else (0, spp)
        return [(p * spp + s0, p * spp + s1) for p in range(s.width * s.height)]

    def film_clear(self, sensor):
        s = self.data.sensors[sensor]; self.films[sensor] = torch.zeros((s.height, s.width, 4))
    def film_raw(self, sensor): return self.films[sensor]
    def render_accumulate(self, sensor, seed, spp, max_depth, rr_depth, lane_begin=0, lane_end=0, counters=None):
        for lb, le in self._lane_ranges(sensor, spp):
            self.films[sensor] += torch.from_numpy(self.sim.render_raw(sensor, seed, spp, max_depth, rr_depth, lb, le))
    def film_develop(self, sensor, out=None):
        f = self.films[sensor]; w = torch.where(f[..., 3:] == 0, torch.ones_like(f[..., 3:]), f[..., 3:])
        return torch.cat([f[..., :3] / w, torch.zeros_like(f[..., :2])], -1)

    def render_backward_raw(self, variant, sensor, seed, spp, max_depth, rr_depth, grad_in, grad_pos=None, grad_nrm=None, **kw):
        g = grad_in.numpy()
        for lb, le in self._lane_ranges(sensor, spp):
            gp, gn = self.sim.backward(variant, sensor, seed, spp, max_depth, rr_depth, g, lb, le)
            grad_pos += torch.from_numpy(gp); grad_nrm += torch.from_numpy(gn)
        return grad_pos, grad_nrm


def _run(rank, world, port, variant, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import epsm_mitsuba3_b200 as ep
    from epsm_mitsuba3_b200 import scenes as S
    from tests import parity
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    sd = S.glass_sphere_caustic(12, 4, 12, subdiv=2) if variant == "manifold_caustic" else S.mirror_plane_scene(12, 4, 12)
    scene = SimBackedScene(sd)
    integ = ep.load_dict({"type": variant, "max_depth": 6})
    img = integ.render(scene, sensor=1, seed=3, spp=4)
    res = sd.sensors[2].width
    g = torch.from_numpy(parity.synthetic_grad_in(res))
    gp, gn = integ.backward_buffers(scene, g, 2, ep.sample_tea_32(3, 1)[0], 8)
    if rank == 0:
        np.savez(out_path, img=img.numpy(), gp=gp.numpy(), gn=gn.numpy())
    if world > 1:
        # every rank holds the reduced result (allreduce, not reduce: each rank replays apply_transformation)
        ref = [torch.zeros_like(gp) for _ in range(world)]; dist.all_gather(ref, gp.contiguous())
        assert all(torch.equal(ref[0], r) for r in ref)
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); return s.getsockname()[1]


@pytest.mark.parametrize("variant", ["manifold", "manifold_caustic"])
def test_two_ranks_equal_one_rank(tmp_path, variant):
    one, two = str(tmp_path / "one.npz"), str(tmp_path / "two.npz")
    _run(0, 1, 0, variant, one)
    mp.spawn(_run, args=(2, _free_port(), variant, two), nprocs=2, join=True)
    a, b = np.load(one), np.load(two)
    assert np.abs(a["gp"]).max() > 0, "test scene produced no gradient"
    for k, tol in (("img", 1e-5), ("gp", 1e-5), ("gn", 1e-5)):
        scale = max(float(np.abs(a[k]).max()), 1e-12)
        assert float(np.abs(a[k] - b[k]).max()) <= tol * scale, k


def test_sample_shards_partition_every_pixel():
    import epsm_mitsuba3_b200 as ep
    for spp in (1, 3, 8, 32, 64):
        for world in (1, 2, 4, 8):
            cover = []
            for r in range(world):
                s0, s1 = ep.sample_shard(spp, r, world); cover += list(range(s0, s1))
            assert cover == list(range(spp)), (spp, world)
    # scanline bands: whole tile rows, contiguous, covering all lanes
    for h, world in ((128, 8), (30, 4), (5, 8)):
        edges = [ep.band_lanes(h, 16, 4, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == h * 16 * 4 and all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
