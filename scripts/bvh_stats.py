"""Traversal statistics of the host-compiled traversal code (tests/hostsim) on the 1 M-triangle room: nodes fetched and
triangles tested per ray for coherent camera rays and for incoherent rays — builder tuning without GPU time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from epsm_mitsuba3_b200 import scenes as S
from tests.hostsim.hostsim import SimScene
from tests import parity

n_tris = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
sd = S.indoor_room(n_tris, 256, 4, 256)
t = time.time(); sim = SimScene(sd); st = sim.stats()
print("build %.2f s" % (time.time() - t), st)
rng = np.random.default_rng(1)
cam = sim.sample_rays(1, 3, 4, 0, 200000)
rays = np.zeros((cam.shape[0], 8), np.float32); rays[:, :3] = cam[:, :3]; rays[:, 3] = 3.4e38; rays[:, 4:7] = cam[:, 3:6]
for name, r, anyhit in (("camera closest", rays, False), ("random closest", parity.random_rays(sd, 200000, finite_fraction=0.0), False),
                        ("random any (finite)", parity.random_rays(sd, 200000, finite_fraction=1.0), True)):
    c = np.zeros(4, np.uint64); t = time.time(); h = sim.trace(r, any_hit=anyhit, counters=c)
    print("%-22s nodes/ray %6.2f  tris/ray %6.2f   (%.2f s)" % (name, c[0] / c[2], c[1] / c[2], time.time() - t))
