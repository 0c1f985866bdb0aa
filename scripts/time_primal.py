"""Times the wavefront primal render of the room at BASELINE size (median of 3 after a warm-up)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import epsm_mitsuba3_b200 as ep
from epsm_mitsuba3_b200 import scenes as S
sd = S.indoor_room(1000000, 1024, 32, 1024); scene = ep.Scene(sd, device=0); integ = ep.load_dict({"type": "manifold", "max_depth": 8})
ts = []
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter(); integ.render(scene, sensor=1, seed=it, spp=32); scene.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("primal ms", round(float(np.median(ts[1:])), 2))
