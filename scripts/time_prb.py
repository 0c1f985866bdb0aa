"""Times the prb_reparam reverse pass (k_prb_backward) on the room scene: scripts/time_prb.py [res] [spp]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import epsm_mitsuba3_b200 as ep
from epsm_mitsuba3_b200 import scenes as S
res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sd = S.indoor_room(1_000_000, res, spp, res)
scene = ep.Scene(sd)
integ = ep.load_dict({"type": "prb_reparam", "max_depth": 8})
g = torch.from_numpy(np.random.default_rng(1).normal(0, 0.1, (res, res, 3)).astype(np.float32)).cuda()
for k in range(int(sys.argv[3]) if len(sys.argv) > 3 else 3):
    torch.cuda.synchronize(); t = time.perf_counter()
    gp, gn = integ.vertex_gradients(scene, g, 0, k, spp)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"lib={os.environ.get('EPSM_B200_LIB','default')} res={res} spp={spp} prb backward {dt*1e3:.1f} ms  ({(res+4)**2*spp/dt/1e6:.2f} M lanes/s)  |g|={float(gp.abs().sum()):.4g}")
