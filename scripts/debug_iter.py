"""Debug helper: primal render / refit / gradient pass sequence of bench.iteration_leg on a reduced room."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import epsm_mitsuba3_b200 as ep
from epsm_mitsuba3_b200 import scenes as S
tris, res, spp = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sd = S.indoor_room(tris, res, spp, res)
scene = ep.Scene(sd, device=0)
integ = ep.load_dict({"type": "manifold", "max_depth": 8})
print("scene up", scene.stats(), flush=True)
img = integ.render(scene, sensor=1, seed=0, spp=spp); scene.synchronize(); torch.cuda.synchronize()
print("primal ok", float(img[..., :3].mean()), flush=True)
if len(sys.argv) > 4 and sys.argv[4] == "primal":
    sys.exit(0)
params = ep.traverse(scene)
keys = [k for k, i in params._mesh_of.items() if sd.meshes[i].grad] or list(params._mesh_of)
print("grad meshes", len(keys), flush=True)
g = torch.from_numpy(np.random.default_rng(0).normal(0, 0.02, (res, res, 5)).astype(np.float32)).cuda()
buf = torch.zeros(2, scene.n_vertices, 3, device="cuda"); torch.cuda.synchronize()
for mode in ("refit", "rebuild"):
    scene.accel_update = mode
    for it in range(2):
        for k in keys:
            params[k] = params[k] + 1e-3
        params.update(); scene.synchronize(); print(mode, it, "update ok", scene.accel_cost(), flush=True)
        img = integ.render(scene, sensor=1, seed=it, spp=spp); scene.synchronize(); torch.cuda.synchronize()
        print(mode, it, "primal ok", float(img[..., :3].mean()), flush=True)
        buf.zero_(); torch.cuda.synchronize()
        scene.render_backward_raw("manifold", 1, 5 + it, spp, 8, 5, g, buf[0], buf[1]); scene.synchronize()
        print(mode, it, "backward ok", float(buf.abs().sum()), flush=True)
