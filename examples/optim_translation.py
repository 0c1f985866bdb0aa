"""The loop of EPSM/optim.py:87-157 on this package's API: recover the translation of one mesh from a target image with the
matcher's pixel-motion gradient and the manifold integrators' vertex gradients.

  python examples/optim_translation.py [--scene mirror|glass|smpl] [--iters 40] [--res 128] [--spp 16] [--match-res 64]

Per iteration (optim.py:110-157): apply the parameters → params.update() (device BVH refit) → img = render(scene, params,
sensor=1, integrator, seed=it, spp) → low-resolution clouds → Matcher.match_Sinkhorn → grad (tiled to the film, optim.py:134)
→ backward of sum(img · grad) (render_backward: sensor 2, 8 spp, crop to match_res, epsm.py:142-145,238-240) → Adam step.
Prints one JSON line per iteration and a summary; exits non-zero if the parameter error did not shrink."""
import argparse, json, os, sys, time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F

import epsm_mitsuba3_b200 as ep
from epsm_mitsuba3_b200 import scenes as S


def run(scene_name="mirror", iters=40, res=128, spp=16, match_res=64, lr=0.02, offset=(0.25, -0.15, 0.0), quiet=False, device=0):
    per_vertex = False
    if scene_name == "glass":
        sd, mesh, variant, md = S.glass_sphere_caustic(res, spp, match_res, subdiv=4), "glass", "manifold_caustic", 6
    elif scene_name == "smpl":
        # BASELINE configs[3]: 6890-vertex mesh in the human.py stage, per-vertex offsets, Adam lr 0.01, max_depth 3
        # (phase 1 of manifold_caustic_hybrid; phase 2, prb_reparam, is not built)
        sd, variant, md, per_vertex = S.articulated_stage(res, spp, match_res), "manifold_caustic", 3, True
        mesh = max((m for m in sd.meshes if m.grad), key=lambda m: m.positions.shape[0]).name
    else:
        sd, mesh, variant, md = S.mirror_plane_scene(res, spp, match_res), "wall", "manifold", 6
    scene = ep.Scene(sd, device=device)
    params = ep.traverse(scene)
    key = f"{mesh}.vertex_positions"
    integ = ep.load_dict({"type": variant, "max_depth": md})
    matcher = ep.Matcher(match_res, f"cuda:{device}")
    base = params[key].clone()

    def low(img):                                              # optim.py:130: resize to match_res (area average here)
        return F.adaptive_avg_pool2d(img[..., :3].permute(2, 0, 1)[None], match_res)[0].permute(1, 2, 0).reshape(-1, 3)

    target = low(integ.render(scene, sensor=1, seed=12345, spp=4 * spp)).clone()
    if per_vertex:
        g0 = torch.Generator(device="cpu").manual_seed(0)
        theta = (0.02 * torch.randn(base.shape, generator=g0)).to(scene.device).requires_grad_(True)
    else:
        theta = torch.tensor(offset, dtype=torch.float32, device=scene.device, requires_grad=True)
    opt = torch.optim.Adam([theta], lr=lr)
    hist = []
    for it in range(iters):
        t0 = time.perf_counter()
        params[key] = base + theta                              # apply_transformation (optim.py:110)
        params.update()
        img = ep.render(scene, params, sensor=1, integrator=integ, seed=it, spp=spp)
        g = matcher.match_Sinkhorn(low(img.detach()), target, want_loss=True).reshape(match_res, match_res, 5)
        grad = g.repeat(res // match_res, res // match_res, 1)  # optim.py:134 (tiled; render_backward crops to match_res)
        opt.zero_grad()
        (img * grad).sum().backward()                           # dr.backward(img * grad), optim.py:136
        if theta.grad is not None:
            theta.grad = torch.nan_to_num(theta.grad)           # optim.py:146-150
        opt.step()
        torch.cuda.synchronize()
        rec = {"it": it, "sinkhorn": matcher.last_loss, "err": float(theta.detach().norm()),
               "theta": None if per_vertex else [round(float(x), 4) for x in theta.detach()], "ms": (time.perf_counter() - t0) * 1e3}
        hist.append(rec)
        if not quiet:
            print(json.dumps(rec))
    return hist


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="mirror", choices=["glass", "mirror", "smpl"])
    ap.add_argument("--iters", type=int, default=40); ap.add_argument("--res", type=int, default=128)
    ap.add_argument("--spp", type=int, default=16); ap.add_argument("--match-res", type=int, default=64)
    ap.add_argument("--lr", type=float, default=0.02)
    ap.add_argument("--offset", type=float, nargs=3, default=[0.25, -0.15, 0.0])
    a = ap.parse_args()
    h = run(a.scene, a.iters, a.res, a.spp, a.match_res, a.lr, tuple(a.offset))
    e0, e1 = (h[0]["err"] if a.scene == "smpl" else float(np.linalg.norm(a.offset))), float(np.mean([r["err"] for r in h[-5:]]))
    print(json.dumps({"summary": True, "err_start": e0, "err_end": e1, "sinkhorn_start": h[0]["sinkhorn"],
                      "sinkhorn_end": float(np.mean([r["sinkhorn"] for r in h[-5:]])), "ms_per_iteration": float(np.median([r["ms"] for r in h[3:]]))}))
    sys.exit(0 if (e1 < 0.5 * e0 or a.scene == "smpl") else 1)
