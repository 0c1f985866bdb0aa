"""Times Matcher.match_Sinkhorn on the GPU at the reference's match resolutions (synthetic images)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import epsm_mitsuba3_b200 as ep
out = {"reorder": os.environ.get("EPSM_SINK_REORDER", "1")}


def smooth(res, seed, shift):
    """Blobs on a gradient: the kind of image the renderer produces (the matcher's clouds are smooth functions of the pixel grid)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:res, 0:res] / (res - 1)
    img = np.stack([0.2 + 0.5 * xx, 0.3 + 0.4 * yy, 0.5 + 0.0 * xx], -1)
    for _ in range(6):
        cx, cy, r = rng.uniform(0.2, 0.8) + shift, rng.uniform(0.2, 0.8), rng.uniform(0.05, 0.15)
        img += rng.uniform(0.2, 0.6, 3) * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * r * r))[..., None]
    return torch.from_numpy(np.clip(img, 0, 1).reshape(-1, 3).astype(np.float32)).cuda()


for res in [int(a) for a in sys.argv[1:] if a.isdigit()] or [64, 128, 256]:
  for kind in ("smooth", "noise"):
    rng = np.random.default_rng(res)
    if kind == "noise":
        a = torch.from_numpy(rng.uniform(0, 1, (res * res, 3)).astype(np.float32)).cuda()
        b = torch.from_numpy(rng.uniform(0, 1, (res * res, 3)).astype(np.float32)).cuda()
    else:
        a, b = smooth(res, 1, 0.0), smooth(res, 1, 0.05)
    m = ep.Matcher(res, "cuda:0")
    m.match_Sinkhorn(a, b)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); g = m.match_Sinkhorn(a, b); ts.append(time.perf_counter() - t0)
    n = res * res; pairs = 4.0 * n * n * (m.n_steps + 2)
    out[f"{res}_{kind}"] = {"ms": 1e3 * float(np.median(ts)), "n_points": n, "eps_steps": m.n_steps, "pair_terms": pairs,
                "Gpairs_per_s": pairs / float(np.median(ts)) / 1e9, "finite": bool(torch.isfinite(g).all())}
print(json.dumps(out))
