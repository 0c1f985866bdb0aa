"""Times Matcher.match_Sinkhorn on the GPU at the reference's match resolutions (synthetic images)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import epsm_mitsuba3_b200 as ep
out = {}
for res in [int(a) for a in sys.argv[1:]] or [64, 128, 256]:
    rng = np.random.default_rng(res)
    a = torch.from_numpy(rng.uniform(0, 1, (res * res, 3)).astype(np.float32)).cuda()
    b = torch.from_numpy(rng.uniform(0, 1, (res * res, 3)).astype(np.float32)).cuda()
    m = ep.Matcher(res, "cuda:0")
    m.match_Sinkhorn(a, b)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); g = m.match_Sinkhorn(a, b); ts.append(time.perf_counter() - t0)
    n = res * res; pairs = 4.0 * n * n * (m.n_steps + 2)
    out[res] = {"ms": 1e3 * float(np.median(ts)), "n_points": n, "eps_steps": m.n_steps, "pair_terms": pairs,
                "Gpairs_per_s": pairs / float(np.median(ts)) / 1e9, "finite": bool(torch.isfinite(g).all())}
print(json.dumps(out))
