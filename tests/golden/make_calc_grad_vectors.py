"""Pins the gradient oracle to THE REFERENCE'S OWN calc_grad CODE, executed here.

`ManifoldIntegrator.calc_grad` (src/python/python/ad/integrators/epsm.py:744-946) and
`ManifoldCausticIntegrator.calc_grad` (epsm.py:951-1200) need only torch: this script reads those source
ranges from /root/reference AT RUN TIME (nothing is copied into the repository), `exec`s them with
  * `torch.Tensor.cuda` → identity (the reference hard-codes `.cuda()`, epsm.py:768-773),
  * `mi.has_flag(flags, mi.BSDFFlags.X)` → object with `.torch()` (integer and), `mi.Point3f` → identity,
  * `EPSMIntegrator` → object,
feeds them the `path_info` list (layout of epsm.py:547,648-654) built from the oracle's path log of the
parity-test scenes, and stores what the reference returns — in float32 (the reference's precision) and in
float64 (default dtype switched) — in tests/golden/calc_grad_reference.npz together with the inputs.

tests/test_calc_grad_pinned.py holds oracle/calc_grad_ref.py to these vectors (and re-executes the reference
live whenever /root/reference is present, so the fixture stays regenerable); the `-m gpu` test holds the CUDA
calc_grad to them directly.

    python tests/golden/make_calc_grad_vectors.py        # rewrites the fixture
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
EPSM_PY = os.environ.get("EPSM_REFERENCE", "/root/reference") + "/src/python/python/ad/integrators/epsm.py"
FIXTURE = os.path.join(ROOT, "tests", "golden", "calc_grad_reference.npz")
WINDOW = 512                     # three windows of WINDOW lanes of the gradient sensor's wavefront per case
WINDOW_AT = (0.0, 0.45, 0.8)     # at these fractions of the wavefront (top, middle, lower part of the image)

BF_NULL, BF_DIFFUSE = 0x1, 0x6   # include/mitsuba/render/bsdf.h:37-110


def reference_available() -> bool:
    return os.path.isfile(EPSM_PY)


class _Flag:
    def __init__(self, t):
        self.t = t

    def torch(self):
        return self.t


class _MiStub:
    """The four `mi.` symbols calc_grad touches."""
    class BSDFFlags:
        Null = BF_NULL
        Diffuse = BF_DIFFUSE

    @staticmethod
    def has_flag(flags, flag):
        return _Flag((flags & int(flag)) != 0)

    @staticmethod
    def Point3f(t):
        return t

    @staticmethod
    def register_integrator(*a, **k):
        pass


def _load_reference_classes():
    """exec the two class bodies from the reference file (line ranges located by their class statements)."""
    with open(EPSM_PY) as f:
        lines = f.read().split("\n")
    out = {}
    for cls in ("ManifoldIntegrator", "ManifoldCausticIntegrator"):
        start = next(i for i, l in enumerate(lines) if l.startswith(f"class {cls}(EPSMIntegrator):"))
        end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith("mi.register_integrator("))
        ns = {"torch": torch, "mi": _MiStub, "EPSMIntegrator": object, "dr": None}
        exec(compile("\n" * start + "\n".join(lines[start:end]), EPSM_PY, "exec"), ns)    # keeps the reference's line numbers
        out[cls] = ns[cls]
    return out


def path_info_from_log(paths: np.ndarray, nl: int, dtype):
    """The list `sample_path(log_path=True)` returns (epsm.py:547,648-654) from the oracle's structured path log."""
    f = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=dtype)
    info = [{"cam": f(paths["cam"])}]
    for k in range(nl):
        v = paths["v"][:, k]
        info.append({
            "it": k, "active": torch.tensor(v["active"] > 0), "bsdf": torch.tensor(v["bsdf_flags"].astype(np.int64)),
            "ismesh": f(v["ismesh"]), "light": f(v["light"]), "active_em": torch.tensor(v["active_em"] > 0),
            "points": [f(v["p0"]), f(v["p1"]), f(v["p2"]), f(v["p"])], "uv": [f(v["b0"]), f(v["b1"])],
            "normal": f(v["sh_n"]), "normals": [f(v["n0"]), f(v["n1"]), f(v["n2"])], "eta": f(v["eta"]), "hf": f(v["hf"]),
        })
    return info


def run_reference(paths: np.ndarray, variant: str, nl: int, dtype=torch.float32):
    """The reference's calc_grad on CPU torch.  Returns (path_grad, light_grad, diffuse_grad) as lists of numpy arrays."""
    cls = _load_reference_classes()["ManifoldCausticIntegrator" if variant == "manifold_caustic" else "ManifoldIntegrator"]
    old_cuda, old_dtype = torch.Tensor.cuda, torch.get_default_dtype()
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.set_default_dtype(dtype)
    try:
        info = path_info_from_log(paths, nl, dtype)
        n = paths.shape[0]
        dlduv = torch.zeros((n, 1, 2 * len(info)), dtype=dtype)                       # epsm.py:256
        dlduv[:, 0, 0:2] = torch.tensor(np.ascontiguousarray(paths["dlduv"]), dtype=dtype)   # epsm.py:268-269
        dldp1 = torch.tensor(np.ascontiguousarray(paths["dldp1"]), dtype=dtype)        # epsm.py:270
        Lt = torch.tensor(np.ascontiguousarray(paths["L"]), dtype=dtype).sum(-1)       # epsm.py:273-274
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")                                            # torch.cross without dim (epsm.py:754)
            pg, lg, dg = cls.calc_grad(object.__new__(cls), path_info=info, dlduv=dlduv, dldp=dldp1, Lt=Lt)
    finally:
        torch.Tensor.cuda = old_cuda
        torch.set_default_dtype(old_dtype)
    cv = lambda l: [t.detach().numpy().copy() for t in l]
    return cv(pg), cv(lg), cv(dg)


def cases():
    """The parity-test scenes (tests/test_gpu_parity.py CASES) — kept in one place: tests/cases.py."""
    sys.path.insert(0, ROOT)
    from tests.cases import CASES
    return CASES


def case_paths(name):
    from oracle import oracle
    from tests import parity
    make, variant, max_depth = cases()[name]
    sd = make()
    o = oracle.OracleScene(sd)
    s2 = sd.sensors[2]
    g = parity.synthetic_grad_in(s2.width)
    total = s2.width * s2.height * 8
    begins = [min(total - WINDOW, int(f * total) // 8 * 8) for f in WINDOW_AT]
    o.set_accel(True)
    paths = np.concatenate([o.path_log(2, 11, 8, max_depth, 5, g, b, b + WINDOW) for b in begins])
    return paths, variant, parity.n_logged(max_depth), begins


def main():
    assert reference_available(), f"{EPSM_PY} not found"
    out = {}
    for name in sorted(cases()):
        paths, variant, nl, begins = case_paths(name)
        out[f"{name}/paths"] = paths.view(np.uint8)           # raw bytes of the structured array (dtype = oracle.path_dtype())
        out[f"{name}/begins"] = np.array(begins, np.int64)
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            pg, lg, dg = run_reference(paths, variant, nl, dt)
            out[f"{name}/{tag}/pg"] = np.stack(pg); out[f"{name}/{tag}/lg"] = np.stack(lg); out[f"{name}/{tag}/dg"] = np.stack(dg)
        nz = sum(int((a != 0).any(1).sum()) for a in pg + lg + dg)
        print(f"{name:12s} {variant:17s} nl={nl} paths={paths.shape[0]} lists={len(pg)}/{len(lg)}/{len(dg)} non-zero rows={nz}")
    np.savez_compressed(FIXTURE, **out)
    print("wrote", FIXTURE, os.path.getsize(FIXTURE) // 1024, "KiB")


if __name__ == "__main__":
    main()
