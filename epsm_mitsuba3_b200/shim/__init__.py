"""Stand-in `mitsuba`, `drjit` and `geomloss` modules over the B200 path, so that the reference's drivers
(EPSM/optim.py, EPSM/exp/*.py, EPSM/utils/matcher.py) run UNMODIFIED against it (BASELINE.json north_star; SURVEY.md §8b
"Python surface" row).  The image has no Mitsuba, Dr.Jit or geomloss: these modules serve the symbols the
drivers use — arrays are thin wrappers over torch tensors (torch is the tensor boundary of optim.py anyway), autograd is
torch's, and everything that renders goes through libepsm_b200.so.

    import epsm_mitsuba3_b200.shim as shim; shim.install()      # puts this directory first on sys.path
    import mitsuba as mi, drjit as dr                             # … the reference's scripts from here on

or  PYTHONPATH=<repo>/epsm_mitsuba3_b200/shim python optim.py manifold_hybrid <exp>
"""
import os
import sys

SHIM_DIR = os.path.dirname(os.path.abspath(__file__))


def install():
    """Make `import mitsuba`, `import drjit`, `import geomloss` resolve to the stand-ins (idempotent)."""
    root = os.path.dirname(os.path.dirname(SHIM_DIR))
    if root not in sys.path:
        sys.path.insert(0, root)
    if SHIM_DIR in sys.path:
        sys.path.remove(SHIM_DIR)
    sys.path.insert(0, SHIM_DIR)
    for name in ("mitsuba", "drjit", "geomloss"):
        m = sys.modules.get(name)
        if m is not None and not getattr(m, "__file__", "").startswith(SHIM_DIR):
            raise RuntimeError(f"another `{name}` module is already imported from {getattr(m, '__file__', '?')}")
    return SHIM_DIR
