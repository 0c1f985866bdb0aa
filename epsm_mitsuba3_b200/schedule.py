"""The `*_hybrid` schedule of the reference's optimisation driver (EPSM/optim.py:87-119), host logic only.

`METHOD` strings ending in `hybrid` strip the 7-character suffix (`manifold_hybrid` → `manifold`,
`manifold_caustic_hybrid` → `manifold_caustic`), run the EPSM integrator on sensor 1 for iterations `it < thres`,
then reset the optimiser state once and switch to `prb_reparam` (engine.PRBReparamIntegrator over
`epsm_prb_reparam_backward`) on sensor 0 with an L2 image loss.  examples/ and bench.py's `iteration` leg drive their loops
through this class; EPSM/optim.py itself runs unmodified through the stand-in modules (epsm_mitsuba3_b200/shim).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

from .engine import EPSMIntegrator, load_dict

NO_SWITCH = 10000                                  # optim.py:95: `thres = 10000` for non-hybrid methods


@dataclass
class Phase:
    integrator: Optional[EPSMIntegrator]
    integrator_type: str
    sensor: int                                    # optim.py:103-106: EPSM integrators render sensor 1, PRB sensor 0
    loss: str                                      # 'matcher' (5-channel image, optim.py:130-136) or 'l2' (optim.py:137-141)
    reset_optimizer: bool                          # optim.py:115-117: opt.reset(key) for every key when it == thres


class HybridSchedule:
    """Resolves (method, max_depth, thres) the way optim.py does and answers `phase(it)` per iteration."""

    def __init__(self, method: str, max_depth: int, thres: int = NO_SWITCH,
                 phase2_factory: Optional[Callable[[dict], object]] = None):
        self.hybrid = method.endswith("hybrid")
        if self.hybrid:
            method = method[:-7]                   # optim.py:88
            self.thres = int(thres)
        else:
            self.thres = NO_SWITCH
        self.method = method
        self.max_depth = int(max_depth)
        self.integrator1 = load_dict({"type": method, "max_depth": self.max_depth})     # optim.py:97-100
        self.sensor1 = 1 if method.startswith("manifold") else 0                         # optim.py:103-106
        self._phase2_factory = phase2_factory
        self._integrator2 = None

    def _phase2(self):
        if self._integrator2 is None:
            props = {"type": "prb_reparam", "max_depth": self.max_depth}                 # optim.py:89-92
            self._integrator2 = self._phase2_factory(props) if self._phase2_factory else load_dict(props)
        return self._integrator2

    def phase(self, it: int) -> Phase:
        if it < self.thres:
            return Phase(self.integrator1, self.method, self.sensor1, "matcher", False)
        return Phase(self._phase2(), "prb_reparam", 0, "l2", it == self.thres)
