"""epsm_mitsuba3_b200 — B200-native implementation of the EPSM manifold-integrator hot path
(jkxing/EPSM_Mitsuba3: `manifold`, `manifold_caustic` and the `*_hybrid` schedules built on them).

Only what the path needs lives here: `csrc/` (CUDA kernels + the C ABI of include/epsm_b200.h), the
host-side mirror of the reference's integrator interface (`engine`), the seeded synthetic scene
generators (`scenes`), the hybrid schedule helper (`schedule`) the matcher that produces the image gradient (`matcher`) and OBJ / PLY mesh ingestion (`meshio`).
"""
from . import capi, scenes, meshio  # noqa: F401
from .engine import (EpsmError, Scene, SceneParameters, Integrator, PathIntegrator, EPSMIntegrator, ManifoldIntegrator,  # noqa: F401
                     ManifoldCausticIntegrator, load_dict, register_integrator, traverse, render,
                     sample_tea_32, band_lanes, sample_shard)
from .schedule import HybridSchedule, Phase  # noqa: F401
from .matcher import Matcher  # noqa: F401

__all__ = ["capi", "scenes", "meshio", "EpsmError", "Scene", "SceneParameters", "Integrator", "PathIntegrator", "EPSMIntegrator", "ManifoldIntegrator",
           "ManifoldCausticIntegrator", "load_dict", "register_integrator", "traverse", "render", "sample_tea_32",
           "band_lanes", "sample_shard", "HybridSchedule", "Phase", "Matcher"]
