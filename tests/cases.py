"""The parity-test scenes, shared by the GPU tests, the host-harness tests and the golden-vector generators.
name → (scene factory, integrator variant, max_depth).  `stage` is BASELINE.json configs[3] (6890-vertex articulated
mesh, manifold_caustic, max_depth 3: the shadow-term branch epsm.py:609-620) at test size."""
from epsm_mitsuba3_b200 import scenes as S

CASES = {
    "mirror": (lambda: S.mirror_plane_scene(32, 4, 24), "manifold", 6),
    "rough": (lambda: S.mirror_plane_scene(32, 4, 24, rough=True), "manifold", 6),
    "glass": (lambda: S.glass_sphere_caustic(32, 4, 24, subdiv=3), "manifold_caustic", 6),
    "cbox_d3": (lambda: S.cornell_box(32, 4, 24), "manifold", 3),
    "cbox_mirror": (lambda: S.cornell_box(32, 4, 24, box_bsdf=S.Bsdf.conductor()), "manifold", 6),
    "cbox_glass": (lambda: S.cornell_box(32, 4, 24, box_bsdf=S.Bsdf.dielectric(1.5)), "manifold_caustic", 5),
    "room": (lambda: S.indoor_room(20000, 32, 4, 24), "manifold", 6),
    "stage": (lambda: S.articulated_stage(32, 4, 24), "manifold_caustic", 3),
}
