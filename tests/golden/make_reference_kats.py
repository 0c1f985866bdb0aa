"""How tests/golden/reference_kats.json was produced.

The values are NOT computed by this repository.  They come from two sources inside the reference:

1. Literal asserts in the reference's own tests — copied as data:
     src/core/tests/test_random.py:8-27           (TEA float32 / float64)
     src/render/tests/test_mesh.py:257-291        (2-triangle rectangle hit record)
2. Values printed by the reference's `scalar_rgb` Python module, built out-of-tree from
   /root/reference during the survey.  With that build on PYTHONPATH
   this script re-reads them; without it (the normal case: neither mitsuba nor drjit is installed in
   this image and the GPU box has no /root/reference) it only prints the committed file.

    PYTHONPATH=<build>/python LD_LIBRARY_PATH=<build>:<build>/python/drjit python tests/golden/make_reference_kats.py
"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def from_reference():
    import drjit as dr          # noqa: F401  (reference build only)
    import mitsuba as mi
    mi.set_variant("scalar_rgb")
    out = {}
    out["tea_float32"] = [[a, b, mi.sample_tea_float32(a, b, 4)] for a, b in
                          [(1, 1), (1, 2), (1, 3), (1, 4), (1, 5), (2, 1), (3, 1), (4, 1)]]
    s = mi.load_dict({"type": "independent"}); s.seed(0)
    out["pcg32_seed0"] = [float(s.next_1d()) for _ in range(3)]
    scene = mi.load_dict(mi.cornell_box())
    ray = mi.Ray3f([0.3, -0.5, 3.9], [0, 0, -1])
    pi = scene.ray_intersect_preliminary(ray)
    si = pi.compute_surface_interaction(ray)
    out["ray_mesh"] = dict(t=float(pi.t), prim=int(pi.prim_index), prim_uv=list(pi.prim_uv), b0=float(si.b0),
                           b1=float(si.b1), p0=list(si.p0), n0=list(si.n0), ismesh=int(si.ismesh))
    ray2 = mi.Ray3f([0, 0.9, 3.9], [0, 0, -1])
    si2 = scene.ray_intersect(ray2)
    ds, _ = scene.sample_emitter_direction(si2, [0.2, 0.6], True)
    out["sample_emitter_direction"] = dict(p=list(ds.p), pdf=float(ds.pdf))
    r, _ = scene.sensors()[0].sample_ray_differential(0, 0, [0.25, 0.75], [0, 0])
    out["sample_ray_differential"] = dict(o=list(r.o), d=list(r.d), d_x=list(r.d_x), maxt=float(r.maxt))
    return out


if __name__ == "__main__":
    try:
        print(json.dumps(from_reference(), indent=1))
    except ImportError:
        print("reference module not importable here; committed values:")
        print(open(os.path.join(HERE, "reference_kats.json")).read())
