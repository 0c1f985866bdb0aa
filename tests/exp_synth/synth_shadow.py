"""Synthetic experiment module in the layout EPSM/optim.py expects of `exp.<config>` (optim.py:42-49,69): the same kind of
task as EPSM/exp/shadow.py (a ring of small spheres casting shadows on a floor, optimised by per-sphere translations), cut
down to test size: 64² film, 8 spheres, 4 iterations, hybrid switch at iteration 2.  The mesh file is written by the test."""
import math

import drjit as dr
import mitsuba as mi
from mitsuba.scalar_rgb import Transform4f as T

it = 4
spp = 4
resolution = 64
thres = 2
max_depth = 2
match_res = 32
n_spheres = 8


def _sensor(border, res, count):
    return {
        "type": "perspective", "fov": 90, "near_clip": 1.0, "far_clip": 1000.0,
        "to_world": mi.ScalarTransform4f.look_at(origin=[0.1, 3.5, 0.1], target=[0.1, 0, 0.1], up=[0, 0, 1]),
        "myfilm": {"type": "hdrfilm", "rfilter": {"type": "gaussian"}, "sample_border": border, "width": res, "height": res},
        "mysampler": {"type": "independent", "sample_count": count},
    }


def _scene(offset):
    base = {
        "type": "scene",
        "integrator": {"type": "prb_reparam", "max_depth": max_depth},
        "sensor0": _sensor(True, resolution, 4),
        "sensor1": _sensor(False, resolution, 4),
        "sensor2": _sensor(False, match_res, 8),
        "DiffuseBSDF": {"type": "diffuse", "reflectance": {"type": "rgb", "value": [0.5, 0.5, 0.5]}},
        "DiffuseBSDF2": {"type": "diffuse", "reflectance": {"type": "rgb", "value": [0.8, 0.5, 0.5]}},
        "myemitter": {"type": "rectangle", "to_world": T.translate([0.0, 8.0, 0.0]).rotate([1, 0, 0], 90.0).scale(0.01),
                      "emitter": {"type": "area", "radiance": {"type": "rgb", "value": 300000.0}}},
        "Floor": {"type": "rectangle", "to_world": T.rotate([1, 0, 0], -90.0).scale(10), "mybsdf": {"type": "ref", "id": "DiffuseBSDF"}},
    }
    for i in range(n_spheres):
        a = i / n_spheres * math.pi * 2
        base[f"sphere{i}"] = {"type": "obj", "filename": "data/meshes/sphere.obj",
                              "to_world": T.translate([math.sin(a) * 0.8 + 0.03 + offset, 6.0, math.cos(a) * 0.8 + 0.03]).scale(0.12),
                              "mybsdf": {"type": "ref", "id": "DiffuseBSDF2"}}
    return mi.load_dict(base)


gt_scene = _scene(0.25)
scene = _scene(0.0)
params = mi.traverse(scene)


def optim_settings():
    objlist = [f"sphere{i}" for i in range(n_spheres)]
    opt = mi.ad.Adam(lr=0.02)
    init = {}
    for obj in objlist:
        opt[obj] = mi.Vector3f([0.0, 0.0, 0.0])
        init[obj] = dr.unravel(mi.Point3f, params[f"{obj}.vertex_positions"])

    def apply_transformation(params, opt):
        for obj in objlist:
            opt[obj][1] = 0
            trans_mat = mi.Transform4f.translate(opt[obj])
            params[f"{obj}.vertex_positions"] = dr.ravel(trans_mat @ init[obj])
        params.update()

    def output(opt):
        return opt["sphere0"]

    return opt, apply_transformation, output, params
