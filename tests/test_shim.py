"""The `mitsuba` / `drjit` / `geomloss` stand-ins (epsm_mitsuba3_b200/shim): host logic that needs no GPU.
What the reference's drivers do with these symbols (EPSM/optim.py:18-157, EPSM/exp/shadow.py:98-225) is replayed here on
CPU torch; everything that renders is in tests/test_shim_optim.py (-m gpu)."""
import math
import os
import sys

import numpy as np
import pytest
import torch

import epsm_mitsuba3_b200.shim as shim


@pytest.fixture(scope="module")
def mods():
    shim.install()
    import drjit as dr
    import mitsuba as mi
    mi.set_variant("cuda_ad_rgb")
    return mi, dr


def test_modules_resolve_to_the_stand_ins(mods):
    mi, dr = mods
    assert mi.__file__.startswith(shim.SHIM_DIR) and dr.__file__.startswith(shim.SHIM_DIR)
    from mitsuba.scalar_rgb import Transform4f as T       # optim.py:16
    assert T is mi.Transform4f
    assert mi.variant() == "cuda_ad_rgb"
    with pytest.raises(ImportError):
        mi.set_variant("scalar_spectral_polarized")
    import geomloss
    assert geomloss.__file__.startswith(shim.SHIM_DIR)
    with pytest.raises(NotImplementedError):
        geomloss.SamplesLoss("gaussian")


def test_transform_chain_matches_scalar_transform(mods):
    """T.translate(a).rotate(axis, deg).scale(s) right-multiplies, like include/mitsuba/core/transform.h; compared with
    the float64 host transform the scene generators use."""
    mi, _ = mods
    from epsm_mitsuba3_b200.scenes import Transform4f as H
    t = mi.Transform4f.translate([0.05, 6.0, 0.05]).scale(0.7).rotate([0, 1, 0], 180)          # exp/shadow.py:88
    h = H().translate([0.05, 6.0, 0.05]).scale(0.7).rotate([0, 1, 0], 180)
    np.testing.assert_allclose(t.numpy(), h.matrix, atol=1e-6)
    la = mi.ScalarTransform4f.look_at(origin=[0.1, 3.5, 0.1], target=[0.1, 0, 0.1], up=[0, 0, 1])   # exp/shadow.py:34-36
    np.testing.assert_allclose(la.numpy(), H.look_at([0.1, 3.5, 0.1], [0.1, 0, 0.1], [0, 0, 1]).matrix, atol=1e-6)
    p = mi.Point3f([[1.0, 2.0, 3.0], [0.0, 0.0, 0.0]])
    np.testing.assert_allclose((t @ p).numpy(), h.apply_points(p.numpy()), atol=1e-5)
    v = mi.Vector3f([[1.0, 2.0, 3.0]])
    np.testing.assert_allclose((t @ v).numpy(), np.asarray(v.numpy(), np.float64) @ h.matrix[:3, :3].T, atol=1e-5)
    comp = t @ la
    np.testing.assert_allclose(comp.numpy(), h.matrix @ H.look_at([0.1, 3.5, 0.1], [0.1, 0, 0.1], [0, 0, 1]).matrix, atol=1e-5)


def test_gradients_flow_from_vertices_to_the_optimised_variable(mods):
    """apply_transformation of exp/shadow.py:213-220: opt[obj][1] = 0; ravel(translate(opt[obj]) @ init) — the vertex buffer
    must carry the autograd graph back to opt[obj]."""
    mi, dr = mods
    opt = mi.ad.Adam(lr=0.005)
    opt["s"] = mi.Vector3f([0.1, 0.2, 0.3])
    dr.enable_grad(opt["s"])
    init = dr.unravel(mi.Point3f, mi.Float(np.arange(12, dtype=np.float32)))
    opt["s"][1] = 0
    flat = dr.ravel(mi.Transform4f.translate(opt["s"]) @ init)
    assert flat.shape == (12,)
    np.testing.assert_allclose(flat.numpy().reshape(4, 3), np.arange(12).reshape(4, 3) + [0.1, 0.0, 0.3], atol=1e-6)
    w = mi.TensorXf(torch.arange(12, dtype=torch.float32))
    dr.backward(flat * w)
    g = dr.grad(opt["s"]).numpy().reshape(3)
    np.testing.assert_allclose(g, [0 + 3 + 6 + 9, 1 + 4 + 7 + 10, 2 + 5 + 8 + 11])
    # optim.py:139-147: x[dr.isnan(x)] = 0; dr.set_grad
    x = dr.grad(opt["s"]); x.t[0, 1] = float("nan")
    x[dr.isnan(x)] = 0
    dr.set_grad(opt["s"], x)
    assert dr.grad(opt["s"]).numpy().reshape(3)[1] == 0
    # rotation angle and look_at origin given as differentiable Floats (exp/bedroom.py:27, exp/cornellbox.py:116-124)
    a = mi.Float(30.0); dr.enable_grad(a)
    r = mi.Transform4f.rotate([0, 1, 0], a * 2) @ mi.Point3f([[1.0, 0.0, 0.0]])
    dr.backward(r)
    th = math.radians(60)
    assert abs(float(dr.grad(a)) - 2 * math.pi / 180 * (-math.sin(th) - math.cos(th))) < 1e-5


def test_adam_follows_the_reference_update(mods):
    """optimizers.py:236-268: lr_t = lr·sqrt(1−β2^t)/(1−β1^t); step = lr_t·m/(sqrt(v)+ε); reset() zeroes state and t."""
    mi, dr = mods
    opt = mi.ad.Adam(lr=0.01)
    opt["x"] = mi.Float([1.0, -2.0, 3.0])
    x = np.array([1.0, -2.0, 3.0]); m = np.zeros(3); v = np.zeros(3)
    rng = np.random.default_rng(0)
    for t in range(1, 6):
        g = rng.normal(size=3)
        dr.set_grad(opt["x"], mi.Float(g.astype(np.float32)))
        opt.step()
        m = 0.9 * m + 0.1 * g; v = 0.999 * v + 0.001 * g * g
        x = x - 0.01 * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * m / (np.sqrt(v) + 1e-8)
        np.testing.assert_allclose(opt["x"].numpy(), x, rtol=1e-5)
        assert opt["x"].t.requires_grad and opt["x"].t.grad is None
    opt.reset("x")                                   # optim.py:117-119 at it == thres
    assert opt.t["x"] == 0 and float(opt.state["x"][0].abs().sum()) == 0
    opt["x"] = dr.clamp(opt["x"], -1, 1)             # exp/glassslab.py:270: re-assignment keeps the state
    assert opt["x"].numpy().max() <= 1 and opt["x"].t.is_leaf
    with pytest.raises(Exception):
        opt["y"] = 3.0
    keys = list(opt.keys())
    assert keys == ["x"]


def test_tensor_semantics_used_by_the_drivers(mods):
    mi, dr = mods
    img = mi.TensorXf(torch.rand(4, 6, 5))
    assert img.shape == (4, 6, 5) and len(img) == 120                      # optim.py:138 divides by len(img) = #entries
    assert img[..., :3].shape == (4, 6, 3)
    ref = mi.TensorXf(torch.rand(4, 6, 3))
    loss = dr.sum(dr.sqr(img[..., :3] - ref)) / len(img)
    assert loss.shape == (1,)
    b = mi.util.convert_to_bitmap(mi.TensorXf(torch.tensor([[[0.0, 0.0031308, 1.5], [0.5, 0.2, 0.75]]])))
    a = np.array(b)
    assert a.dtype == np.uint8 and a.shape == (1, 2, 3)
    srgb = lambda x: 12.92 * x if x <= 0.0031308 else 1.055 * x ** (1 / 2.4) - 0.055
    assert a[0, 0].tolist() == [0, int(srgb(0.0031308) * 255 + 0.5), 255]
    assert a[0, 1].tolist() == [int(srgb(c) * 255 + 0.5) for c in (0.5, 0.2, 0.75)]
    U, V = dr.meshgrid(dr.linspace(mi.Float, 0, 1, 3), dr.linspace(mi.Float, 0, 1, 4), indexing="ij")     # exp/glassslab.py:25-29
    assert U.shape == (12,) and float(V[1]) == pytest.approx(1 / 3)
    v = dr.zeros(mi.cuda_ad_rgb.Vector3f, 5)                                 # exp/glassslab.py:253-257
    v[1, 3] = 0.25
    assert v.numpy()[3].tolist() == [0, 0.25, 0]


def _write_obj(path, v, f):
    with open(path, "w") as fh:
        for p in v:
            fh.write("v %f %f %f\n" % tuple(p))
        for t in f:
            fh.write("f %d %d %d\n" % tuple(int(i) + 1 for i in t))


def test_load_dict_of_a_driver_style_scene(mods, tmp_path):
    """The dict layout of exp/shadow.py:98-203 (sensor trio, named BSDFs + refs, rectangle emitter, obj shapes)."""
    mi, dr = mods
    from epsm_mitsuba3_b200.scenes import icosphere
    v, f = icosphere(1)
    _write_obj(tmp_path / "sphere.obj", v, f)
    T = mi.Transform4f

    def sensor(border, res, spp):
        return {"type": "perspective", "fov": 90, "near_clip": 1.0, "far_clip": 1000.0,
                "to_world": mi.ScalarTransform4f.look_at(origin=[0.1, 3.5, 0.1], target=[0.1, 0, 0.1], up=[0, 0, 1]),
                "myfilm": {"type": "hdrfilm", "rfilter": {"type": "gaussian"}, "sample_border": border, "width": res, "height": res},
                "mysampler": {"type": "independent", "sample_count": spp}}
    base = {"type": "scene", "integrator": {"type": "prb_reparam", "max_depth": 2},
            "sensor0": sensor(True, 64, 4), "sensor1": sensor(False, 64, 4), "sensor2": sensor(False, 32, 8),
            "DiffuseBSDF": {"type": "diffuse", "reflectance": {"type": "rgb", "value": [0.5, 0.5, 0.5]}},
            "DiffuseBSDF2": {"type": "diffuse", "reflectance": {"type": "rgb", "value": [0.8, 0.5, 0.5]}},
            "myemitter": {"type": "rectangle", "to_world": T.translate([0.0, 8.0, 0.0]).rotate([1, 0, 0], 90.0).scale(0.01),
                          "emitter": {"type": "area", "radiance": {"type": "rgb", "value": 300000.0}}},
            "Floor": {"type": "rectangle", "to_world": T.rotate([1, 0, 0], -90.0).scale(10), "mybsdf": {"type": "ref", "id": "DiffuseBSDF"}}}
    for i in range(3):
        base[f"sphere{i}"] = {"type": "obj", "filename": str(tmp_path / "sphere.obj"), "to_world": T.translate([0.3 * i, 6.0, 0.0]).scale(0.025),
                              "mybsdf": {"type": "ref", "id": "DiffuseBSDF2"}}
    scene = mi.load_dict(base)
    d = scene._data
    assert [m.name for m in d.meshes] == ["myemitter", "Floor", "sphere0", "sphere1", "sphere2"]
    assert len(d.sensors) == 3 and d.sensors[0].sample_border and not d.sensors[1].sample_border and d.sensors[2].width == 32
    assert d.integrator == {"type": "prb_reparam", "max_depth": 2}
    assert d.meshes[2].bsdf == d.meshes[3].bsdf != d.meshes[1].bsdf
    assert d.emitters[0].radiance == (300000.0,) * 3 and d.meshes[0].emitter == 0
    np.testing.assert_allclose(d.meshes[3].positions.mean(0), [0.3, 6.0, 0.0], atol=1e-6)
    with pytest.raises(RuntimeError, match='Plugin "envmap" not found'):
        mi.load_dict({"type": "scene", "e": {"type": "envmap"}})
    integ = mi.load_dict({"type": "manifold", "max_depth": 2})                  # optim.py:97-100
    assert integ.max_depth == 2 and integ.rr_depth == 5
    assert len(scene.sensors()) == 3 and scene.sensors()[2].film().size() == [32, 32]
