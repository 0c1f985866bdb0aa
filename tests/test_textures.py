"""Texture coordinates and bitmap reflectances (src/textures/bitmap.cpp, mesh.cpp:731-763, used by
EPSM/exp/human.py:101-163).  The oracle and the product's device code (host harness / GPU) are held to vectors produced by the
reference's own scalar_rgb module (tests/golden/make_texture_vectors.py → texture_vectors.npz)."""
import os

import numpy as np
import pytest

from epsm_mitsuba3_b200 import scene_dict, scenes as S
from tests import parity

V = np.load(os.path.join(os.path.dirname(__file__), "golden", "texture_vectors.npz"))
VARIANTS = [(f, w) for f in ("nearest", "bilinear") for w in ("repeat", "clamp", "mirror")]
TO_UV = np.array([[1.5, 0.2, 0.1], [-0.3, 2.0, 0.25], [0, 0, 1]])


@pytest.mark.parametrize("flt,wrap", VARIANTS)
def test_texture_lookup_matches_the_reference(flt, wrap, orc):
    """Dr.Jit Texture::eval_nonaccel behind bitmap.cpp:582-625 on 1024 uv pairs in [-1.6, 2.6]² (64 of them exactly on texel
    borders), with and without a to_uv transform."""
    for tag, to_uv in (("", None), ("_touv", TO_UV)):
        tex = S.Texture(V["tex/texels"], filter=flt, wrap=wrap, to_uv=to_uv)
        got = orc.texture_eval(tex, V["tex/uv"])
        ref = V[f"tex/{flt}_{wrap}{tag}/rgb"]
        bad = np.abs(got - ref).max(1) > 6e-6                        # float32 rounding of the to_uv transform and the bilinear weights
        # a uv that lands within an ulp of a texel border may round to the other side in the reference's vectorised arithmetic
        assert bad.mean() <= 0.002, (flt, wrap, tag, float(bad.mean()), float(np.abs(got - ref).max()))


def test_eight_bit_files_are_linearised_like_the_reference(tmp_path):
    """Bitmap loading (bitmap.cpp:118-200): inverse sRGB gamma unless raw."""
    from PIL import Image
    png = str(tmp_path / "t.png"); Image.fromarray(V["png/bytes"]).save(png)
    lin = scene_dict._bitmap_texture({"type": "bitmap", "filename": png, "filter_type": "nearest"})
    raw = scene_dict._bitmap_texture({"type": "bitmap", "filename": png, "filter_type": "nearest", "raw": True})
    assert lin.filter == "nearest" and lin.wrap == "repeat"
    assert np.allclose(lin.texels, V["png/linear"], rtol=2e-6, atol=1e-7)
    assert np.allclose(raw.texels, V["png/raw"], rtol=1e-6, atol=1e-7)
    b = scene_dict.parse_bsdf({"type": "twosided", "bsdf": {"type": "diffuse", "reflectance": {"type": "bitmap", "filename": png}}})
    assert b.twosided and b.texture is not None and b.texture.filter == "bilinear"
    assert np.allclose(b.reflectance, lin.texels.reshape(-1, 3).mean(0), rtol=1e-6)


def _scene(flt, wrap):
    sd = S.textured_scene(32, 4, 24, flt, wrap)
    for m in sd.meshes:
        m.ismesh = True; m.is_rectangle = False
    return sd


def _check_hits(probe, key):
    o_, d_ = V[f"si/{key}/o"], V[f"si/{key}/d"]
    valid = V[f"si/{key}/valid"] > 0
    n = 0
    for i in np.flatnonzero(valid):
        r = probe(o_[i], d_[i])
        assert np.allclose(r["uv"], V[f"si/{key}/uv"][i], rtol=1e-5, atol=2e-6), (i, r["uv"], V[f"si/{key}/uv"][i])
        assert np.allclose(r["sh_s"], V[f"si/{key}/sh_s"][i], atol=2e-5) and np.allclose(r["sh_t"], V[f"si/{key}/sh_t"][i], atol=2e-5), i
        refl = V[f"si/{key}/refl"][i]
        n += int(np.allclose(r["reflectance"], refl, rtol=1e-5, atol=2e-6) or refl.max() == 0)      # conductors evaluate to 0
    assert n >= 0.998 * valid.sum(), (n, int(valid.sum()))                                         # texel-border uv, as above
    for i in np.flatnonzero(~valid)[:64]:
        assert not np.any(probe(o_[i], d_[i])["reflectance"])


@pytest.mark.parametrize("flt,wrap", [("nearest", "repeat"), ("bilinear", "mirror")])
def test_uv_frame_and_reflectance_match_the_reference(flt, wrap, orc):
    """si.uv, the uv-aligned shading frame (sh_frame.s / t from dp_du, mesh.cpp:746-763 + interaction.h:262-275) and the
    reflectance the diffuse BSDF sees, on 1536 rays: oracle and the product's device code compiled for the host."""
    from tests.hostsim.hostsim import SimScene
    sd = _scene(flt, wrap); key = f"{flt}_{wrap}"
    _check_hits(orc.OracleScene(sd).surface_uv, key)
    _check_hits(SimScene(sd).surface_uv, key)


def test_textured_render_converges_to_the_reference_path_integrator(orc):
    from tests.test_reference_vectors import _stat_check
    sd = _scene("nearest", "repeat")
    spp_ref, max_depth = (int(x) for x in V["path/textured_nearest_repeat/meta"])
    img = orc.OracleScene(sd, accel=True).render(1, 0, 512, max_depth, 5)[..., :3]
    _stat_check(img, V["path/textured_nearest_repeat/image"], 512, spp_ref)
    # and the texture is visible: the same scene with the mean colours instead is a different image
    flat = _scene("nearest", "repeat")
    for b in flat.bsdfs:
        b.texture = None
    img_flat = orc.OracleScene(flat, accel=True).render(1, 0, 64, max_depth, 5)[..., :3]
    assert np.abs(img_flat - V["path/textured_nearest_repeat/image"]).mean() > 3 * np.abs(img - V["path/textured_nearest_repeat/image"]).mean()


def test_scene_description_carries_textures_and_texcoords(tmp_path):
    """mi.load_dict-style description: an OBJ with vt lines and a bitmap reflectance end up as Mesh.texcoords / Bsdf.texture and
    in the C descriptors."""
    from PIL import Image
    png = str(tmp_path / "wood.png"); Image.fromarray(V["png/bytes"]).save(png)
    obj = tmp_path / "quad.obj"
    obj.write_text("v -1 -1 0\nv 1 -1 0\nv 1 1 0\nv -1 1 0\nvt 0 0\nvt 2 0\nvt 2 3\nvt 0 3\nf 1/1 2/2 3/3\nf 1/1 3/3 4/4\n")
    d = {"type": "scene",
         "sensor": {"type": "perspective", "fov": 40, "to_world": S.Transform4f.look_at([0, 0, 4], [0, 0, 0], [0, 1, 0]),
                    "film": {"type": "hdrfilm", "width": 8, "height": 8}},
         "quad": {"type": "obj", "filename": str(obj),
                  "bsdf": {"type": "diffuse", "reflectance": {"type": "bitmap", "filename": png, "filter_type": "nearest", "wrap_mode": "repeat"}}},
         "light": {"type": "rectangle", "to_world": S.Transform4f().translate([0, 0, 3]).rotate([1, 0, 0], 180),
                   "emitter": {"type": "area", "radiance": {"type": "rgb", "value": [5, 5, 5]}}}}
    sd = scene_dict.scene_from_dict(d)
    quad = [m for m in sd.meshes if m.name == "quad"][0]
    assert quad.texcoords is not None and quad.texcoords.shape == (4, 2)
    assert np.allclose(sorted(quad.texcoords[:, 0]), [0, 0, 2, 2]) and np.allclose(sorted(1 - quad.texcoords[:, 1]), [0, 0, 3, 3])   # obj.cpp flips v
    assert sd.bsdfs[quad.bsdf].texture is not None and sd.bsdfs[quad.bsdf].texture.filter == "nearest"
    c = sd.to_c()
    assert bool(c.meshes[sd.meshes.index(quad)].texcoords) and bool(c.bsdfs[quad.bsdf].reflectance_texture)
    assert c.bsdfs[quad.bsdf].reflectance_texture.contents.width == 6 and c.bsdfs[quad.bsdf].reflectance_texture.contents.height == 4


# ------------------------------------------------------------------------------------------------
#  GPU: the kernels through the C ABI
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("flt,wrap,variant,max_depth", [("nearest", "repeat", "manifold", 6), ("bilinear", "mirror", "manifold", 4)])
def test_gpu_textured_scene_parity(flt, wrap, variant, max_depth, orc):
    """Primal film, per-path records and accumulated gradients of the textured scene against the oracle (the same checks as
    tests/test_gpu_parity.py), and the converged image against the reference's `path` integrator."""
    from tests.test_gpu_parity import GpuBackend
    sd = S.textured_scene(32, 4, 24, flt, wrap)
    b, o = GpuBackend(sd), orc.OracleScene(sd, accel=True)
    parity.check_trace(b, o, parity.random_rays(sd, 4000))
    parity.check_film(b.render_raw(1, 3, 4, max_depth, 5), o.render_raw(1, 3, 4, max_depth, 5))
    res = sd.sensors[2].width; g = parity.synthetic_grad_in(res); n = res * res * 8
    paths = o.path_log(2, 11, 8, max_depth, 5, g)
    rec = b.path_records(variant, 2, 11, 8, max_depth, 5, g, 0, n)
    stats, (pg, lg, dg) = parity.check_path_records(rec, paths, variant, max_depth)
    assert stats["nonzero"] > 0
    ref_p, ref_n = parity.scatter_gradients(paths, variant, parity.n_logged(max_depth), pg, lg, dg, sd.n_vertices, max_depth)
    gp, gn = b.backward(variant, 2, 11, 8, max_depth, 5, g)
    parity.check_accumulated(gp, gn, ref_p, ref_n)


@pytest.mark.gpu
def test_gpu_textured_render_converges_to_the_reference_path_integrator():
    import epsm_mitsuba3_b200 as ep
    from tests.test_reference_vectors import _stat_check
    sd = _scene("nearest", "repeat")
    spp_ref, max_depth = (int(x) for x in V["path/textured_nearest_repeat/meta"])
    scene = ep.Scene(sd, device=0)
    img = ep.load_dict({"type": "manifold", "max_depth": max_depth}).render(scene, sensor=1, seed=0, spp=4096).cpu().numpy()[..., :3]
    _stat_check(img, V["path/textured_nearest_repeat/image"], 4096, spp_ref)
