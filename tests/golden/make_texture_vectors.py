"""Golden vectors for texture coordinates and bitmap reflectances from THE REFERENCE ITSELF (scalar_rgb module built unmodified
into baseline/_ref/kd, see make_reference_vectors.py):

    baseline/build_ref.sh && python tests/golden/make_texture_vectors.py   →  tests/golden/texture_vectors.npz

  Dr.Jit Texture lookups behind src/textures/bitmap.cpp (nearest / bilinear × repeat / clamp / mirror, to_uv)   → tex/<filter>_<wrap>/{uv,rgb}
  Bitmap loading of an 8-bit sRGB file (inverse gamma unless raw, bitmap.cpp:118-200)                             → png/{bytes,linear,raw}
  Mesh::compute_surface_interaction with vertex texcoords: si.uv, the uv-aligned shading frame (mesh.cpp:731-763,
  interaction.h:262-275) and the diffuse reflectance seen by the BSDF (diffuse.cpp:107-140)                       → si/<scene>/{o,d,uv,sh_s,sh_t,refl}
  `path` image of the textured scene, 4096 spp (statistical check of the whole primal loop)                       → path/<scene>/image
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_reference_vectors as R   # noqa: E402

FIXTURE = os.path.join(ROOT, "tests", "golden", "texture_vectors.npz")
VARIANTS = [("nearest", "repeat"), ("bilinear", "repeat"), ("nearest", "clamp"), ("bilinear", "clamp"), ("nearest", "mirror"), ("bilinear", "mirror")]


def main():
    dr, mi = R._import_reference()
    from epsm_mitsuba3_b200 import scenes as S
    out = {}
    rng = np.random.default_rng(21)
    texels = rng.uniform(0, 1, (5, 7, 3)).astype(np.float32)
    uv = rng.uniform(-1.6, 2.6, (1024, 2)).astype(np.float32)
    uv[:64] = rng.integers(-2, 4, (64, 2)) / np.array([7.0, 5.0])            # exactly on texel borders
    for flt, wrap in VARIANTS:
        for tag, to_uv in (("", None), ("_touv", np.array([[1.5, 0.2, 0.1], [-0.3, 2.0, 0.25], [0, 0, 1]]))):
            tex = mi.load_dict(R.texture_dict(mi, S.Texture(texels, filter=flt, wrap=wrap, to_uv=to_uv)))
            rgb = np.zeros((uv.shape[0], 3), np.float32)
            for i, p in enumerate(uv):
                si = mi.SurfaceInteraction3f(); si.uv = [float(p[0]), float(p[1])]
                rgb[i] = np.array(tex.eval(si), np.float32)
            out[f"tex/{flt}_{wrap}{tag}/rgb"] = rgb
    out["tex/uv"] = uv; out["tex/texels"] = texels
    with tempfile.TemporaryDirectory() as tmp:
        # 8-bit sRGB file through the reference's Bitmap loader
        from PIL import Image
        b = rng.integers(0, 256, (4, 6, 3), dtype=np.uint8); b[0, 0] = (0, 10, 11); b[0, 1] = (255, 254, 128)
        png = os.path.join(tmp, "t.png"); Image.fromarray(b).save(png)
        centres = np.stack(np.meshgrid((np.arange(6) + .5) / 6, (np.arange(4) + .5) / 4, indexing="xy"), -1).reshape(-1, 2)
        for key, raw in (("linear", False), ("raw", True)):
            tex = mi.load_dict({"type": "bitmap", "filename": png, "filter_type": "nearest", "raw": raw})
            v = np.zeros((centres.shape[0], 3), np.float32)
            for i, p in enumerate(centres):
                si = mi.SurfaceInteraction3f(); si.uv = [float(p[0]), float(p[1])]
                v[i] = np.array(tex.eval(si), np.float32)
            out[f"png/{key}"] = v.reshape(4, 6, 3)
        out["png/bytes"] = b
        # textured scenes: hits, uv, shading frame, reflectance
        for flt, wrap in (("nearest", "repeat"), ("bilinear", "mirror")):
            sd = S.textured_scene(32, 4, 24, flt, wrap)
            for m in sd.meshes:
                m.ismesh = True; m.is_rectangle = False
            key = f"{flt}_{wrap}"
            scene = R.to_reference_scene(mi, sd, tmp)
            org, d = R.rays_for(sd, 1536, 5)
            rec = {k: [] for k in ("uv", "sh_s", "sh_t", "refl", "valid")}
            ctx = mi.BSDFContext()
            for o_, d_ in zip(org, d):
                ray = mi.Ray3f(o_.tolist(), d_.tolist())
                si = scene.ray_intersect(ray)
                ok = si.is_valid()
                rec["valid"].append(1 if ok else 0)
                if not ok:
                    rec["uv"].append([0, 0]); rec["sh_s"].append([0, 0, 0]); rec["sh_t"].append([0, 0, 0]); rec["refl"].append([0, 0, 0]); continue
                rec["uv"].append(np.array(si.uv)); rec["sh_s"].append(np.array(si.sh_frame.s)); rec["sh_t"].append(np.array(si.sh_frame.t))
                si.wi = [0, 0, 1]
                rec["refl"].append(np.array(si.bsdf().eval(ctx, si, [0, 0, 1])) * np.pi)        # diffuse: f·cos = ρ/π at normal incidence
            out[f"si/{key}/o"] = org; out[f"si/{key}/d"] = d
            for k, v in rec.items():
                out[f"si/{key}/{k}"] = np.asarray(v, np.float32)
            print(key, "hits", int(np.sum(rec["valid"])))
        sd = S.textured_scene(32, 4, 24, "nearest", "repeat")
        for m in sd.meshes:
            m.ismesh = True; m.is_rectangle = False
        R.dump_path_render(dr, mi, sd, tmp, out, "textured_nearest_repeat", 4096, 6)
    np.savez_compressed(FIXTURE, **out)
    print("wrote", FIXTURE, os.path.getsize(FIXTURE) // 1024, "KiB,", len(out), "arrays")


if __name__ == "__main__":
    main()
