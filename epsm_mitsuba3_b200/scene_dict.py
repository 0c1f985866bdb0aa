"""Mitsuba scene descriptions → SceneData: `mi.load_dict` (src/core/python/xml_v.cpp:89-213) and `mi.load_file`
(src/core/xml.cpp:1422-1456) for the plugins on the EPSM path (§8f rank 4).

Supported plugins — everything BASELINE's configs and the programmatic EPSM/exp scenes use:
  integrator  path | prb | prb_reparam | manifold | manifold_caustic            (kept as a dict; instantiated by the caller)
  sensor      perspective (+ hdrfilm with gaussian rfilter, independent sampler)   src/sensors/perspective.cpp, src/films/hdrfilm.cpp
  bsdf        diffuse | dielectric | conductor | roughconductor | twosided         src/bsdfs/*.cpp (row 7)
  shape       obj | ply | serialized | rectangle | cube | sphere                   src/shapes/*.cpp (analytic shapes are tessellated
              and flagged non-mesh, which is how the EPSM fields see them: si.ismesh = 0, epsm.py:793-796)
  emitter     area                                                                 src/emitters/area.cpp
  texture     rgb / spectrum / float values; `bitmap` as a diffuse reflectance (src/textures/bitmap.cpp: nearest / bilinear,
              repeat / clamp / mirror, to_uv, raw) — elsewhere a bitmap is reduced to its mean colour
Anything else raises `Plugin "<type>" not found`, as the reference's plugin manager does (src/core/plugin.cpp:174-211)."""
from __future__ import annotations

import math
import os
import warnings
import xml.etree.ElementTree as ET
from typing import Any, Dict, Optional

import numpy as np

from . import meshio
from . import scenes as S

INTEGRATORS = ("path", "prb", "prb_reparam", "direct_reparam", "manifold", "manifold_caustic")
BSDFS = ("diffuse", "dielectric", "conductor", "roughconductor", "twosided")
SHAPES = ("obj", "ply", "serialized", "rectangle", "cube", "sphere")


class PluginNotFound(RuntimeError):
    def __init__(self, name):
        super().__init__(f'Plugin "{name}" not found')


def _matrix(t) -> np.ndarray:
    """4×4 float64 from a Transform4f of any flavour (the shim's torch one, scenes.Transform4f), a nested list or None."""
    if t is None:
        return np.eye(4)
    m = getattr(t, "matrix", t)              # Transform4f.matrix: numpy, torch or a shim Matrix4f
    m = getattr(m, "t", m)                   # shim arrays keep their tensor in .t
    if hasattr(m, "detach"):
        m = m.detach().cpu().numpy()
    return np.asarray(m, np.float64).reshape(4, 4)


def _rgb(v, default=None):
    """rgb / spectrum / float property → (r, g, b)."""
    if v is None:
        return default
    if isinstance(v, dict):
        t = v.get("type")
        if t in ("rgb", "srgb", "spectrum", "uniform", "d65"):
            return _rgb(v.get("value", v.get("color", 1.0)))
        if t == "bitmap":
            return _bitmap_texture(v).mean()
        raise PluginNotFound(t)
    a = np.asarray(v.numpy() if hasattr(v, "numpy") else v, np.float64).reshape(-1)
    if a.size == 1:
        return (float(a[0]),) * 3
    if a.size != 3:
        raise ValueError(f"rgb value with {a.size} components")
    return tuple(float(x) for x in a)


def _bitmap_texture(v, base_dir: str = "") -> S.Texture:
    """`bitmap` texture plugin (src/textures/bitmap.cpp:118-200): 8/16-bit files are converted to linear floats with the
    inverse sRGB gamma unless raw; `data` may hand over a float array (H,W[,3]) directly, as mi.Bitmap objects do."""
    if v.get("data") is not None or v.get("bitmap") is not None:
        img = np.asarray(v.get("data") if v.get("data") is not None else v.get("bitmap"), np.float64)
    else:
        fn = v.get("filename")
        path = fn if fn and (os.path.isabs(fn) or not base_dir) else os.path.join(base_dir, fn or "")
        if not fn or not os.path.exists(path):
            raise FileNotFoundError(f'bitmap texture "{fn}" not found')
        try:
            import cv2
            raw_img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
            if raw_img is not None and raw_img.ndim == 3:
                raw_img = raw_img[..., :3][..., ::-1]
        except ImportError:
            raw_img = None
        if raw_img is None:
            from PIL import Image
            raw_img = np.asarray(Image.open(path).convert("RGB"))
        if raw_img.dtype in (np.uint8, np.uint16):
            img = raw_img.astype(np.float64) / (65535.0 if raw_img.dtype == np.uint16 else 255.0)
            if not v.get("raw", False):                 # Bitmap::convert with srgb_gamma: the inverse sRGB transfer function
                img = np.where(img <= 0.04045, img / 12.92, ((img + 0.055) / 1.055) ** 2.4)
        else:
            img = raw_img.astype(np.float64)            # float formats are linear already
    m = None
    if v.get("to_uv") is not None:                      # ScalarTransform4f → the 3×3 of transform_affine(Point2f): rows/cols 0, 1, 3
        t4 = _matrix(v.get("to_uv")); m = np.array([[t4[0, 0], t4[0, 1], t4[0, 3]], [t4[1, 0], t4[1, 1], t4[1, 3]], [0, 0, 1]])
    return S.Texture(np.ascontiguousarray(img, np.float32), filter=v.get("filter_type", "bilinear"), wrap=v.get("wrap_mode", "repeat"), to_uv=m)


_IOR = {"vacuum": 1.0, "air": 1.000277, "water": 1.3330, "bk7": 1.5046, "diamond": 2.419, "glass": 1.5046, "acrylic glass": 1.49, "pyrex": 1.470,
        "polypropylene": 1.49, "pet": 1.5750, "fused quartz": 1.458, "water ice": 1.31, "acetone": 1.36, "ethanol": 1.361, "sodium chloride": 1.544,
        "amber": 1.55, "silicon": 3.969, "helium": 1.00004, "hydrogen": 1.00013, "carbon dioxide": 1.00045, "bromine": 1.661}   # include/mitsuba/render/ior.h


def _ior(v, default):
    if v is None:
        return default
    if isinstance(v, str):
        if v.lower() not in _IOR:
            raise ValueError(f'unknown IOR material "{v}"')
        return _IOR[v.lower()]
    return float(v)


def parse_bsdf(d: dict, base_dir: str = "") -> S.Bsdf:
    t = d.get("type")
    if t == "twosided":
        inner = [v for k, v in d.items() if isinstance(v, dict) and v.get("type") in BSDFS]
        if len(inner) != 1:
            raise ValueError("twosided expects exactly one nested BSDF")        # twosided.cpp (a second BSDF for the back side is not on the path)
        b = parse_bsdf(inner[0], base_dir); b.twosided = True
        return b
    if t == "diffuse":
        r = d.get("reflectance")
        if isinstance(r, dict) and r.get("type") == "bitmap":
            tex = _bitmap_texture(r, base_dir)
            b = S.Bsdf.diffuse(tex.mean()); b.texture = tex       # `reflectance` keeps the mean colour for summaries; the kernels sample the texture
            return b
        return S.Bsdf.diffuse(_rgb(r, (0.5, 0.5, 0.5)))
    if t == "dielectric":
        b = S.Bsdf.dielectric(_ior(d.get("int_ior"), 1.5046), _ior(d.get("ext_ior"), 1.000277))
        b.reflectance = _rgb(d.get("specular_reflectance"), (1.0, 1.0, 1.0)); b.transmittance = _rgb(d.get("specular_transmittance"), (1.0, 1.0, 1.0))
        return b
    if t in ("conductor", "roughconductor"):
        material = d.get("material", "none")
        if "eta" in d or "k" in d:
            eta, k = _rgb(d.get("eta"), (0.0, 0.0, 0.0)), _rgb(d.get("k"), (1.0, 1.0, 1.0))
        elif str(material).lower() == "none":
            eta, k = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)                               # conductor.cpp: 'none' = perfect mirror
        else:
            raise FileNotFoundError(f'conductor material "{material}": the IOR tables live in the absent mitsuba-data submodule; give eta / k')
        if t == "conductor":
            b = S.Bsdf.conductor(eta, k)
        else:
            alpha = d.get("alpha", 0.1)
            b = S.Bsdf.roughconductor(float(alpha), eta, k, ggx=str(d.get("distribution", "beckmann")).lower() == "ggx")
            if "alpha_u" in d:
                b.alpha_u, b.alpha_v = float(d["alpha_u"]), float(d.get("alpha_v", d["alpha_u"]))
            b.sample_visible = bool(d.get("sample_visible", True))
        b.reflectance = _rgb(d.get("specular_reflectance"), (1.0, 1.0, 1.0))
        return b
    raise PluginNotFound(t)


def parse_sensor(d: dict) -> S.Sensor:
    if d.get("type") != "perspective":
        raise PluginNotFound(d.get("type"))
    film = next((v for v in d.values() if isinstance(v, dict) and v.get("type") in ("hdrfilm", "specfilm")), {"type": "hdrfilm"})
    if film.get("type") != "hdrfilm":
        raise PluginNotFound(film.get("type"))
    rf = next((v for v in film.values() if isinstance(v, dict) and "type" in v), {"type": "gaussian"})
    if rf.get("type") != "gaussian":
        raise PluginNotFound(rf.get("type"))                                       # only the Gaussian is on the path
    sampler = next((v for v in d.values() if isinstance(v, dict) and v.get("type") in ("independent", "stratified", "ldsampler", "multijitter", "orthogonal")),
                   {"type": "independent"})
    if sampler.get("type") != "independent":
        raise PluginNotFound(sampler.get("type"))
    w, h = int(film.get("width", 768)), int(film.get("height", 576))
    if "fov" not in d:
        raise ValueError("perspective: `fov` is required (focal_length is not supported)")
    fov, axis = float(d["fov"]), str(d.get("fov_axis", "x"))
    if axis in ("y",) or (axis == "smaller" and h < w) or (axis == "larger" and h > w):
        fov = math.degrees(2 * math.atan(math.tan(math.radians(fov) / 2) * w / h))  # perspective.cpp parse_fov
    return S.Sensor(S.Transform4f(_matrix(d.get("to_world"))), fov, w, h, float(d.get("near_clip", 1e-2)), float(d.get("far_clip", 1e4)),
                    int(sampler.get("sample_count", 4)), bool(film.get("sample_border", False)))


def _shape_mesh(key: str, d: dict, base_dir: str) -> S.Mesh:
    t = d["type"]
    tw = S.Transform4f(_matrix(d.get("to_world")))
    if t in ("obj", "ply", "serialized"):
        fn = d["filename"]
        path = fn if os.path.isabs(fn) or not base_dir else os.path.join(base_dir, fn)
        kw = dict(to_world=tw, face_normals=bool(d.get("face_normals", False)), name=key)
        kw["return_texcoords"] = True
        if t == "obj":
            m, uv = meshio.load_obj(path, **kw)
        elif t == "ply":
            m, uv = meshio.load_ply(path, **kw)
        else:
            m, uv = meshio.load_serialized(path, shape_index=int(d.get("shape_index", 0)), **kw)
        m.texcoords = uv                                 # si.uv and the uv-aligned dp_du / dp_dv of the shading frame (mesh.cpp:737-763)
    elif t == "rectangle":
        m = S.rectangle(tw, name=key)
    elif t == "cube":
        m = S.cube(tw, name=key)
    elif t == "sphere":
        c, r = np.asarray(d.get("center", [0, 0, 0]), np.float64), float(d.get("radius", 1.0))
        v, f = S.icosphere(int(d.get("_tessellation", 4)))
        p = tw.apply_points(v * r + c)
        m = S.Mesh(key, p, f, normals=tw.apply_normals(v), ismesh=False)
    else:
        raise PluginNotFound(t)
    m.flip_normals = bool(d.get("flip_normals", False))
    return m


def scene_from_dict(d: dict, base_dir: str = "") -> S.SceneData:
    """`mi.load_dict({'type': 'scene', …})`.  Returns SceneData with `.integrator` (dict or None) attached; meshes are named
    by their dict keys, so parameters are `<key>.vertex_positions` as in mi.traverse."""
    if d.get("type") != "scene":
        raise ValueError("scene_from_dict expects {'type': 'scene', …}")
    sd = S.SceneData(name=str(d.get("id", "scene")))
    sd.integrator = None
    named_bsdfs: Dict[str, int] = {}
    default_bsdf: Optional[int] = None

    def bsdf_id(spec, own_key=None) -> int:
        if spec.get("type") == "ref":
            rid = spec.get("id")
            if rid not in named_bsdfs:
                raise KeyError(f'reference to unknown object "{rid}"')
            return named_bsdfs[rid]
        i = sd.add_bsdf(parse_bsdf(spec, base_dir))
        for name in (spec.get("id"), own_key):
            if name:
                named_bsdfs[name] = i
        return i

    # first pass: named BSDFs (references may point backwards or forwards, as in XML)
    for key, v in d.items():
        if isinstance(v, dict) and v.get("type") in BSDFS:
            bsdf_id(v, key)
    for key, v in d.items():
        if not isinstance(v, dict):
            continue
        t = v.get("type")
        if t in INTEGRATORS:
            sd.integrator = dict(v)
        elif t == "perspective":
            sd.sensors.append(parse_sensor(v))
        elif t in BSDFS:
            continue
        elif t in SHAPES:
            m = _shape_mesh(key, v, base_dir)
            specs = [x for x in v.values() if isinstance(x, dict) and (x.get("type") in BSDFS or x.get("type") == "ref")]
            if specs:
                m.bsdf = bsdf_id(specs[0])
            else:
                if default_bsdf is None:
                    default_bsdf = sd.add_bsdf(S.Bsdf.diffuse((0.5, 0.5, 0.5)))      # Shape without a BSDF gets a default diffuse one (shape.cpp)
                m.bsdf = default_bsdf
            em = next((x for x in v.values() if isinstance(x, dict) and x.get("type") == "area"), None)
            other = next((x.get("type") for x in v.values() if isinstance(x, dict) and x.get("type") not in BSDFS + ("ref", "area", None)
                          and not isinstance(x.get("type"), type(None)) and x.get("type") not in ("rgb",)), None)
            if other is not None:
                raise PluginNotFound(other)
            sd.add_mesh(m, radiance=_rgb(em.get("radiance"), (1.0, 1.0, 1.0)) if em else None)
        elif t in ("ref",):
            continue
        elif t is not None:
            raise PluginNotFound(t)
    if not sd.meshes:
        raise ValueError("scene without shapes")
    return sd


# ---------------------------------------------------------------------------------------------------------------------
#  Mitsuba XML (src/core/xml.cpp): the subset the EPSM scenes use — default / $parameters, include, typed properties,
#  transform blocks, ref, nested objects.  Produces the dict that scene_from_dict consumes.
# ---------------------------------------------------------------------------------------------------------------------
_OBJECT_TAGS = ("scene", "integrator", "sensor", "sampler", "film", "rfilter", "bsdf", "shape", "emitter", "texture", "medium", "phase", "volume", "spectrum")


def _xml_transform(node, subst) -> S.Transform4f:
    m = np.eye(4)
    val = lambda n, k, dflt=None: subst(n.get(k)) if n.get(k) is not None else dflt
    for op in node:
        if op.tag == "translate":
            v = [float(x) for x in val(op, "value").replace(",", " ").split()] if op.get("value") else [float(val(op, a, 0)) for a in "xyz"]
            t = S.Transform4f().translate(v).matrix
        elif op.tag == "scale":
            if op.get("value"):
                v = [float(x) for x in val(op, "value").replace(",", " ").split()]
                v = v * 3 if len(v) == 1 else v
            else:
                v = [float(val(op, a, 1)) for a in "xyz"]
            t = S.Transform4f().scale(v).matrix
        elif op.tag == "rotate":
            axis = [float(x) for x in val(op, "value").replace(",", " ").split()] if op.get("value") else [float(val(op, a, 0)) for a in "xyz"]
            t = S.Transform4f().rotate(axis, float(val(op, "angle"))).matrix
        elif op.tag == "lookat":
            p = lambda k: [float(x) for x in val(op, k).replace(",", " ").split()]
            t = S.Transform4f.look_at(p("origin"), p("target"), p("up") if op.get("up") else [0, 1, 0]).matrix
        elif op.tag == "matrix":
            v = [float(x) for x in val(op, "value").replace(",", " ").split()]
            t = np.array(v).reshape(4, 4) if len(v) == 16 else np.block([[np.array(v).reshape(3, 3), np.zeros((3, 1))], [np.zeros((1, 3)), np.ones((1, 1))]])
        else:
            raise ValueError(f"unsupported transform operation <{op.tag}>")
        m = t @ m                                                      # later operations are applied after earlier ones (xml.cpp)
    return S.Transform4f(m)


def xml_to_dict(path: str, **parameters) -> dict:
    """Parse a Mitsuba 3 scene file into the dict form of mi.load_dict.  `parameters` override <default> values ($name)."""
    base = os.path.dirname(os.path.abspath(path))
    defaults: Dict[str, str] = {k: str(v) for k, v in parameters.items()}
    counter = [0]

    def subst(s: Optional[str]) -> Optional[str]:
        if s is None or "$" not in s:
            return s
        for k in sorted(defaults, key=len, reverse=True):
            s = s.replace("$" + k, defaults[k])
        if "$" in s:
            raise ValueError(f'undefined parameter in "{s}"')
        return s

    def obj(node, base_dir) -> Dict[str, Any]:
        d: Dict[str, Any] = {"type": subst(node.get("type")) if node.tag != "scene" else "scene"}
        if node.get("id"):
            d["id"] = node.get("id")
        for ch in node:
            tag = ch.tag
            name = ch.get("name")
            if tag == "default":
                defaults.setdefault(ch.get("name"), subst(ch.get("value")))
            elif tag in ("integer", "float", "boolean", "string"):
                v = subst(ch.get("value"))
                d[name] = int(v) if tag == "integer" else float(v) if tag == "float" else (v.lower() == "true") if tag == "boolean" else v
                if tag == "string" and name == "filename" and not os.path.isabs(v):
                    d[name] = os.path.join(base_dir, v)
            elif tag in ("rgb", "spectrum"):
                v = [float(x) for x in subst(ch.get("value")).replace(",", " ").split()]
                d[name] = {"type": "rgb", "value": v if len(v) > 1 else v[0]}
            elif tag in ("point", "vector"):
                d[name] = ([float(x) for x in subst(ch.get("value")).replace(",", " ").split()] if ch.get("value")
                           else [float(subst(ch.get(a, "0"))) for a in "xyz"])
            elif tag == "transform":
                d[name] = _xml_transform(ch, subst)
            elif tag == "ref":
                counter[0] += 1
                d[name or f"_ref_{counter[0]}"] = {"type": "ref", "id": ch.get("id")}
            elif tag == "include":
                inc = subst(ch.get("filename"))
                inc = inc if os.path.isabs(inc) else os.path.join(base_dir, inc)
                sub = obj(ET.parse(inc).getroot(), os.path.dirname(inc))
                d.update({k: v for k, v in sub.items() if k not in ("type", "id")})
            elif tag == "alias":
                continue
            elif tag in _OBJECT_TAGS:
                counter[0] += 1
                key = ch.get("id") or name or f"_{tag}_{counter[0]}"
                d[key] = obj(ch, base_dir)
            else:
                raise ValueError(f"unsupported XML element <{tag}>")
        return d

    root = ET.parse(path).getroot()
    if root.tag != "scene":
        raise ValueError("not a Mitsuba scene file")
    return obj(root, base)


def scene_from_file(path: str, **parameters) -> S.SceneData:
    """`mi.load_file(path, **parameters)`."""
    return scene_from_dict(xml_to_dict(path, **parameters), os.path.dirname(os.path.abspath(path)))
