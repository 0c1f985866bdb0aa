"""SceneData → Mitsuba 3 scene file (XML + one PLY per mesh): the inverse of scene_dict.scene_from_file.

Used to hand the synthetic BASELINE scenes to the UNMODIFIED reference binary (baseline/_ref/{kd,embree}/mitsuba, built by
baseline/build_ref.sh) for the CPU traversal baseline of BASELINE.md §2 (`bench.py`), and by users who want to look at a
scene in stock Mitsuba.  Plugin names and properties are those of the reference (src/core/xml.cpp, src/shapes/ply.cpp,
src/bsdfs/*.cpp, src/emitters/area.cpp, src/sensors/perspective.cpp, src/films/hdrfilm.cpp)."""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from . import capi, meshio
from .scenes import SceneData


def _f(v):
    return " ".join(repr(float(x)) for x in np.asarray(v, np.float64).reshape(-1))


def _bsdf_xml(b, ident: str) -> str:
    if b.kind == capi.BSDF_DIFFUSE:
        inner = f'<bsdf type="diffuse" {{id}}><rgb name="reflectance" value="{_f(b.reflectance)}"/></bsdf>'
    elif b.kind == capi.BSDF_DIELECTRIC:
        inner = (f'<bsdf type="dielectric" {{id}}><float name="int_ior" value="{float(b.eta[0])!r}"/><float name="ext_ior" value="1.0"/>'
                 f'<rgb name="specular_reflectance" value="{_f(b.reflectance)}"/><rgb name="specular_transmittance" value="{_f(b.transmittance)}"/></bsdf>')
    elif b.kind == capi.BSDF_CONDUCTOR:
        inner = (f'<bsdf type="conductor" {{id}}><rgb name="eta" value="{_f(b.eta)}"/><rgb name="k" value="{_f(b.k)}"/>'
                 f'<rgb name="specular_reflectance" value="{_f(b.reflectance)}"/></bsdf>')
    elif b.kind == capi.BSDF_ROUGHCONDUCTOR:
        inner = (f'<bsdf type="roughconductor" {{id}}><string name="distribution" value="{"ggx" if b.ggx else "beckmann"}"/>'
                 f'<float name="alpha_u" value="{float(b.alpha_u)!r}"/><float name="alpha_v" value="{float(b.alpha_v)!r}"/>'
                 f'<boolean name="sample_visible" value="{"true" if b.sample_visible else "false"}"/>'
                 f'<rgb name="eta" value="{_f(b.eta)}"/><rgb name="k" value="{_f(b.k)}"/><rgb name="specular_reflectance" value="{_f(b.reflectance)}"/></bsdf>')
    else:
        raise ValueError(f"unknown BSDF kind {b.kind}")
    if b.twosided:
        return f'<bsdf type="twosided" id="{ident}">' + inner.format(id="") + "</bsdf>"
    return inner.format(id=f'id="{ident}"')


def export_scene(sd: SceneData, out_dir: str, integrator: Optional[dict] = None, sensors=None, spp: Optional[int] = None,
                 resolution: Optional[int] = None, name: str = "scene.xml") -> str:
    """Writes `<out_dir>/<name>` + `<out_dir>/meshes/*.ply`; returns the XML path.  `sensors`: indices to export (default all);
    `spp` / `resolution` override the sensors' sample count / square film size."""
    os.makedirs(os.path.join(out_dir, "meshes"), exist_ok=True)
    integ = dict(integrator or getattr(sd, "integrator", None) or {"type": "path"})
    x = ['<scene version="3.0.0">']
    x.append(f'<integrator type="{integ["type"]}">' + "".join(
        f'<{"float" if isinstance(v, float) else "integer"} name="{k}" value="{v}"/>' for k, v in integ.items() if k != "type") + "</integrator>")
    for i in (range(len(sd.sensors)) if sensors is None else sensors):
        s = sd.sensors[i]
        w, h = (resolution, resolution) if resolution else (s.width, s.height)
        x.append(f'<sensor type="perspective"><float name="fov" value="{float(s.fov)!r}"/><string name="fov_axis" value="x"/>'
                 f'<float name="near_clip" value="{float(s.near_clip)!r}"/><float name="far_clip" value="{float(s.far_clip)!r}"/>'
                 f'<transform name="to_world"><matrix value="{_f(s.to_world.matrix)}"/></transform>'
                 f'<sampler type="independent"><integer name="sample_count" value="{int(spp or s.spp)}"/></sampler>'
                 f'<film type="hdrfilm"><integer name="width" value="{w}"/><integer name="height" value="{h}"/>'
                 f'<boolean name="sample_border" value="{"true" if s.sample_border else "false"}"/><rfilter type="gaussian"/></film></sensor>')
    for i, b in enumerate(sd.bsdfs):
        x.append(_bsdf_xml(b, f"bsdf{i}"))
    for i, m in enumerate(sd.meshes):
        rel = os.path.join("meshes", f"m{i:04d}.ply")
        meshio.save_ply(os.path.join(out_dir, rel), m)
        em = ""
        if m.emitter >= 0:
            em = f'<emitter type="area"><rgb name="radiance" value="{_f(sd.emitters[m.emitter].radiance)}"/></emitter>'
        x.append(f'<shape type="ply" id="{m.name}_{i}"><string name="filename" value="{rel}"/>'
                 + ('<boolean name="face_normals" value="true"/>' if m.normals is None else "")
                 + ('<boolean name="flip_normals" value="true"/>' if m.flip_normals else "")
                 + f'<ref id="bsdf{m.bsdf}"/>{em}</shape>')
    x.append("</scene>")
    path = os.path.join(out_dir, name)
    with open(path, "w") as f:
        f.write("\n".join(x) + "\n")
    return path
