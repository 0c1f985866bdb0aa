"""Golden vectors from THE REFERENCE ITSELF (its `scalar_rgb` Python module, built unmodified from /root/reference into
baseline/_ref/kd by baseline/build_ref.sh) for every function of the primal / path-logging side of the hot path
(Appendix C).  Thousands of seeded inputs per function; every entry of
tests/golden/reference_vectors.npz (and of reference_kats.json's module-read values) is regenerable with

    baseline/build_ref.sh && python tests/golden/make_reference_vectors.py

The scenes are this repository's seeded SceneData generators exported to PLY meshes + a Mitsuba scene dict, so the
reference traces EXACTLY the triangles the oracle and the CUDA path trace (prim indices comparable).

Functions pinned (reference file:line → npz group):
  Scene::ray_intersect_preliminary (kd-tree, kdtree.h:2188-2290; Möller–Trumbore mesh.h:342-367)  → si/<scene>/{t,prim,prim_uv}
  Mesh::compute_surface_interaction + EPSM fields (mesh.cpp:632-892; interaction.h:221-224)        → si/<scene>/{p,n,sh_n,wi,b0,b1,ismesh,p0..n2}
  BSDF::sample / eval / pdf: diffuse.cpp:101-175, twosided.cpp:111-250, dielectric.cpp:250-380,
      conductor.cpp:248-310, roughconductor.cpp:226-440 (hf :255)                                  → bsdf/<name>/…
  Scene::sample_emitter_direction / pdf_emitter_direction / Emitter::eval (scene.cpp:225-290,
      area.cpp:82-200, shape.cpp:352-383, mesh.cpp:516-571, rectangle.cpp:152-171)                 → emitter/<scene>/…
  PerspectiveCamera::sample_ray_differential (perspective.cpp:238-279)                              → sensor/<scene>/…
  GaussianFilter::eval (gaussian.cpp:48-103, the non-CUDA polynomial)                              → rfilter/gaussian
  ImageBlock::put + HDRFilm::develop (imageblock.cpp:174-412 scalar branch with the DISCRETISED filter table,
      hdrfilm.cpp:286-391) — loose check only: the JIT branch (:414-528) that the path uses evaluates the filter directly → film/…
  `path` integrator (src/integrators/path.cpp:1-324) converged image, for the statistical check of the whole primal loop   → path/<scene>
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REFDIR = os.path.join(ROOT, "baseline", "_ref", "kd")
FIXTURE = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
N = 2048                                     # vectors per function and scene / BSDF


def _import_reference():
    if not os.path.isfile(os.path.join(REFDIR, "libmitsuba.so")):
        raise SystemExit("baseline/_ref/kd not built: run baseline/build_ref.sh")
    import ctypes
    for lib in ("libnanothread.so", "libasmjit-mitsuba.so", "libmitsuba.so"):          # so that no LD_LIBRARY_PATH is needed
        p = os.path.join(REFDIR, lib)
        if os.path.exists(p):
            ctypes.CDLL(p, mode=ctypes.RTLD_GLOBAL)
    sys.path.insert(0, os.path.join(REFDIR, "python"))
    import warnings
    warnings.simplefilter("ignore", SyntaxWarning)
    import drjit as dr
    import mitsuba as mi
    mi.set_variant("scalar_rgb")
    return dr, mi


# ---------------------------------------------------------------------------------------------------------------------
#  SceneData → reference scene
# ---------------------------------------------------------------------------------------------------------------------
def bsdf_dict(b, mi=None):
    from epsm_mitsuba3_b200 import capi
    rgb = lambda v: {"type": "rgb", "value": [float(x) for x in v]}
    unb = lambda v: {"type": "srgb", "color": [float(x) for x in v], "unbounded": True}       # conductor eta / k exceed 1
    if b.kind == capi.BSDF_DIFFUSE and getattr(b, "texture", None) is not None:
        d = {"type": "diffuse", "reflectance": texture_dict(mi, b.texture)}
    elif b.kind == capi.BSDF_DIFFUSE:
        d = {"type": "diffuse", "reflectance": rgb(b.reflectance)}
    elif b.kind == capi.BSDF_DIELECTRIC:
        d = {"type": "dielectric", "int_ior": float(b.eta[0]), "ext_ior": 1.0}
    elif b.kind == capi.BSDF_CONDUCTOR:
        d = {"type": "conductor", "eta": unb(b.eta), "k": unb(b.k)}
    else:
        d = {"type": "roughconductor", "eta": unb(b.eta), "k": unb(b.k), "alpha": float(b.alpha_u),
             "distribution": "ggx" if b.ggx else "beckmann", "sample_visible": bool(b.sample_visible)}
    return {"type": "twosided", "nested": d} if b.twosided else d


def texture_dict(mi, tex):
    """scenes.Texture → the reference's `bitmap` plugin on the same linear float texels (raw: no colour conversion)."""
    d = {"type": "bitmap", "bitmap": mi.Bitmap(np.ascontiguousarray(tex.texels, np.float32)), "raw": True,
         "filter_type": tex.filter, "wrap_mode": tex.wrap}
    if tex.to_uv is not None:
        m = np.asarray(tex.to_uv, np.float32)
        d["to_uv"] = mi.ScalarTransform4f([[m[0, 0], m[0, 1], 0, m[0, 2]], [m[1, 0], m[1, 1], 0, m[1, 2]], [0, 0, 1, 0], [0, 0, 0, 1]])
    return d


def sensor_dict(mi, s):
    return {"type": "perspective", "fov": float(s.fov), "near_clip": float(s.near_clip), "far_clip": float(s.far_clip),
            "to_world": mi.ScalarTransform4f(s.to_world.matrix.astype(np.float32).tolist()),
            "film": {"type": "hdrfilm", "width": int(s.width), "height": int(s.height), "rfilter": {"type": "gaussian"},
                     "sample_border": bool(s.sample_border)},
            "sampler": {"type": "independent", "sample_count": int(s.spp)}}


def to_reference_scene(mi, sd, tmpdir, integrator=None):
    """Every mesh of `sd` as a PLY file loaded by the reference's `ply` plugin (vertex normals from the file, or
    face_normals=True for flat meshes), BSDFs / area emitters / sensors as dict entries, in upload order."""
    from epsm_mitsuba3_b200 import meshio
    d = {"type": "scene"}
    if integrator:
        d["integrator"] = integrator
    for i, s in enumerate(sd.sensors):
        d[f"sensor{i}"] = sensor_dict(mi, s)
    for i, m in enumerate(sd.meshes):
        path = os.path.join(tmpdir, f"{sd.name}_{i}.ply")
        meshio.save_ply(path, m)
        e = {"type": "ply", "filename": path, "face_normals": m.normals is None, "flip_normals": bool(m.flip_normals),
             "bsdf": bsdf_dict(sd.bsdfs[m.bsdf], mi)}
        if m.emitter >= 0:
            e["emitter"] = {"type": "area", "radiance": {"type": "rgb", "value": [float(x) for x in sd.emitters[m.emitter].radiance]}}
        d[f"mesh{i:04d}"] = e
    scene = mi.load_dict(d)
    shapes = scene.shapes()
    assert len(shapes) == len(sd.meshes)
    for sh, m in zip(shapes, sd.meshes):                       # upload order is preserved
        assert sh.primitive_count() == m.faces.shape[0]
    return scene


# ---------------------------------------------------------------------------------------------------------------------
def rays_for(sd, n, seed):
    """Half camera rays through random film positions of sensor 1, half random rays through the scene's bounding box."""
    rng = np.random.default_rng(seed)
    P = np.concatenate([m.positions for m in sd.meshes]); lo, hi = P.min(0), P.max(0)
    org = rng.uniform(lo - 0.2 * (hi - lo), hi + 0.2 * (hi - lo), (n, 3)).astype(np.float32)
    tgt = rng.uniform(lo, hi, (n, 3)); d = tgt - org; d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    cam = np.asarray(sd.sensors[1].to_world.matrix[:3, 3], np.float32)
    k = n // 2
    org[:k] = cam
    tgt = P[rng.integers(0, P.shape[0], k)] + rng.normal(0, 0.02, (k, 3)) * (hi - lo); dd = tgt - cam; d[:k] = (dd / np.linalg.norm(dd, axis=1, keepdims=True)).astype(np.float32)
    return org, d


def dump_si(dr, mi, sd, scene, out, key):
    org, d = rays_for(sd, N, 11)
    foff = sd.face_offsets()
    shape_id = {sh.id(): i for i, sh in enumerate(scene.shapes())}
    rec = {k: [] for k in ("t", "prim", "prim_uv", "p", "n", "sh_n", "wi", "b0", "b1", "ismesh", "p0", "p1", "p2", "n0", "n1", "n2",
                           "uv", "flags")}
    for o_, d_ in zip(org, d):
        ray = mi.Ray3f(o_.tolist(), d_.tolist())
        pi = scene.ray_intersect_preliminary(ray)
        si = pi.compute_surface_interaction(ray)
        if not pi.is_valid():
            for k in rec:
                rec[k].append(np.inf if k == "t" else (0xffffffff if k == "prim" else 0))
            continue
        mesh = shape_id[pi.shape.id()]
        rec["t"].append(float(pi.t)); rec["prim"].append(int(foff[mesh]) + int(pi.prim_index)); rec["prim_uv"].append(list(pi.prim_uv))
        for k in ("p", "n", "wi", "p0", "p1", "p2", "n0", "n1", "n2", "uv"):
            rec[k].append(list(getattr(si, k)))
        rec["sh_n"].append(list(si.sh_frame.n)); rec["b0"].append(float(si.b0)); rec["b1"].append(float(si.b1))
        rec["ismesh"].append(float(si.ismesh)); rec["flags"].append(int(si.bsdf().flags()))
    out[f"si/{key}/o"] = org; out[f"si/{key}/d"] = d
    for k, v in rec.items():
        width = {"prim_uv": 2, "uv": 2, "t": 0, "prim": 0, "b0": 0, "b1": 0, "ismesh": 0, "flags": 0}.get(k, 3)
        a = np.zeros((N, width) if width else N, np.uint32 if k in ("prim", "flags") else np.float32)
        for i, x in enumerate(v):
            a[i] = x
        out[f"si/{key}/{k}"] = a


def dump_emitter(dr, mi, sd, scene, out, key):
    """sample_emitter_direction from the first hit of N rays; pdf_emitter_direction + Emitter::eval on rays aimed at the lights."""
    rng = np.random.default_rng(23)
    org, d = rays_for(sd, N, 12)
    smp = rng.uniform(0, 1, (N, 2)).astype(np.float32)
    res = np.zeros((N, 14), np.float32); ok = np.zeros(N, bool)
    for i in range(N):
        si = scene.ray_intersect(mi.Ray3f(org[i].tolist(), d[i].tolist()))
        if not si.is_valid():
            continue
        ds, w = scene.sample_emitter_direction(si, smp[i].tolist(), True)
        ok[i] = True
        res[i] = list(ds.p) + list(ds.n) + list(ds.d) + [float(ds.pdf), float(ds.dist)] + list(w)
    out[f"emitter/{key}/o"] = org; out[f"emitter/{key}/d"] = d; out[f"emitter/{key}/sample"] = smp
    out[f"emitter/{key}/valid"] = ok; out[f"emitter/{key}/ds"] = res
    # rays from random reference points towards random points on the emitter meshes
    lights = [m for m in sd.meshes if m.emitter >= 0]
    P = np.concatenate([m.positions for m in sd.meshes]); lo, hi = P.min(0), P.max(0)
    ref = rng.uniform(lo + 0.1 * (hi - lo), hi - 0.1 * (hi - lo), (N, 3)).astype(np.float32)
    dirs = np.zeros((N, 3), np.float32); pe = np.zeros((N, 4), np.float32); okp = np.zeros(N, bool)
    for i in range(N):
        m = lights[i % len(lights)]
        f = m.faces[rng.integers(m.faces.shape[0])]; b = rng.dirichlet([1, 1, 1])
        tgt = (m.positions[f].astype(np.float64) * b[:, None]).sum(0)
        dd = tgt - ref[i]; dd /= np.linalg.norm(dd); dirs[i] = dd
        ray = mi.Ray3f(ref[i].tolist(), dirs[i].tolist())
        si = scene.ray_intersect(ray)
        if not si.is_valid():
            continue
        rsi = mi.SurfaceInteraction3f(); rsi.p = ref[i].tolist()
        ds = mi.DirectionSample3f(scene, si, rsi)
        okp[i] = True
        em = si.emitter(scene)
        if em is not None:          # scalar mode dereferences ds.emitter: a non-emitter hit has pdf 0 / eval 0 by definition
            pe[i, 0] = float(scene.pdf_emitter_direction(rsi, ds))
            pe[i, 1:4] = list(em.eval(si))
    out[f"emitter/{key}/ref_p"] = ref; out[f"emitter/{key}/ref_d"] = dirs; out[f"emitter/{key}/pdf_eval"] = pe
    out[f"emitter/{key}/pdf_valid"] = okp


def dump_sensor(dr, mi, sd, scene, out, key):
    rng = np.random.default_rng(31)
    for si_, sensor in enumerate(scene.sensors()):
        pos = rng.uniform(0, 1, (N, 2)).astype(np.float32)
        pos[:4] = [[0, 0], [1, 1], [0.5, 0.5], [0.25, 0.75]]
        res = np.zeros((N, 13), np.float32)
        for i in range(N):
            r, _ = sensor.sample_ray_differential(0.0, 0.0, pos[i].tolist(), [0.0, 0.0])
            res[i] = list(r.o) + list(r.d) + list(r.d_x) + list(r.d_y) + [float(r.maxt)]
        out[f"sensor/{key}/{si_}/pos"] = pos; out[f"sensor/{key}/{si_}/ray"] = res


BSDFS = None


def bsdf_cases():
    from epsm_mitsuba3_b200.scenes import Bsdf
    return {
        "diffuse": Bsdf.diffuse((0.7, 0.4, 0.3)),
        "diffuse_twosided": Bsdf.diffuse((0.7, 0.4, 0.3), twosided=True),
        "dielectric": Bsdf.dielectric(1.5),
        "dielectric_1p33": Bsdf.dielectric(1.33),
        "conductor": Bsdf.conductor(),
        "conductor_twosided": Bsdf.conductor(twosided=True),
        "roughconductor": Bsdf.roughconductor(0.1),
        "roughconductor_twosided": Bsdf.roughconductor(0.25, twosided=True),
        "roughconductor_ggx": Bsdf.roughconductor(0.15, ggx=True),
    }


def dump_bsdfs(dr, mi, out):
    rng = np.random.default_rng(5)
    ctx = mi.BSDFContext()
    for name, b in bsdf_cases().items():
        bsdf = mi.load_dict(bsdf_dict(b))
        wi = rng.normal(0, 1, (N, 3)); wi /= np.linalg.norm(wi, axis=1, keepdims=True)
        wi[: N // 8, 2] = np.abs(wi[: N // 8, 2])                     # some guaranteed front-side
        wi[:2] = [[0, 0, 1], [0.6, 0, 0.8]]
        wi = wi.astype(np.float32)
        s1 = rng.uniform(0, 1, N).astype(np.float32); s2 = rng.uniform(0, 1, (N, 2)).astype(np.float32)
        wo_e = rng.normal(0, 1, (N, 3)); wo_e /= np.linalg.norm(wo_e, axis=1, keepdims=True); wo_e = wo_e.astype(np.float32)
        smp = np.zeros((N, 12), np.float32); ev = np.zeros((N, 4), np.float32)
        for i in range(N):
            si = mi.SurfaceInteraction3f()
            si.wi = wi[i].tolist(); si.n = [0, 0, 1]; si.sh_frame = mi.Frame3f([0, 0, 1]); si.t = 1.0; si.wavelengths = []
            bs, w = bsdf.sample(ctx, si, float(s1[i]), s2[i].tolist(), True)
            smp[i] = list(bs.wo) + [float(bs.pdf), float(bs.eta), float(int(bs.sampled_type))] + list(w) + list(bs.hf)
            v, p = bsdf.eval_pdf(ctx, si, wo_e[i].tolist(), True)
            ev[i] = list(v) + [float(p)]
        out[f"bsdf/{name}/wi"] = wi; out[f"bsdf/{name}/s1"] = s1; out[f"bsdf/{name}/s2"] = s2; out[f"bsdf/{name}/sample"] = smp
        out[f"bsdf/{name}/wo_eval"] = wo_e; out[f"bsdf/{name}/eval_pdf"] = ev
        out[f"bsdf/{name}/flags"] = np.array([int(bsdf.flags())], np.uint32)


def dump_film(dr, mi, out):
    rf = mi.load_dict({"type": "gaussian"})
    x = np.linspace(-2.5, 2.5, 2001).astype(np.float32)
    out["rfilter/gaussian/x"] = x
    out["rfilter/gaussian/y"] = np.array([float(rf.eval(float(v))) for v in x], np.float32)
    out["rfilter/gaussian/radius"] = np.array([float(rf.radius())], np.float32)
    # ImageBlock::put (scalar branch, discretised filter) + develop through an hdrfilm
    W = H = 24
    film = mi.load_dict({"type": "hdrfilm", "width": W, "height": H, "rfilter": {"type": "gaussian"}, "pixel_format": "rgb",
                         "component_format": "float32"})
    film.prepare([])
    block = film.create_block()
    rng = np.random.default_rng(9)
    pos = rng.uniform(0, W, (N, 2)).astype(np.float32); rgb = rng.uniform(0, 2, (N, 3)).astype(np.float32)
    for i in range(N):
        block.put(pos[i].tolist(), [float(rgb[i, 0]), float(rgb[i, 1]), float(rgb[i, 2]), 1.0], True)
    out["film/pos"] = pos; out["film/rgb"] = rgb
    out["film/block"] = np.array(block.tensor(), np.float32)                       # (H+2b, W+2b, 4) with the border
    out["film/border"] = np.array([block.border_size()], np.int32)
    film.put_block(block)
    out["film/develop"] = np.array(film.develop(), np.float32)                      # (H, W, 3)


def dump_normals(dr, mi, tmpdir, out):
    """Mesh::recompute_vertex_normals (mesh.cpp:257-330) as the `ply` loader calls it for files without normals
    (ply.cpp:435-437)."""
    from epsm_mitsuba3_b200 import meshio, scenes as S
    for key, m in normal_cases().items():
        m = S.Mesh(m.name, m.positions, m.faces)                        # no normals in the file
        path = os.path.join(tmpdir, f"normals_{key}.ply")
        meshio.save_ply(path, m)
        shape = mi.load_dict({"type": "ply", "filename": path})
        assert shape.has_vertex_normals()
        p = mi.traverse(shape)
        out[f"normals/{key}/vertex_normals"] = np.array(p["vertex_normals"], np.float32).reshape(-1, 3)


def normal_cases():
    from epsm_mitsuba3_b200 import scenes as S
    body = S.articulated_stage(32, 4, 24).meshes[0]
    mirror = S.mirror_plane_scene(32, 4, 24).meshes[0]
    v, f = S.icosphere(2)
    rng = np.random.default_rng(3)
    blob = S.Mesh("blob", (v * (1 + 0.2 * rng.normal(0, 1, (v.shape[0], 1)))).astype(np.float32), f)
    return {"body": body, "mirror": mirror, "blob": blob}


def dump_path_render(dr, mi, sd, tmpdir, out, key, spp, max_depth, sensor=1):
    """Converged `path` image on sensor 1 (sample_border False) — or sensor 0 (sample_border True, the target-render
    sensor of EPSM/optim.py:56-58): the statistical oracle of the whole primal loop."""
    scene = to_reference_scene(mi, sd, tmpdir, integrator={"type": "path", "max_depth": max_depth, "rr_depth": 5})
    img = mi.render(scene, sensor=sensor, spp=spp, seed=0)
    out[f"path/{key}/image"] = np.array(img, np.float32)
    out[f"path/{key}/meta"] = np.array([spp, max_depth], np.int32)


def scene_cases():
    """The parity scenes with every shape a real mesh, as the reference sees a PLY export (si.ismesh = 1, emitter
    sampling by Mesh::sample_position); the analytic `rectangle` flags of the generators are cleared."""
    from epsm_mitsuba3_b200 import scenes as S
    cases = _scene_cases(S)
    for sd in cases.values():
        for m in sd.meshes:
            m.ismesh = True; m.is_rectangle = False
    return cases


def _scene_cases(S):
    return {
        "mirror": S.mirror_plane_scene(32, 4, 24),
        "rough": S.mirror_plane_scene(32, 4, 24, rough=True),
        "glass": S.glass_sphere_caustic(32, 4, 24, subdiv=3),
        "cbox": S.cornell_box(32, 4, 24),
        "stage": S.articulated_stage(32, 4, 24),
        "room": S.indoor_room(20000, 32, 4, 24),
    }


def main():
    sys.path.insert(0, ROOT)
    dr, mi = _import_reference()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for key, sd in scene_cases().items():
            scene = to_reference_scene(mi, sd, tmp)
            dump_si(dr, mi, sd, scene, out, key)
            dump_emitter(dr, mi, sd, scene, out, key)
            dump_sensor(dr, mi, sd, scene, out, key)
            print("scene", key, "hits", int(np.isfinite(out[f"si/{key}/t"]).sum()), "emitter samples", int(out[f"emitter/{key}/valid"].sum()),
                  "pdf probes", int(out[f"emitter/{key}/pdf_valid"].sum()))
        dump_bsdfs(dr, mi, out)
        dump_film(dr, mi, out)
        dump_normals(dr, mi, tmp, out)
        from epsm_mitsuba3_b200 import scenes as S
        cases = scene_cases()
        dump_path_render(dr, mi, cases["cbox"], tmp, out, "cbox", 4096, 6)
        dump_path_render(dr, mi, cases["mirror"], tmp, out, "mirror", 4096, 6)
        dump_path_render(dr, mi, cases["glass"], tmp, out, "glass", 4096, 6)
        dump_path_render(dr, mi, cases["rough"], tmp, out, "rough", 4096, 3)
        dump_path_render(dr, mi, cases["cbox"], tmp, out, "cbox_border", 4096, 6, sensor=0)
    np.savez_compressed(FIXTURE, **out)
    print("wrote", FIXTURE, os.path.getsize(FIXTURE) // 1024, "KiB,", len(out), "arrays")


if __name__ == "__main__":
    main()
