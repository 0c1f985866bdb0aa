"""`mitsuba` stand-in: the Python surface EPSM/optim.py, EPSM/optim_human.py and EPSM/exp/*.py use ("Python
surface" row), over libepsm_b200.so.  Nothing here renders on the CPU: scenes become `epsm_mitsuba3_b200.Scene` objects
(device resident, hand-written sm_100a kernels) the first time they are rendered or traversed.

Mirrored reference interfaces (paths relative to the reference root):
  mi.set_variant / mi.variant                      src/python/python/__init__.py (variant switch; one variant exists here)
  mi.load_dict / mi.load_file                      src/core/python/xml_v.cpp:89-213, src/core/xml.cpp:1422-1456
  mi.traverse / SceneParameters                    src/python/python/util.py:24-228
  mi.render                                        src/python/python/util.py:352-519 (seed_grad = sample_tea_32(seed, 1)[0])
  mi.ad.Adam / mi.ad.SGD                           src/python/python/ad/optimizers.py
  mi.Transform4f / mi.ScalarTransform4f            include/mitsuba/core/transform.h (translate/rotate/scale/look_at chain by
                                                   right-multiplication, `@`)
  mi.util.convert_to_bitmap, mi.Bitmap             src/python/python/util.py:530-557, src/core/bitmap.cpp (sRGB, uint8)
  mi.Mesh                                          src/render/mesh.cpp (programmatic mesh + write_ply), EPSM/exp/glassslab.py:22-63
Arrays (mi.Float, mi.Point3f, mi.TensorXf …) are the `drjit` stand-in's torch-backed arrays.
"""
from __future__ import annotations

import math as _math
import sys as _sys
import types as _types

import numpy as _np
import torch as _torch

import drjit as _dr
from drjit import (Float, UInt32, Int32, Mask as Bool, TensorXf, Vector2f, Vector3f, Point2f, Point3f, Normal3f, Color3f,  # noqa: F401
                   Vector3u, Matrix3f, Matrix4f, ArrayBase as _ArrayBase)

import epsm_mitsuba3_b200 as _ep
from epsm_mitsuba3_b200 import engine as _engine, scene_dict as _sd, scenes as _S, meshio as _meshio

UInt = UInt32
Int = Int32
Spectrum = Color3f
Vector3i = Vector3u
Point3u = Vector3u
ScalarVector3f = ScalarPoint3f = ScalarNormal3f = ScalarColor3f = list
ScalarVector2f = ScalarPoint2f = list
Float32 = Float

MI_VERSION = "3.2.0"
__version__ = "3.2.0-epsm-b200-shim"
_VARIANTS = ("scalar_rgb", "cuda_ad_rgb", "llvm_ad_rgb", "cuda_rgb", "llvm_rgb")
_variant = None


def variants():
    return list(_VARIANTS)


def variant():
    return _variant


def set_variant(*names):
    """Every variant name of the RGB family selects the one implementation there is (float32 RGB on the B200)."""
    global _variant
    for n in names:
        if n in _VARIANTS:
            _variant = n
            return
    raise ImportError(f"Requested an unsupported variant \"{names[0] if names else ''}\". The following variants are available: "
                      + ", ".join(_VARIANTS) + ".")


def sample_tea_32(v0, v1, rounds=4):
    return _engine.sample_tea_32(int(v0), int(v1), rounds)


# ---------------------------------------------------------------------------------------------------------------------
#  Transform4f — differentiable (torch) 4×4 affine transform
# ---------------------------------------------------------------------------------------------------------------------
def _dev():
    return _dr._device()


def _vec3(v):
    """(3,) tensor from a literal, a Vector3f of width 1 or a list mixing floats and Floats (keeps the autograd graph)."""
    if isinstance(v, _ArrayBase):
        t = v.t
    elif isinstance(v, (list, tuple)) and any(isinstance(e, (_ArrayBase, _torch.Tensor)) for e in v):
        t = _torch.stack([(e.t if isinstance(e, _ArrayBase) else _torch.as_tensor(e, dtype=_torch.float32, device=_dev())).reshape(-1)[0]
                          .to(_dev(), _torch.float32) for e in v])
    else:
        t = _torch.as_tensor(_np.asarray(v, _np.float32), device=_dev())
    t = t.to(_torch.float32).reshape(-1)
    if t.numel() == 1:
        t = t.expand(3)
    if t.numel() != 3:
        raise TypeError("expected a 3-vector")
    return t


def _scalar(v):
    if isinstance(v, _ArrayBase):
        return v.t.reshape(-1)[0].to(_torch.float32)
    return _torch.as_tensor(float(v), dtype=_torch.float32, device=_dev())


class _chainable:
    """`Transform4f.translate(v)` (static) and `t.translate(v)` (= t @ translate(v)), as in transform.h."""

    def __init__(self, f):
        self.f = f
        self.__doc__ = f.__doc__

    def __get__(self, obj, cls):
        if obj is None:
            return lambda *a, **k: cls(self.f(*a, **k))
        return lambda *a, **k: cls(obj.matrix_t @ self.f(*a, **k))


class Transform4f:
    def __init__(self, m=None):
        if m is None:
            t = _torch.eye(4, device=_dev())
        elif isinstance(m, Transform4f):
            t = m.matrix_t
        elif isinstance(m, _S.Transform4f):
            t = _torch.as_tensor(m.matrix, dtype=_torch.float32, device=_dev())
        elif isinstance(m, _ArrayBase):
            t = m.t.reshape(4, 4)
        elif isinstance(m, _torch.Tensor):
            t = m.to(_dev(), _torch.float32).reshape(4, 4)
        else:
            t = _torch.as_tensor(_np.asarray(m, _np.float32).reshape(4, 4), device=_dev())
        self.matrix_t = t

    @property
    def matrix(self):
        return Matrix4f(self.matrix_t)

    def numpy(self):
        return self.matrix_t.detach().cpu().numpy().astype(_np.float64)

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def __repr__(self):
        return f"Transform4f({self.numpy().tolist()})"

    @_chainable
    def translate(v):
        t = _vec3(v)
        m = _torch.eye(4, device=t.device)
        return _torch.cat([_torch.cat([m[:3, :3], t.reshape(3, 1)], dim=1), m[3:4]], dim=0)

    @_chainable
    def scale(v):
        t = _vec3(v)
        return _torch.diag(_torch.cat([t, _torch.ones(1, device=t.device)]))

    @_chainable
    def rotate(axis, angle):
        a = _vec3(axis); a = a / _torch.linalg.norm(a)
        ang = _scalar(angle) * (_math.pi / 180.0)
        s, c = _torch.sin(ang), _torch.cos(ang)
        x, y, z = a[0], a[1], a[2]
        o = _torch.zeros((), device=a.device)
        rows = [[c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s, o],
                [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s, o],
                [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c), o],
                [o, o, o, o + 1]]
        return _torch.stack([_torch.stack([e.to(a.device) for e in r]) for r in rows])

    @_chainable
    def look_at(origin, target, up):
        o, t, u = _vec3(origin), _vec3(target), _vec3(up)
        d = t - o; d = d / _torch.linalg.norm(d)
        left = _torch.linalg.cross(u, d); left = left / _torch.linalg.norm(left)
        new_up = _torch.linalg.cross(d, left)
        top = _torch.stack([left, new_up, d, o], dim=1)
        return _torch.cat([top, _torch.tensor([[0.0, 0.0, 0.0, 1.0]], device=top.device)], dim=0)

    def inverse(self):
        return Transform4f(_torch.linalg.inv(self.matrix_t))

    def translation(self):
        return Vector3f(self.matrix_t[:3, 3].reshape(1, 3))

    def __matmul__(self, o):
        m = self.matrix_t
        if isinstance(o, (Transform4f, _S.Transform4f)):
            return Transform4f(m @ Transform4f(o).matrix_t)
        if isinstance(o, Normal3f):
            return Normal3f(o.t @ _torch.linalg.inv(m[:3, :3]))                 # inverse transpose applied to a row vector
        if isinstance(o, Point3f):
            return Point3f(o.t @ m[:3, :3].T + m[:3, 3])
        if isinstance(o, _ArrayBase) and o.t.ndim == 2 and o.t.shape[1] == 3:
            return type(o)(o.t @ m[:3, :3].T)                                    # Vector3f: no translation
        p = _torch.as_tensor(_np.asarray(o, _np.float32), device=m.device)
        return Point3f((p.reshape(-1, 3) @ m[:3, :3].T + m[:3, 3]))

    transform_affine = __matmul__


ScalarTransform4f = Transform4f
AffineTransform4f = Transform4f


# ---------------------------------------------------------------------------------------------------------------------
#  Scene objects
# ---------------------------------------------------------------------------------------------------------------------
class Sensor:
    def __init__(self, scene, index):
        self._scene, self.index = scene, index

    def _s(self):
        return self._scene._data.sensors[self.index]

    def film(self):
        return Film(self)

    def sampler(self):
        return _types.SimpleNamespace(sample_count=lambda: self._s().spp)

    def world_transform(self):
        return Transform4f(self._s().to_world)


class Film:
    def __init__(self, sensor):
        self._sensor = sensor

    def size(self):
        s = self._sensor._s()
        return [s.width, s.height]

    crop_size = size

    def crop_offset(self):
        return [0, 0]


class Scene:
    """mi.Scene: holds the parsed description; the device scene (`.native`) is created on first use."""

    def __init__(self, data: _S.SceneData):
        self._data = data
        self._native = None
        self._integrator = None
        self._params = None

    @property
    def native(self) -> _engine.Scene:
        if self._native is None:
            dev = _torch.cuda.current_device() if _torch.cuda.is_available() else 0
            self._native = _engine.Scene(self._data, dev)
        return self._native

    def integrator(self):
        if self._integrator is None:
            d = getattr(self._data, "integrator", None) or {"type": "path"}
            self._integrator = _engine.load_dict(d)
        return self._integrator

    def sensors(self):
        return [Sensor(self, i) for i in range(len(self._data.sensors))]

    def shapes(self):
        return list(self._data.meshes)

    def emitters(self):
        return [m for m in self._data.meshes if getattr(m, "emitter", -1) >= 0]

    def bbox(self):
        p = _np.concatenate([m.positions for m in self._data.meshes])
        return _types.SimpleNamespace(min=p.min(0).tolist(), max=p.max(0).tolist())

    def _native_params(self) -> _engine.SceneParameters:
        if self._params is None:
            self._params = _engine.traverse(self.native)
        return self._params


def load_dict(d: dict, parallel: bool = True):
    """mi.load_dict: `{'type': 'scene', …}` → Scene, an integrator dict → the integrator plugin; a `mi.Mesh` object may
    appear as a value (glassslab.py)."""
    t = d.get("type")
    if t == "scene":
        d = {k: (v._as_dict() if isinstance(v, Mesh) else v) for k, v in d.items()}
        return Scene(_sd.scene_from_dict(d))
    if t in _sd.INTEGRATORS or t in _engine._REGISTRY:
        return _engine.load_dict(d)
    raise RuntimeError(f'Plugin "{t}" not found')


def load_file(path: str, parallel: bool = True, **kwargs):
    return Scene(_sd.scene_from_file(path, **kwargs))


def load_string(s: str, **kwargs):
    import tempfile, os
    with tempfile.NamedTemporaryFile("w", suffix=".xml", delete=False) as f:
        f.write(s)
    try:
        return load_file(f.name, **kwargs)
    finally:
        os.unlink(f.name)


def cornell_box():
    """mi.cornell_box() (src/python/python/util.py:559-693) as a parsed scene description."""
    raise NotImplementedError("use epsm_mitsuba3_b200.scenes.cornell_box(): the dict form is not needed by the EPSM drivers")


def register_integrator(name, factory):
    _engine.register_integrator(name, factory)


# ---------------------------------------------------------------------------------------------------------------------
#  mi.traverse / SceneParameters
# ---------------------------------------------------------------------------------------------------------------------
class SceneParameters:
    """`params[key]` are drjit-style arrays over the native parameter tensors; assignments keep the autograd graph so
    that `dr.backward(loss(mi.render(scene, params, …)))` reaches `opt[...]` (util.py:181-228)."""

    def __init__(self, scene: Scene):
        self._scene = scene
        self._p = scene._native_params()
        self._extra = {}
        for i, s in enumerate(scene._data.sensors):
            name = getattr(s, "name", None) or ("PerspectiveCamera" if i == 0 else f"PerspectiveCamera_{i}")
            self._extra[f"{name}.to_world"] = ("sensor", i)

    def keys(self):
        return list(self._p.keys()) + list(self._extra.keys())

    def __contains__(self, k):
        return k in self._p or k in self._extra

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self.keys())

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def __getitem__(self, key):
        if key in self._extra:
            kind, i = self._extra[key]
            return Transform4f(self._scene._data.sensors[i].to_world)
        return Float(self._p[key].reshape(-1))

    def __setitem__(self, key, value):
        if key in self._extra:
            kind, i = self._extra[key]
            self._scene._data.sensors[i].to_world = _S.Transform4f(Transform4f(value).numpy())
            self._sensor_dirty = True
            return
        t = value.t if isinstance(value, _ArrayBase) else _torch.as_tensor(value)
        self._p[key] = t.to(self._p.scene.device, _torch.float32).reshape(-1, 3)

    def flags(self, key):
        return 0

    def set_dirty(self, key):
        self._p._dirty.add(key)

    def keep(self, keys):
        return None

    def update(self, values=None):
        if values:
            for k, v in values.items():
                self[k] = v
        if getattr(self, "_sensor_dirty", False):
            self._p.scene.update_sensors()
            self._sensor_dirty = False
        self._p.update()
        return []

    def __repr__(self):
        return "SceneParameters[\n" + "\n".join(f"  {k}," for k in self.keys()) + "\n]"


class _MeshParameters(dict):
    def __init__(self, mesh):
        super().__init__()
        self._mesh = mesh
        self["vertex_positions"] = Float(_torch.zeros(mesh._nv * 3))
        self["vertex_normals"] = Float(_torch.zeros(mesh._nv * 3))
        self["vertex_texcoords"] = Float(_torch.zeros(mesh._nv * 2 if mesh._uv else 0))
        self["faces"] = UInt32(_torch.zeros(mesh._nf * 3, dtype=_torch.int64))

    def update(self, *a):
        m = self._mesh
        m.positions = _np.asarray(self["vertex_positions"].numpy(), _np.float32).reshape(-1, 3)
        m.faces = _np.asarray(self["faces"].numpy(), _np.uint32).reshape(-1, 3)
        uv = self["vertex_texcoords"].numpy()
        m.texcoords = _np.asarray(uv, _np.float32).reshape(-1, 2) if uv.size else None
        m.normals = _S.vertex_normals(m.positions, m.faces)                 # Mesh::parameters_changed → recompute_vertex_normals
        return []


def traverse(obj):
    if isinstance(obj, Mesh):
        return _MeshParameters(obj)
    if isinstance(obj, Scene):
        return SceneParameters(obj)
    raise TypeError("mi.traverse(): expected a Scene or a Mesh")


# ---------------------------------------------------------------------------------------------------------------------
#  mi.render
# ---------------------------------------------------------------------------------------------------------------------
def render(scene: Scene, params=None, sensor=0, integrator=None, seed: int = 0, seed_grad: int = 0, spp: int = 0, spp_grad: int = 0):
    """util.py:352-519.  Returns a TensorXf attached to the autograd graph of the vertex parameters in `params`."""
    if not isinstance(scene, Scene):
        raise TypeError("mi.render(): scene must be a mitsuba.Scene")
    if isinstance(sensor, Sensor):
        sensor = sensor.index
    integ = integrator if integrator is not None else scene.integrator()
    if params is not None and not isinstance(params, SceneParameters):
        raise Exception("params should be an instance of mi.SceneParameters!")          # util.py:470-473 (optimizers are unwrapped there)
    p = scene._native_params()
    img = _engine.render(scene.native, p, sensor=int(sensor), integrator=integ, seed=int(seed), spp=int(spp),
                         seed_grad=int(seed_grad), spp_grad=int(spp_grad))
    return TensorXf(img)


def _install_trace():
    import json, os
    path = os.environ.get("EPSM_SHIM_TRACE")
    if path:
        def hook(rec):
            with open(path, "a") as f:
                f.write(json.dumps(rec) + "\n")
        _engine.trace_hook = hook


_install_trace()


# ---------------------------------------------------------------------------------------------------------------------
#  Bitmap / util
# ---------------------------------------------------------------------------------------------------------------------
class Bitmap:
    """Enough of mi.Bitmap for convert_to_bitmap / write: float32 RGB in, `np.array(bitmap)` out."""

    class PixelFormat:
        Y, YA, RGB, RGBA, XYZ, MultiChannel = range(6)

    def __init__(self, data, pixel_format=None):
        if isinstance(data, str):
            import cv2
            img = cv2.imread(data, cv2.IMREAD_UNCHANGED)
            if img is None:
                raise RuntimeError(f'"{data}": file does not exist!')
            data = img[..., ::-1] if img.ndim == 3 else img
        self.data = _np.asarray(data.numpy() if isinstance(data, _ArrayBase) else data)

    def __array__(self, dtype=None, copy=None):
        return self.data.astype(dtype) if dtype is not None else self.data

    def size(self):
        return [self.data.shape[1], self.data.shape[0]]

    def convert(self, pixel_format=None, component_format=None, srgb_gamma=True):
        a = self.data.astype(_np.float32)
        if srgb_gamma:
            a = _np.where(a <= 0.0031308, a * 12.92, 1.055 * _np.power(_np.maximum(a, 0.0031308), 1 / 2.4) - 0.055)
        return Bitmap(_np.clip(a * 255.0 + 0.5, 0, 255).astype(_np.uint8))

    def write(self, path, *a, **k):
        import cv2
        d = self.data
        cv2.imwrite(path, d[..., ::-1] if d.ndim == 3 and d.shape[-1] >= 3 else d)


def _convert_to_bitmap(data, uint8_srgb=True):
    """util.py:530-557: float tensor → Bitmap, converted to 8-bit sRGB unless uint8_srgb=False."""
    a = data.numpy() if isinstance(data, _ArrayBase) else _np.asarray(data.detach().cpu() if isinstance(data, _torch.Tensor) else data)
    b = Bitmap(a.astype(_np.float32))
    return b.convert(srgb_gamma=True) if uint8_srgb else b


def _write_bitmap(filename, data, write_async=True, quality=-1):
    _convert_to_bitmap(data, not filename.lower().endswith((".exr", ".hdr", ".pfm"))).write(filename)


util = _types.ModuleType("mitsuba.util")
util.convert_to_bitmap = _convert_to_bitmap
util.write_bitmap = _write_bitmap
util.traverse = traverse
util.render = render
util.cornell_box = cornell_box
util.SceneParameters = SceneParameters
_sys.modules["mitsuba.util"] = util


class OptixDenoiser:
    """Constructed by optim_human.py:34,45 and never called; OptiX does not exist on B200 (no RT cores)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        raise RuntimeError("OptixDenoiser is not available")


# ---------------------------------------------------------------------------------------------------------------------
#  mi.Mesh (programmatic meshes; EPSM/exp/glassslab.py:22-63)
# ---------------------------------------------------------------------------------------------------------------------
class Mesh:
    def __init__(self, name, vertex_count, face_count, props=None, has_vertex_normals=False, has_vertex_texcoords=False):
        self.name, self._nv, self._nf, self._uv = name, int(vertex_count), int(face_count), bool(has_vertex_texcoords)
        self.positions = _np.zeros((self._nv, 3), _np.float32)
        self.faces = _np.zeros((self._nf, 3), _np.uint32)
        self.normals = None
        self.texcoords = None
        self._bsdf = None

    def vertex_count(self):
        return self._nv

    def face_count(self):
        return self._nf

    def write_ply(self, path):
        _meshio.save_ply(path, _S.Mesh(self.name, self.positions, self.faces, normals=self.normals))

    def _as_dict(self):
        raise RuntimeError("passing mi.Mesh objects to load_dict: write the mesh with write_ply() and load it as {'type': 'ply'}")


# ---------------------------------------------------------------------------------------------------------------------
#  submodules: mitsuba.ad, mitsuba.scalar_rgb / cuda_ad_rgb / llvm_ad_rgb (variant namespaces)
# ---------------------------------------------------------------------------------------------------------------------
from . import ad  # noqa: E402,F401

for _name in _VARIANTS:
    _m = _types.ModuleType(f"mitsuba.{_name}")
    for _k in ("Float", "UInt32", "Int32", "Bool", "TensorXf", "Vector2f", "Vector3f", "Point2f", "Point3f", "Normal3f", "Color3f",
               "Vector3u", "Matrix3f", "Matrix4f", "Transform4f", "ScalarTransform4f", "load_dict", "load_file", "traverse", "render",
               "Bitmap", "Mesh", "Scene", "SceneParameters"):
        setattr(_m, _k, globals()[_k])
    _sys.modules[f"mitsuba.{_name}"] = _m
    globals()[_name] = _m
