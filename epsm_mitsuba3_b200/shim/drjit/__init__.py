"""`drjit` stand-in: the subset EPSM/optim.py and EPSM/exp/*.py use, over torch.

Arrays wrap a torch tensor (`.t`); differentiation is torch autograd (`dr.enable_grad`, `dr.backward`, `dr.grad`).
Layout: `Float` is a 1-D array of width W; `Vector3f` / `Point3f` hold (W, 3) and are indexed BY COMPONENT like Dr.Jit's
structure-of-arrays vectors (`v[1] = 0` clears y, `v.x`); `TensorXf` is an N-d tensor with NumPy-style indexing."""
from __future__ import annotations

import builtins as _b
import math as _math

import numpy as _np
import torch as _torch

__version__ = "0.4.3-epsm-b200-shim"


def _device():
    return _torch.device("cuda", _torch.cuda.current_device()) if _torch.cuda.is_available() else _torch.device("cpu")


def _raw(x, dtype=_torch.float32):
    """torch tensor of anything array-like (keeps the autograd graph of wrapped arrays)."""
    if isinstance(x, ArrayBase):
        return x.t
    if isinstance(x, _torch.Tensor):
        return x.to(device=_device(), dtype=dtype) if x.dtype.is_floating_point or dtype != _torch.float32 else x.to(_device())
    if isinstance(x, (list, tuple)) and _b.any(isinstance(e, (ArrayBase, _torch.Tensor)) for e in x):
        parts = [_raw(e, dtype).reshape(-1) for e in x]
        w = _b.max(p.shape[0] for p in parts)
        return _torch.stack([p.expand(w) for p in parts], dim=-1)
    return _torch.as_tensor(_np.asarray(x), dtype=dtype, device=_device())


class ArrayBase:
    """Common behaviour: arithmetic with broadcasting, `.torch()`, `.numpy()`, NumPy protocol."""
    __array_priority__ = 1000
    IsTensor = False

    def __init__(self, t):
        self.t = t

    # --- conversions
    def torch(self):
        return self.t.detach()

    def numpy(self):
        return self.t.detach().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def __len__(self):
        return self.t.shape[0] if self.t.ndim else 1

    @property
    def shape(self):
        return tuple(self.t.shape)

    def __repr__(self):
        return f"{type(self).__name__}({self.t.detach().cpu().numpy().tolist()})"

    __str__ = __repr__

    def __float__(self):
        return float(self.t.detach().reshape(-1)[0])

    def __bool__(self):
        return bool(self.t.detach().reshape(-1)[0])

    # --- arithmetic
    def _wrap(self, t):
        return type(self)(t)

    def _other(self, o):
        t = _raw(o, self.t.dtype if self.t.dtype.is_floating_point else _torch.float32)
        if isinstance(self, _VecBase) and t.ndim == 1 and not isinstance(o, _VecBase):
            if isinstance(o, (list, tuple, _np.ndarray)) and t.shape[0] == self.Size:
                t = t.reshape(1, -1)                                     # a literal vector
            else:
                t = t.reshape(-1, 1)                                     # a Float broadcasts over the components
        return t

    def _bin(self, o, f, swap=False):
        if isinstance(o, ArrayBase) and isinstance(o, _VecBase) and not isinstance(self, _VecBase):
            return NotImplemented
        b = self._other(o)
        return self._wrap(f(b, self.t) if swap else f(self.t, b))

    def __add__(self, o): return self._bin(o, _torch.add)
    def __radd__(self, o): return self._bin(o, _torch.add, True)
    def __sub__(self, o): return self._bin(o, _torch.sub)
    def __rsub__(self, o): return self._bin(o, _torch.sub, True)
    def __mul__(self, o): return self._bin(o, _torch.mul)
    def __rmul__(self, o): return self._bin(o, _torch.mul, True)
    def __truediv__(self, o): return self._bin(o, _torch.div)
    def __rtruediv__(self, o): return self._bin(o, _torch.div, True)
    def __pow__(self, o): return self._bin(o, _torch.pow)
    def __neg__(self): return self._wrap(-self.t)
    def __abs__(self): return self._wrap(self.t.abs())

    def _cmp(self, o, f):
        return Mask(f(self.t, self._other(o)))

    def __lt__(self, o): return self._cmp(o, _torch.lt)
    def __le__(self, o): return self._cmp(o, _torch.le)
    def __gt__(self, o): return self._cmp(o, _torch.gt)
    def __ge__(self, o): return self._cmp(o, _torch.ge)

    # --- element access (Float / TensorXf: positional; vectors override with component access)
    def __getitem__(self, k):
        k = k.t if isinstance(k, ArrayBase) else k
        r = self.t[k]
        return self._wrap(r if r.ndim else r.reshape(1))

    def __setitem__(self, k, v):
        k = k.t if isinstance(k, ArrayBase) else k
        v = _raw(v, self.t.dtype) if not isinstance(v, (int, float)) else v
        if self.t.requires_grad and self.t.is_leaf:
            with _torch.no_grad():
                self.t[k] = v
        else:
            self.t[k] = v


class Mask(ArrayBase):
    def __init__(self, t):
        super().__init__(_raw(t, _torch.bool) if not isinstance(t, _torch.Tensor) else t.to(_torch.bool))

    def __and__(self, o): return Mask(self.t & _raw(o, _torch.bool))
    def __or__(self, o): return Mask(self.t | _raw(o, _torch.bool))
    def __invert__(self): return Mask(~self.t)


class Float(ArrayBase):
    def __init__(self, *args):
        if len(args) == 0:
            t = _torch.zeros(1, device=_device())
        elif len(args) == 1:
            t = _raw(args[0])
        else:
            t = _raw(list(args))
        if not isinstance(args[0] if args else None, ArrayBase):
            t = t.reshape(-1) if t.ndim != 1 else t
        super().__init__(t.reshape(-1) if t.ndim == 0 else t)

    def _wrap(self, t):
        return Float(t) if t.ndim <= 1 else TensorXf(t)


class UInt32(ArrayBase):
    def __init__(self, *args):
        t = _raw(args[0] if len(args) == 1 else list(args), _torch.int64)
        super().__init__(t.reshape(-1) if t.ndim == 0 else t)


Int32 = UInt32


class TensorXf(ArrayBase):
    IsTensor = True

    def __init__(self, t=0.0, shape=None):
        t = _raw(t)
        if shape is not None:
            t = t.reshape(tuple(shape))
        super().__init__(t)

    def _wrap(self, t):
        return TensorXf(t)

    def __len__(self):                       # Tensor::size: number of entries (ext/drjit/src/python/tensor.h:41)
        return self.t.numel()

    @property
    def array(self):
        return Float(self.t.reshape(-1))


class _VecBase(ArrayBase):
    Size = 3

    def __init__(self, *args):
        n = self.Size
        if len(args) == 0:
            t = _torch.zeros((1, n), device=_device())
        elif len(args) == 1 and isinstance(args[0], _VecBase):
            t = args[0].t
        elif len(args) == 1 and isinstance(args[0], (int, float)):
            t = _torch.full((1, n), float(args[0]), device=_device())
        elif len(args) == 1:
            t = _raw(args[0])
            if t.ndim == 1:
                t = t.reshape(1, n) if t.shape[0] == n and not isinstance(args[0], Float) else t.reshape(-1, 1).expand(-1, n)
        else:
            assert len(args) == n, f"{type(self).__name__} takes {n} components"
            t = _raw(list(args))
        assert t.ndim == 2 and t.shape[1] == n, f"{type(self).__name__}: bad shape {tuple(t.shape)}"
        super().__init__(t)

    def __len__(self):
        return self.Size

    def __getitem__(self, i):
        if isinstance(i, int):
            if not -self.Size <= i < self.Size:
                raise IndexError(i)                  # lets list(v) / unpacking terminate
            return Float(self.t[:, i])
        if isinstance(i, tuple) and len(i) == 2 and isinstance(i[0], int):      # v[component, lane]
            return Float(self.t[i[1], i[0]].reshape(-1))
        return super().__getitem__(i)

    def __setitem__(self, i, v):
        if isinstance(i, tuple) and len(i) == 2 and isinstance(i[0], int):      # v[component, lane] = x
            val = float(v) if isinstance(v, (int, float)) else _raw(v).reshape(-1)
            with _torch.no_grad():
                self.t[i[1], i[0]] = val
            return
        if isinstance(i, int):
            v = _raw(v).reshape(-1) if not isinstance(v, (int, float)) else v
            if self.t.requires_grad and self.t.is_leaf:
                with _torch.no_grad():
                    self.t[:, i] = v
            elif self.t.requires_grad:
                t = self.t.clone(); t[:, i] = v; self.t = t
            else:
                self.t[:, i] = v
            return
        super().__setitem__(i, v)

    x = property(lambda s: s[0], lambda s, v: s.__setitem__(0, v))
    y = property(lambda s: s[1], lambda s, v: s.__setitem__(1, v))
    z = property(lambda s: s[2], lambda s, v: s.__setitem__(2, v))


class Vector3f(_VecBase):
    pass


class Point3f(_VecBase):
    pass


class Normal3f(_VecBase):
    pass


class Color3f(_VecBase):
    pass


class Vector2f(_VecBase):
    Size = 2


Point2f = Vector2f


class Vector3u(_VecBase):
    def __init__(self, *args):
        super().__init__(*args)
        self.t = self.t.to(_torch.int64)


class _MatBase(ArrayBase):
    Size = 4

    def __init__(self, m=None):
        n = self.Size
        t = _torch.eye(n, device=_device()) if m is None else (m.t if isinstance(m, ArrayBase) else _raw(m))
        super().__init__(t.reshape(n, n) if t.numel() == n * n else t)

    def __matmul__(self, o):
        if isinstance(o, _MatBase):
            return type(self)(self.t @ o.t)
        if isinstance(o, _VecBase):
            return type(o)(o.t @ self.t[: o.Size, : o.Size].T)
        return type(self)(self.t @ _raw(o))


class Matrix3f(_MatBase):
    Size = 3


class Matrix4f(_MatBase):
    Size = 4


# ---------------------------------------------------------------------------------------------------------------------
#  Functions
# ---------------------------------------------------------------------------------------------------------------------
def _lift(x):
    return x if isinstance(x, ArrayBase) else Float(x)


def _unary(f):
    def g(x):
        if isinstance(x, (int, float)):
            return f(_torch.tensor(float(x))).item()
        return x._wrap(f(x.t))
    return g


sin, cos, tan = _unary(_torch.sin), _unary(_torch.cos), _unary(_torch.tan)
asin, acos, atan = _unary(_torch.asin), _unary(_torch.acos), _unary(_torch.atan)
exp, log, sqrt, abs = _unary(_torch.exp), _unary(_torch.log), _unary(_torch.sqrt), _unary(_torch.abs)
sqr = _unary(lambda t: t * t)
rcp = _unary(lambda t: 1.0 / t)
rsqrt = _unary(_torch.rsqrt)
pi = _math.pi
inf = _math.inf


def atan2(y, x):
    return _lift(y)._wrap(_torch.atan2(_raw(y), _raw(x)))


def clamp(x, lo, hi):
    if isinstance(x, (int, float)):
        return _b.min(_b.max(x, lo), hi)
    return x._wrap(_torch.maximum(_torch.minimum(x.t, _raw(hi).to(x.t)), _raw(lo).to(x.t)))


clip = clamp


def minimum(a, b):
    return _lift(a)._bin(b, _torch.minimum)


def maximum(a, b):
    return _lift(a)._bin(b, _torch.maximum)


def select(m, a, b):
    a, b = _lift(a), _lift(b)
    mt = _raw(m, _torch.bool)
    if isinstance(a, _VecBase) and mt.ndim == 1:
        mt = mt.reshape(-1, 1)
    return (a if isinstance(a, _VecBase) or not isinstance(b, _VecBase) else b)._wrap(_torch.where(mt, a.t, a._other(b)))


def isnan(x):
    return Mask(_torch.isnan(x.t))


def isfinite(x):
    return Mask(_torch.isfinite(x.t))


def eq(a, b):
    return Mask(_raw(a) == _raw(b))


def neq(a, b):
    return Mask(_raw(a) != _raw(b))


def any(m):
    return bool(_raw(m, _torch.bool).any())


def all(m):
    return bool(_raw(m, _torch.bool).all())


def sum(x):
    """Horizontal sum: of every entry for Float / TensorXf, per component for vectors (drjit.sum)."""
    if isinstance(x, _VecBase):
        return Float(x.t.sum(dim=1))
    return Float(x.t.sum().reshape(1))


def mean(x):
    return Float(x.t.mean().reshape(1))


def max(x):
    return Float(x.t.max().reshape(1))


def min(x):
    return Float(x.t.min().reshape(1))


def dot(a, b):
    return Float((a.t * b.t).sum(dim=-1))


def norm(v):
    return Float(_torch.linalg.norm(v.t, dim=-1))


def squared_norm(v):
    return Float((v.t * v.t).sum(dim=-1))


def normalize(v):
    return v._wrap(v.t / _torch.linalg.norm(v.t, dim=-1, keepdim=True))


def cross(a, b):
    return a._wrap(_torch.cross(a.t.expand_as(_torch.broadcast_tensors(a.t, b.t)[0]), b.t.expand_as(_torch.broadcast_tensors(a.t, b.t)[0]), dim=-1))


def zeros(tp, shape=1):
    if tp in (TensorXf,):
        return TensorXf(_torch.zeros(tuple(shape) if not isinstance(shape, int) else (shape,), device=_device()))
    if isinstance(tp, type) and issubclass(tp, _VecBase):
        return tp(_torch.zeros((shape if isinstance(shape, int) else shape[0], tp.Size), device=_device()))
    return tp(_torch.zeros(shape if isinstance(shape, int) else tuple(shape), device=_device()))


def ones(tp, shape=1):
    return zeros(tp, shape) + 1.0


def full(tp, value, shape=1):
    return zeros(tp, shape) + value


def opaque(tp, value, shape=1):
    return full(tp, value, shape)


def arange(tp, *args):
    return tp(_torch.arange(*args, device=_device()))


def linspace(tp, start, stop, num, endpoint=True):
    if not endpoint:
        stop = start + (stop - start) * (num - 1) / num
    return tp(_torch.linspace(float(start), float(stop), int(num), device=_device()))


def meshgrid(*args, indexing="xy"):
    g = _torch.meshgrid(*[_raw(a).reshape(-1) for a in args], indexing=indexing)
    return tuple(Float(x.reshape(-1)) for x in g)


def ravel(v, order="A"):
    """Vector → flat Float, components interleaved (x0 y0 z0 x1 …) like the mesh buffers."""
    return Float(v.t.reshape(-1))


def unravel(tp, flat, order="A"):
    t = _raw(flat)
    return tp(t.reshape(-1, tp.Size))


def width(x):
    return x.t.shape[0] if x.t.ndim else 1


def shape(x):
    return tuple(x.t.shape)


def detached_t(x):
    return type(x) if isinstance(x, ArrayBase) else x


def is_diff_v(x):
    return isinstance(x, ArrayBase)


def is_float_v(x):
    return isinstance(x, ArrayBase) and x.t.dtype.is_floating_point


# ---- automatic differentiation (torch autograd underneath) ----------------------------------------------------------
def enable_grad(*xs):
    for x in xs:
        for a in (x if isinstance(x, (list, tuple)) else [x]):
            if not a.t.is_leaf:
                a.t = a.t.detach()
            a.t.requires_grad_(True)


def disable_grad(*xs):
    for x in xs:
        x.t = x.t.detach()


def grad_enabled(*xs):
    return builtins_any(isinstance(x, ArrayBase) and x.t.requires_grad for x in xs)


def detach(x, preserve_type=True):
    return x._wrap(x.t.detach()) if isinstance(x, ArrayBase) else x


def grad(x, preserve_type=True):
    g = x.t.grad
    return type(x)(g.clone() if g is not None else _torch.zeros_like(x.t))


def set_grad(x, g):
    x.t.grad = _raw(g).to(x.t).reshape(x.t.shape).clone() if not isinstance(g, (int, float)) else _torch.full_like(x.t, float(g))


def accum_grad(x, g):
    gg = _raw(g).to(x.t).reshape(x.t.shape)
    x.t.grad = gg.clone() if x.t.grad is None else x.t.grad + gg


def backward(x, flags=None):
    """dr.backward(x): reverse-mode from x with a unit gradient on every entry."""
    if not x.t.requires_grad:
        raise RuntimeError("drjit.backward(): the array is not attached to the AD graph (no input has gradients enabled)")
    x.t.backward(_torch.ones_like(x.t))


backward_from = backward


def schedule(*a):
    return None


def eval(*a):
    return None


def sync_thread():
    if _torch.cuda.is_available():
        _torch.cuda.synchronize()


def flush_malloc_cache():
    return None


def wrap_ad(source: str, target: str):
    """dr.wrap_ad(source='drjit', target='torch'): call a torch function on Dr.Jit arrays.  Arrays are torch tensors
    here, so the wrapper only unwraps the arguments and wraps the result; the autograd graph is shared."""
    def deco(f):
        def g(*args, **kw):
            conv = lambda a: a.t if isinstance(a, ArrayBase) else a
            r = f(*[conv(a) for a in args], **{k: conv(v) for k, v in kw.items()})
            if isinstance(r, _torch.Tensor):
                return Float(r.reshape(1)) if r.ndim == 0 else TensorXf(r)
            return r
        return g
    return deco


builtins_any = _b.any


class ADMode:
    Primal, Forward, Backward = 0, 1, 2


class ADFlag:
    ClearNone, ClearEdges, ClearInput, ClearInterior, ClearVertices, Default = 0, 1, 2, 4, 6, 7


class scalar:                      # drjit.scalar.* used for literal arrays
    ArrayXf = Float
    Array3f = Vector3f
