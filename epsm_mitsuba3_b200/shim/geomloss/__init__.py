"""`geomloss` stand-in for EPSM/utils/matcher.py:7,12,59-60: `SamplesLoss("sinkhorn", blur, scaling)` between two point
clouds, differentiable w.r.t. the first.  The ε-scaling Sinkhorn loop and the gradient are the hand-written kernels behind
`epsm_sinkhorn_grad` (csrc/epsm_sinkhorn.cu; algorithm restated in oracle/sinkhorn_ref.py); only p = 2, debiased, uniform
weights — the configuration the reference uses — exists."""
from __future__ import annotations

import ctypes as _C

import torch as _torch

from epsm_mitsuba3_b200 import capi as _capi

__version__ = "0.2.6-epsm-b200-shim"


class _SinkhornOp(_torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, owner):
        xs, ys = x.detach().to(_torch.float32).contiguous(), y.detach().to(_torch.float32).contiguous()
        g = _torch.empty_like(xs)
        _torch.cuda.current_stream(xs.device).synchronize()
        loss, steps = _C.c_double(), _C.c_uint32()
        rc = owner.lib.epsm_sinkhorn_grad(owner.ctx, xs.data_ptr(), xs.shape[0], ys.data_ptr(), ys.shape[0], owner.blur, owner.scaling,
                                          1.0, g.data_ptr(), _C.byref(loss), _C.byref(steps), _capi.MEM_DEVICE)
        if rc != 0:
            raise RuntimeError(owner.lib.epsm_last_error(owner.ctx).decode())
        owner.n_steps = steps.value
        ctx.save_for_backward(g)
        return x.new_tensor(loss.value)

    @staticmethod
    def backward(ctx, grad_out):
        (g,) = ctx.saved_tensors
        return g * grad_out, None, None


class SamplesLoss:
    def __init__(self, loss="sinkhorn", p=2, blur=0.05, reach=None, diameter=None, scaling=0.5, debias=True, backend="auto", **kw):
        if loss != "sinkhorn" or p != 2 or reach is not None or not debias:
            raise NotImplementedError('only SamplesLoss("sinkhorn", p=2, debias=True, reach=None) is built (EPSM/utils/matcher.py:12)')
        self.blur, self.scaling = float(blur), float(scaling)
        self.lib, self.ctx, self.n_steps = None, None, None

    def _ensure(self, device):
        if self.ctx is None:
            if device.type != "cuda":
                raise RuntimeError("geomloss stand-in: the Sinkhorn kernels need a CUDA device (sm_100a); there is no CPU fallback")
            self.lib = _capi.load_library()
            ctx = _C.c_void_p()
            if self.lib.epsm_create(device.index or 0, _C.byref(ctx)) != 0:
                raise RuntimeError(self.lib.epsm_last_error(None).decode())
            self.ctx = ctx

    def __call__(self, x, y):
        if x.ndim != 2 or y.ndim != 2 or x.shape[1] != y.shape[1]:
            raise ValueError("SamplesLoss: expected (N, D) and (M, D) point clouds")
        if x.shape[1] != 5:
            raise NotImplementedError("the Sinkhorn kernels are built for the matcher's 5-D clouds (r, g, b, x, y)")
        self._ensure(x.device)
        return _SinkhornOp.apply(x, y, self)

    def __del__(self):
        if getattr(self, "ctx", None):
            self.lib.epsm_destroy(self.ctx); self.ctx = None
