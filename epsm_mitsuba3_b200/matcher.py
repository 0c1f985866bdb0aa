"""Host-side mirror of the reference's matcher (EPSM/utils/matcher.py): same class name, constructor arguments and
method as `Matcher.match_Sinkhorn`, over the native Sinkhorn kernels (`epsm_sinkhorn_grad`, csrc/epsm_sinkhorn.cu)
instead of the third-party geomloss package (README.md:11; not in this image).  torch is touched at the tensor
boundary only: clamping and concatenating the 5-D clouds, as matcher.py:52-58 does."""
from __future__ import annotations

import ctypes as C

import torch

from . import capi
from .engine import EpsmError


class Matcher:
    def __init__(self, res: int, device, blur: float = 0.01, scaling: float = 0.9) -> None:     # matcher.py:10-13
        dev = torch.device(device)
        if dev.type != "cuda" or not torch.cuda.is_available():
            raise EpsmError("epsm_mitsuba3_b200.Matcher needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", dev.index or 0)
        self.resolution = res
        self.blur, self.scaling = float(blur), float(scaling)
        x = torch.linspace(0, 1, self.resolution)
        y = torch.linspace(0, 1, self.resolution)
        pos = torch.meshgrid(x, y, indexing="ij")
        self.pos = torch.cat([pos[1][..., None], pos[0][..., None]], dim=2).to(self.device).reshape(-1, 2)   # matcher.py:17
        self.lib = capi.load_library()
        ctx = C.c_void_p()
        if self.lib.epsm_create(self.device.index, C.byref(ctx)) != 0:
            raise EpsmError(self.lib.epsm_last_error(None).decode())
        self.ctx = ctx
        self.last_loss = None
        self.n_steps = None

    def __del__(self):
        if getattr(self, "ctx", None):
            self.lib.epsm_destroy(self.ctx); self.ctx = None

    def clouds(self, render_point: torch.Tensor, gt_rgb: torch.Tensor):
        """(r, g, b, x, y) per pixel, rgb clamped to [0, 1] (matcher.py:52-58)."""
        t = torch.cat([torch.clamp(gt_rgb.to(self.device, torch.float32), 0, 1), self.pos], dim=1).contiguous()
        r = torch.cat([torch.clamp(render_point.to(self.device, torch.float32), 0, 1), self.pos], dim=1).contiguous()
        return r, t

    def match_Sinkhorn(self, render_point: torch.Tensor, gt_rgb: torch.Tensor, want_loss: bool = False) -> torch.Tensor:
        """g = d( SamplesLoss("sinkhorn", blur, scaling)(render_5d, target_5d) · res² ) / d render_5d, shape (N, 5)."""
        r, t = self.clouds(render_point, gt_rgb)
        g = torch.empty_like(r)
        torch.cuda.current_stream(self.device).synchronize()          # the library runs on its own stream
        loss, steps = C.c_double(), C.c_uint32()
        rc = self.lib.epsm_sinkhorn_grad(self.ctx, r.data_ptr(), r.shape[0], t.data_ptr(), t.shape[0], self.blur, self.scaling,
                                         float(self.resolution * self.resolution), g.data_ptr(),
                                         C.byref(loss) if want_loss else None, C.byref(steps), capi.MEM_DEVICE)
        if rc != 0:
            raise EpsmError(self.lib.epsm_last_error(self.ctx).decode())
        self.n_steps = steps.value
        if want_loss:
            self.last_loss = loss.value
        else:
            self.lib.epsm_synchronize(self.ctx)
        return g
