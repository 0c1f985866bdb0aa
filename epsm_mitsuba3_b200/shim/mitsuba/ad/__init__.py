"""`mitsuba.ad`: the optimizers of src/python/python/ad/optimizers.py (SGD :87-166, Adam :169-285) over the torch-backed
arrays of the `drjit` stand-in.  Same update equations, state handling (`reset`, shape changes), `mask_updates` and
`uniform` options, learning-rate dictionary; `step()` replaces every variable by a fresh leaf with gradients enabled."""
from __future__ import annotations

from collections import defaultdict

import torch as _torch

import drjit as _dr


class Optimizer:
    def __init__(self, lr, params=None):
        self.lr = defaultdict(lambda: self.lr_default)
        self.set_learning_rate(lr)
        self.variables = {}
        self.state = {}
        if params is not None:
            for k, v in params.items():
                self.__setitem__(k, v)

    def __contains__(self, key):
        return key in self.variables

    def __getitem__(self, key):
        return self.variables[key]

    def __setitem__(self, key, value):
        if not (_dr.is_diff_v(value) and _dr.is_float_v(value)):
            raise Exception("Optimizer.__setitem__(): value should be differentiable!")
        needs_reset = (key not in self.variables) or _dr.shape(self.variables[key]) != _dr.shape(value)
        v = type(value)(value.t.detach().clone())
        _dr.enable_grad(v)
        self.variables[key] = v
        if needs_reset:
            self.reset(key)

    def __delitem__(self, key):
        del self.variables[key]

    def __len__(self):
        return len(self.variables)

    def keys(self):
        return self.variables.keys()

    def items(self):
        return self.variables.items()

    def set_learning_rate(self, lr):
        if isinstance(lr, (float, int)):
            self.lr_default = float(lr)
        elif isinstance(lr, dict):
            for k, v in lr.items():
                self.lr[k] = float(v)
        else:
            raise Exception("Optimizer.set_learning_rate(): value should be a float or a dict!")

    def reset(self, key):
        pass

    def _replace(self, k, p, value):
        u = type(p)(value)
        _dr.enable_grad(u)
        self.variables[k] = u


class SGD(Optimizer):
    def __init__(self, lr, momentum=0, mask_updates=False, params=None):
        assert 0 <= momentum < 1 and lr > 0
        self.momentum, self.mask_updates = momentum, mask_updates
        super().__init__(lr, params)

    def step(self):
        for k, p in list(self.variables.items()):
            g = p.t.grad
            if g is None:
                continue
            with _torch.no_grad():
                if self.momentum != 0:
                    if self.state[k].shape != g.shape:
                        self.reset(k)
                    nxt = self.momentum * self.state[k] + g
                    step = self.lr[k] * self.state[k]
                    if self.mask_updates:
                        nz = g != 0
                        nxt = _torch.where(nz, nxt, self.state[k]); step = _torch.where(nz, step, _torch.zeros_like(step))
                    self.state[k] = nxt
                    value = p.t.detach() - step
                else:
                    value = p.t.detach() - self.lr[k] * g
            self._replace(k, p, value)

    def reset(self, key):
        if self.momentum != 0:
            self.state[key] = _torch.zeros_like(self.variables[key].t)


class Adam(Optimizer):
    def __init__(self, lr, beta_1=0.9, beta_2=0.999, epsilon=1e-8, mask_updates=False, uniform=False, params=None):
        assert 0 <= beta_1 < 1 and 0 <= beta_2 < 1 and lr > 0 and epsilon > 0
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon
        self.mask_updates, self.uniform = mask_updates, uniform
        self.t = defaultdict(lambda: 0)
        super().__init__(lr, params)

    def step(self):
        for k, p in list(self.variables.items()):
            self.t[k] += 1
            lr_t = self.lr[k] * (1 - self.beta_2 ** self.t[k]) ** 0.5 / (1 - self.beta_1 ** self.t[k])
            g = p.t.grad
            if g is None:                       # no gradient arrived: dr.grad has shape 0 → skipped (optimizers.py:244-246)
                continue
            with _torch.no_grad():
                if g.shape != self.state[k][0].shape:
                    self.reset(k); self.t[k] = 1
                m_tp, v_tp = self.state[k]
                m_t = self.beta_1 * m_tp + (1 - self.beta_1) * g
                v_t = self.beta_2 * v_tp + (1 - self.beta_2) * g * g
                if self.mask_updates:
                    nz = g != 0
                    m_t = _torch.where(nz, m_t, m_tp); v_t = _torch.where(nz, v_t, v_tp)
                self.state[k] = (m_t, v_t)
                denom = (v_t.max().sqrt() if self.uniform else v_t.sqrt()) + self.epsilon
                step = lr_t * m_t / denom
                if self.mask_updates:
                    step = _torch.where(nz, step, _torch.zeros_like(step))
                value = p.t.detach() - step
            self._replace(k, p, value)

    def reset(self, key):
        z = _torch.zeros_like(self.variables[key].t.detach())
        self.state[key] = (z, z.clone())
        self.t[key] = 0

    def __repr__(self):
        return ("Adam[\n  variables = %s,\n  lr = %s,\n  betas = (%g, %g),\n  eps = %g\n]"
                % (list(self.keys()), dict(self.lr, default=self.lr_default), self.beta_1, self.beta_2, self.epsilon))
