"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT THE PRODUCT.

A CPU restatement (torch CPU fp32 tensors = the reference's own arithmetic
substrate, `torch>=2`, pyproject.toml:21-23) of the K-FAC hot path of
gpauloski/kfac-pytorch v0.4.2.  Each function cites the reference file:line
it restates.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import this module, and only as
the checker / the timed CPU baseline -- never as a fallback for the CUDA path.

Pinned: `tests/test_oracle_golden.py` checks every function here against
fixtures under `tests/golden/` that were produced by running the UNMODIFIED
reference in the build container (`oracle/gen_golden.py`), including the
reference's own exact `get_cov` vectors (tests/layers/utils_test.py:25-72).
"""
from __future__ import annotations

import math
from typing import Callable

import torch


# ---------------------------------------------------------------- factors
def append_bias_ones(t: torch.Tensor) -> torch.Tensor:
    """kfac/layers/utils.py:8-15."""
    return torch.cat([t, t.new_ones(list(t.shape[:-1]) + [1])], dim=-1)


def get_cov(a: torch.Tensor, scale: float | None = None) -> torch.Tensor:
    """kfac/layers/utils.py:18-59 (b=None branch): a^T (a/scale), symmetrised."""
    if a.dim() != 2:
        raise ValueError('Input tensor must have 2 dimensions.')
    if scale is None:
        scale = a.size(0)
    c = a.t() @ (a / scale)
    return (c + c.t()) / 2.0


def linear_a_factor(x: torch.Tensor, has_bias: bool) -> torch.Tensor:
    """kfac/layers/modules.py:123-132."""
    x = x.reshape(-1, x.size(-1))
    if has_bias:
        x = append_bias_ones(x)
    return get_cov(x)


def linear_g_factor(g: torch.Tensor) -> torch.Tensor:
    """kfac/layers/modules.py:134-141."""
    return get_cov(g.reshape(-1, g.size(-1)))


def conv2d_patches(x, kernel_size, stride, padding) -> torch.Tensor:
    """kfac/layers/modules.py:210-237: (B,C,H,W) -> (B,Ho,Wo,C*kh*kw).

    Feature order (C, kh, kw); dilation/groups ignored like the reference.
    """
    if padding[0] + padding[1] > 0:
        x = torch.nn.functional.pad(
            x, (padding[1], padding[1], padding[0], padding[0]))
    x = x.unfold(2, kernel_size[0], stride[0]).unfold(3, kernel_size[1], stride[1])
    x = x.permute(0, 2, 3, 1, 4, 5).contiguous()
    return x.view(x.size(0), x.size(1), x.size(2), -1)


def conv2d_a_factor(x, kernel_size, stride, padding, has_bias) -> torch.Tensor:
    """kfac/layers/modules.py:170-178 (ones appended BEFORE the /spatial)."""
    p = conv2d_patches(x, kernel_size, stride, padding)
    spatial = p.size(1) * p.size(2)
    p = p.view(-1, p.size(-1))
    if has_bias:
        p = append_bias_ones(p)
    return get_cov(p / spatial)


def conv2d_g_factor(g: torch.Tensor) -> torch.Tensor:
    """kfac/layers/modules.py:180-192."""
    spatial = g.size(2) * g.size(3)
    g = g.permute(0, 2, 3, 1).reshape(-1, g.size(1))
    return get_cov(g / spatial)


def ema_update(factor, batch_sum, count: int, alpha: float) -> torch.Tensor:
    """kfac/layers/base.py:375-405: mean over micro-batches, seed I, EMA."""
    new = batch_sum if count <= 1 else (1.0 / count) * batch_sum
    if factor is None:
        factor = torch.eye(new.shape[0], dtype=new.dtype)
    return alpha * factor + (1 - alpha) * new


# ---------------------------------------------------------------- inverses
def eigen_decompose(factor: torch.Tensor):
    """kfac/layers/eigen.py:295-321 / 323-344: eigh on fp32, clamp d >= 0."""
    d, q = torch.linalg.eigh(factor.to(torch.float32))
    return torch.clamp(d, min=0.0), q


def eigen_dgda(dg, da, damping: float) -> torch.Tensor:
    """kfac/layers/eigen.py:345-348."""
    return 1 / (torch.outer(dg, da) + damping)


def damped_inverse(factor: torch.Tensor, damping: float) -> torch.Tensor:
    """kfac/layers/inverse.py:186-213."""
    n = factor.shape[0]
    return torch.linalg.inv(
        (factor + damping * torch.eye(n, dtype=factor.dtype)).to(torch.float32))


# ---------------------------------------------------------------- precondition
def grad_matrix(weight_grad, bias_grad=None) -> torch.Tensor:
    """kfac/layers/modules.py:56-69,194-208: (out, in*kh*kw [+1])."""
    g = weight_grad.reshape(weight_grad.size(0), -1)
    if bias_grad is not None:
        g = torch.cat([g, bias_grad.view(-1, 1)], 1)
    return g


def precondition_eigen(grad, qa, qg, dgda=None, da=None, dg=None, damping=None):
    """kfac/layers/eigen.py:371-385."""
    v1 = qg.t() @ grad @ qa
    if dgda is not None:
        v2 = v1 * dgda
    else:
        v2 = v1 / (torch.outer(dg, da) + damping)
    return qg @ v2 @ qa.t()


def precondition_inverse(grad, a_inv, g_inv) -> torch.Tensor:
    """kfac/layers/inverse.py:231-234."""
    return g_inv @ grad @ a_inv


def grad_scale(precond: list, grads: list, lr: float, kl_clip: float) -> float:
    """kfac/base_preconditioner.py:411-435 (python-float accumulation)."""
    vg = 0.0
    for p, g in zip(precond, grads):
        vg += (p * g * lr ** 2).sum().item()
    if vg == 0.0:
        return 1.0
    return min(1.0, math.sqrt(kl_clip / abs(vg)))


# ---------------------------------------------------------------- full step
class OracleLayer:
    """Per-layer state, mirroring KFACEigenLayer/KFACInverseLayer fields."""

    def __init__(self, name: str, module: torch.nn.Module):
        self.name = name
        self.module = module
        self.is_conv = isinstance(module, torch.nn.Conv2d)
        self.has_bias = module.bias is not None
        self.a_batch = None
        self.g_batch = None
        self.a_count = 0
        self.g_count = 0
        self.A = None
        self.G = None
        self.qa = self.qg = self.da = self.dg = self.dgda = None
        self.a_inv = self.g_inv = None
        self.P = None

    def a_factor_of(self, x):
        m = self.module
        if self.is_conv:
            return conv2d_a_factor(x, m.kernel_size, m.stride, m.padding, self.has_bias)
        return linear_a_factor(x, self.has_bias)

    def g_factor_of(self, g):
        return conv2d_g_factor(g) if self.is_conv else linear_g_factor(g)


def _hp(name):
    def get(self):
        v = getattr(self, '_' + name)
        return v(self.steps) if callable(v) else v

    def put(self, v):
        setattr(self, '_' + name, v)
    return property(get, put)


class OraclePreconditioner:
    """Single-process restatement of BaseKFACPreconditioner.step()
    (kfac/base_preconditioner.py:310-382) + hooks (:437-479) on CPU.

    World size 1 (the reference's single-process path has no collectives,
    kfac/distributed.py:221-222).  Used as the parity checker and as the
    timed CPU baseline (`bench.py`, kind "port").
    """

    factor_update_steps = _hp('factor_update_steps')
    inv_update_steps = _hp('inv_update_steps')
    damping = _hp('damping')
    factor_decay = _hp('factor_decay')
    kl_clip = _hp('kl_clip')
    lr = _hp('lr')

    def __init__(self, model, *, factor_update_steps=1, inv_update_steps=1,
                 damping=0.001, factor_decay=0.95, kl_clip=0.001, lr=0.1,
                 accumulation_steps=1, compute_method='eigen',
                 prediv=True, skip_layers=None,
                 grad_scaler: Callable[[], float] | None = None,
                 update_factors_in_hook=True):
        import re
        # constants or callables of the step count (base_preconditioner.py:160-213)
        self._factor_update_steps = factor_update_steps
        self._inv_update_steps = inv_update_steps
        self._damping = damping
        self._factor_decay = factor_decay
        self._kl_clip = kl_clip
        self._lr = lr
        self.accumulation_steps = accumulation_steps
        self.compute_method = compute_method
        self.prediv = prediv
        self.grad_scaler = grad_scaler
        self.update_factors_in_hook = update_factors_in_hook
        self.steps = 0
        self.mini = {}
        self.layers: dict[torch.nn.Module, OracleLayer] = {}
        skip = [re.compile(p) for p in (skip_layers or [])]
        # kfac/layers/register.py:57-95
        for name, mod in model.named_modules():
            if len(list(mod.children())) != 0:
                continue
            if any(r.search(name) or r.search(mod.__class__.__name__) for r in skip):
                continue
            if not all(p.requires_grad for p in mod.parameters()):
                continue
            if isinstance(mod, (torch.nn.Linear, torch.nn.Conv2d)):
                self.layers[mod] = OracleLayer(name, mod)
                mod.register_forward_pre_hook(self._save_input)
                mod.register_full_backward_hook(self._save_grad_output)
        self.last_scale = None

    # hooks: base_preconditioner.py:437-479 + layers/base.py:345-373
    @torch.no_grad()
    def _save_input(self, module, inp):
        if not module.training or self.steps % self.factor_update_steps != 0:
            return
        L = self.layers[module]
        a = L.a_factor_of(inp[0].detach().clone())
        L.a_batch = a if L.a_batch is None else L.a_batch + a
        L.a_count += 1
        self.mini[L.name] = self.mini.get(L.name, 0) + 1
        if self.update_factors_in_hook and self.mini[L.name] % self.accumulation_steps == 0:
            L.A = ema_update(L.A, L.a_batch, L.a_count, self.factor_decay)
            L.a_batch, L.a_count = None, 0

    @torch.no_grad()
    def _save_grad_output(self, module, gin, gout):
        if not module.training or self.steps % self.factor_update_steps != 0:
            return
        L = self.layers[module]
        g = gout[0] if not isinstance(gout, torch.Tensor) else gout
        if self.grad_scaler is not None:
            g = g / self.grad_scaler()
        g = L.g_factor_of(g)
        L.g_batch = g if L.g_batch is None else L.g_batch + g
        L.g_count += 1
        if self.update_factors_in_hook and self.mini.get(L.name, 0) % self.accumulation_steps == 0:
            L.G = ema_update(L.G, L.g_batch, L.g_count, self.factor_decay)
            L.g_batch, L.g_count = None, 0

    @torch.no_grad()
    def step(self):
        layers = list(reversed(list(self.layers.values())))
        # factors updated here instead of in the hooks (base_preconditioner.py:323-333)
        if not self.update_factors_in_hook and self.steps % self.factor_update_steps == 0:
            for L in layers:
                self.mini[L.name] = 0
                if L.a_count:
                    L.A = ema_update(L.A, L.a_batch, L.a_count, self.factor_decay)
                    L.a_batch, L.a_count = None, 0
                if L.g_count:
                    L.G = ema_update(L.G, L.g_batch, L.g_count, self.factor_decay)
                    L.g_batch, L.g_count = None, 0
        if self.steps % self.inv_update_steps == 0:
            for L in layers:
                if self.compute_method == 'eigen':
                    L.da, L.qa = eigen_decompose(L.A)
                    L.dg, L.qg = eigen_decompose(L.G)
                    if self.prediv:
                        L.dgda = eigen_dgda(L.dg, L.da, self.damping)
                else:
                    L.a_inv = damped_inverse(L.A, self.damping)
                    L.g_inv = damped_inverse(L.G, self.damping)
        grads = []
        for L in layers:
            m = L.module
            g = grad_matrix(m.weight.grad, m.bias.grad if L.has_bias else None)
            grads.append(g)
            if self.compute_method == 'eigen':
                if self.prediv:
                    L.P = precondition_eigen(g, L.qa, L.qg, dgda=L.dgda)
                else:
                    L.P = precondition_eigen(g, L.qa, L.qg, da=L.da, dg=L.dg,
                                             damping=self.damping)
            else:
                L.P = precondition_inverse(g, L.a_inv, L.g_inv)
        scale = grad_scale([L.P for L in layers], grads, self.lr, self.kl_clip)
        self.last_scale = scale
        # layers/base.py:407-423 + modules.py:87-97
        for L in layers:
            m = L.module
            p = scale * L.P
            if L.has_bias:
                m.weight.grad = p[:, :-1].reshape(m.weight.grad.shape).contiguous()
                m.bias.grad = p[:, -1].reshape(m.bias.grad.shape).contiguous()
            else:
                m.weight.grad = p.reshape(m.weight.grad.shape).contiguous()
        self.steps += 1
        self.mini = {}
