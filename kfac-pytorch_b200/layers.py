"""Module adapters and per-layer K-FAC state.

Mirrors the reference's `ModuleHelper` family (kfac/layers/modules.py:13-237)
and the state carried by `KFACBaseLayer` / `KFACEigenLayer` /
`KFACInverseLayer` (kfac/layers/base.py:19, eigen.py:20, inverse.py:20), but
the layer object here is only a *descriptor*: every tensor it exposes is a view
into a contiguous HBM arena owned by the preconditioner, and all arithmetic
happens in libkfac_b200.so.
"""
from __future__ import annotations

import re
from typing import Any, Callable

import torch

from kfac_b200 import _cabi


class ModuleHelper:
    """Geometry + native dispatch for one supported module type."""

    def __init__(self, module: torch.nn.Module):
        self.module = module

    def __repr__(self) -> str:
        return f'{self.__class__.__name__}({repr(self.module)})'

    @property
    def a_factor_shape(self) -> tuple[int, int]:
        raise NotImplementedError

    @property
    def g_factor_shape(self) -> tuple[int, int]:
        raise NotImplementedError

    @property
    def device(self) -> torch.device:
        return next(self.module.parameters()).device

    def has_bias(self) -> bool:
        return getattr(self.module, 'bias', None) is not None

    def has_symmetric_factors(self) -> bool:
        return True

    def get_weight_grad(self) -> torch.Tensor:
        return self.module.weight.grad

    def get_bias_grad(self) -> torch.Tensor:
        return self.module.bias.grad

    def get_grad(self) -> torch.Tensor:
        """(out, in*kh*kw [+1]) gradient matrix; torch glue for inspection /
        tests only -- the hot path reads weight.grad / bias.grad in place."""
        g = self.module.weight.grad.reshape(self.module.weight.grad.size(0), -1)
        if self.has_bias():
            g = torch.cat([g, self.module.bias.grad.view(-1, 1)], 1)
        return g

    # native accumulation of the factor statistics of one micro-batch
    def accumulate_a(self, x: torch.Tensor, acc: torch.Tensor, scratch) -> None:
        raise NotImplementedError

    def accumulate_g(self, g: torch.Tensor, acc: torch.Tensor, grad_scale: float, scratch=None) -> None:
        raise NotImplementedError


_NATIVE_INPUT_DTYPES = (torch.float32, torch.float16, torch.bfloat16)


def cast_for_factor(t: torch.Tensor, factor_dtype: torch.dtype | None) -> torch.Tensor:
    """The reference casts the activation / grad-output to `factor_dtype` before the statistics are formed
    (kfac/layers/base.py:350,364).  Here the statistics are always ACCUMULATED and STORED in float32:
      * float16 / bfloat16: the input is rounded to that dtype (same input rounding as the reference),
        the products are accumulated in fp32 (at least as accurate as the reference's half GEMM);
      * float32 / float64 / None: the input is used as is (float64 inputs are cast to float32 at the
        boundary -- the native library has no fp64 path; documented deviation)."""
    if factor_dtype in (torch.float16, torch.bfloat16) and t.dtype != factor_dtype:
        t = t.to(factor_dtype)
    if t.dtype not in _NATIVE_INPUT_DTYPES:
        t = t.to(torch.float32)
    return t


def _dtype_code(t: torch.Tensor) -> int:
    try:
        return _cabi.DTYPE_CODE[t.dtype]
    except KeyError:
        raise ValueError(f'unsupported activation/gradient dtype {t.dtype}') from None


class LinearModuleHelper(ModuleHelper):
    """torch.nn.Linear: A = [x,1]^T[x,1]/rows, G = g^T g/rows (modules.py:100-141)."""

    @property
    def a_factor_shape(self):
        d = self.module.weight.size(1) + int(self.has_bias())
        return (d, d)

    @property
    def g_factor_shape(self):
        d = self.module.weight.size(0)
        return (d, d)

    def accumulate_a(self, x, acc, scratch):
        _cabi.require_device(x)
        x = x.detach()
        x = x if x.is_contiguous() else x.contiguous()
        feat = x.size(-1)
        rows = x.numel() // feat
        lib = _cabi.load()
        need = lib.kfac_factor_linear_workspace_bytes(rows, feat, int(self.has_bias()))
        ws = scratch.get(need, x.device) if need else None
        _cabi.check(lib.kfac_factor_linear(x.data_ptr(), _dtype_code(x), rows, feat,
                                           int(self.has_bias()), 1.0 / rows, acc.data_ptr(),
                                           ws.data_ptr() if ws is not None else None, need,
                                           _cabi.stream_ptr()), 'kfac_factor_linear')

    def accumulate_g(self, g, acc, grad_scale, scratch=None):
        _cabi.require_device(g)
        g = g.detach()
        g = g if g.is_contiguous() else g.contiguous()
        feat = g.size(-1)
        rows = g.numel() // feat
        lib = _cabi.load()
        need = lib.kfac_factor_linear_workspace_bytes(rows, feat, 0) if scratch is not None else 0
        ws = scratch.get(need, g.device) if need else None
        _cabi.check(lib.kfac_factor_linear(g.data_ptr(), _dtype_code(g), rows, feat, 0,
                                           1.0 / (rows * grad_scale * grad_scale), acc.data_ptr(),
                                           ws.data_ptr() if ws is not None else None, need,
                                           _cabi.stream_ptr()), 'kfac_factor_linear')


class Conv2dModuleHelper(ModuleHelper):
    """torch.nn.Conv2d (modules.py:144-237).  Like the reference, dilation and
    groups are ignored; feature order is (C_in, kh, kw) [+ ones]."""

    @property
    def a_factor_shape(self):
        m = self.module
        d = m.in_channels * m.kernel_size[0] * m.kernel_size[1] + int(self.has_bias())
        return (d, d)

    @property
    def g_factor_shape(self):
        return (self.module.out_channels, self.module.out_channels)

    def _geometry(self, x):
        m = self.module
        pad = m.padding
        if isinstance(pad, str):
            raise ValueError('string padding modes are not supported by K-FAC conv factors')
        B, Cc, H, W = x.shape
        return (B, Cc, H, W, m.kernel_size[0], m.kernel_size[1], m.stride[0], m.stride[1],
                pad[0], pad[1])

    def accumulate_a(self, x, acc, scratch):
        _cabi.require_device(x)
        x = x.detach()
        x = x if x.is_contiguous() else x.contiguous()
        geo = self._geometry(x)
        B, Cc, H, W, kh, kw, sh, sw, ph, pw = geo
        Ho, Wo = (H + 2 * ph - kh) // sh + 1, (W + 2 * pw - kw) // sw + 1
        S = Ho * Wo
        # rows = B*S, features scaled by 1/S  ->  scale = 1/(B*S) * 1/S^2   (modules.py:170-178)
        scale = 1.0 / (float(B) * float(S) ** 3)
        lib = _cabi.load()
        ones = int(self.has_bias())
        need = lib.kfac_factor_conv2d_input_workspace_bytes(*geo, ones)
        ws = scratch.get(need, x.device)
        _cabi.check(lib.kfac_factor_conv2d_input(
            x.data_ptr(), _dtype_code(x), *geo, ones, scale, acc.data_ptr(),
            ws.data_ptr() if need else None, need, _cabi.stream_ptr()), 'kfac_factor_conv2d_input')

    def accumulate_g(self, g, acc, grad_scale, scratch=None):
        _cabi.require_device(g)
        g = g.detach()
        g = g if g.is_contiguous() else g.contiguous()
        B, Cc, Ho, Wo = g.shape
        S = Ho * Wo
        scale = 1.0 / (float(B) * float(S) ** 3 * grad_scale * grad_scale)
        lib = _cabi.load()
        need = lib.kfac_factor_conv2d_gradout_workspace_bytes(B, Cc, Ho, Wo) if scratch is not None else 0
        ws = scratch.get(need, g.device) if need else None
        _cabi.check(lib.kfac_factor_conv2d_gradout(g.data_ptr(), _dtype_code(g), B, Cc, Ho, Wo,
                                                   scale, acc.data_ptr(),
                                                   ws.data_ptr() if ws is not None else None, need,
                                                   _cabi.stream_ptr()),
                    'kfac_factor_conv2d_gradout')


LINEAR_TYPES: tuple[type, ...] = (torch.nn.Linear,)
CONV2D_TYPES: tuple[type, ...] = (torch.nn.Conv2d,)


def get_module_helper(module: torch.nn.Module) -> ModuleHelper | None:
    if isinstance(module, LINEAR_TYPES):
        return LinearModuleHelper(module)
    if isinstance(module, CONV2D_TYPES):
        return Conv2dModuleHelper(module)
    return None


class KFACLayer:
    """Per-layer descriptor.  Attribute names follow the reference layer
    classes (`a_factor`, `g_factor`, `qa`, `qg`, `da`, `dg`, `dgda`, `a_inv`,
    `g_inv`, `grad`) so tests and tools written against them keep working."""

    def __init__(self, module: ModuleHelper, *, method, prediv_eigenvalues: bool = True,
                 grad_scaler: Callable[[], float] | Any | None = None,
                 factor_dtype: torch.dtype | None = None,
                 inv_dtype: torch.dtype = torch.float32,
                 symmetry_aware: bool = False) -> None:
        self.module = module
        self.method = method
        self.prediv_eigenvalues = prediv_eigenvalues
        if grad_scaler is not None and hasattr(grad_scaler, 'get_scale'):
            grad_scaler = grad_scaler.get_scale
        self.grad_scaler = grad_scaler
        self.factor_dtype = factor_dtype
        self.inv_dtype = inv_dtype
        self.symmetry_aware = symmetry_aware
        self.symmetric_factors = module.has_symmetric_factors()
        self.a_dim = module.a_factor_shape[0]
        self.g_dim = module.g_factor_shape[0]
        self.index = -1
        # arena views (bound by the preconditioner)
        self._a_view = self._g_view = None
        self._a_batch_view = self._g_batch_view = None
        self._has_a = self._has_g = False       # running average exists
        self._a_count = self._g_count = 0       # accumulated micro-batches
        self._a_pending = self._g_pending = False
        self._inv = {}                          # name -> storage (rows x ld4(cols)) of qa, qg, da, dg, dgda, a_inv, g_inv
        self._inv_cols = {}                     # name -> logical column count
        self._p_store = None
        self._inv_ready = set()
        self._p_view = None
        self._grad_ready = False

    def __repr__(self) -> str:
        return f'{self.__class__.__name__}({repr(self.module)})'

    # -------- reference-compatible read accessors (views into the arenas)
    def _as_factor_dtype(self, t):
        fd = self.factor_dtype
        return t if (t is None or fd in (None, torch.float32)) else t.to(fd)

    @property
    def a_factor(self):
        """Running average A (fp32 arena view; a cast copy when `factor_dtype` is not float32)."""
        return self._as_factor_dtype(self._a_view if self._has_a else None)

    @property
    def g_factor(self):
        return self._as_factor_dtype(self._g_view if self._has_g else None)

    def _inv_get(self, key):
        if key not in self._inv_ready:
            return None
        t = self._inv[key]
        t = t[:, :self._inv_cols[key]] if t.dim() == 2 else t
        # second-order data is stored in fp32; `inv_dtype=float64` readers get a cast copy (eigen.py:319-320)
        return t if self.inv_dtype == torch.float32 else t.to(self.inv_dtype)

    qa = property(lambda self: self._inv_get('qa'))
    qg = property(lambda self: self._inv_get('qg'))
    da = property(lambda self: self._inv_get('da'))
    dg = property(lambda self: self._inv_get('dg'))
    dgda = property(lambda self: self._inv_get('dgda'))
    a_inv = property(lambda self: self._inv_get('a_inv'))
    g_inv = property(lambda self: self._inv_get('g_inv'))

    @property
    def grad(self):
        """Preconditioned gradient P (g x a) while it is pending write-back."""
        return self._p_view if self._grad_ready else None

    def state_dict(self) -> dict[str, torch.Tensor | None]:
        a, g = self.a_factor, self.g_factor
        return {'A': a.clone() if a is not None else None, 'G': g.clone() if g is not None else None}

    def load_state_dict(self, state_dict: dict[str, torch.Tensor | None]) -> None:
        if 'A' not in state_dict or 'G' not in state_dict:
            raise KeyError("KFACLayer state_dict must contain keys 'A' and 'G'")
        if state_dict['A'] is not None:
            self._a_view.copy_(state_dict['A'].to(torch.float32))
            self._has_a = True
        if state_dict['G'] is not None:
            self._g_view.copy_(state_dict['G'].to(torch.float32))
            self._has_g = True

    def memory_usage(self) -> dict[str, int]:
        b = 4
        out = {
            'a_factors': self.a_dim * self.a_dim * b if self._has_a else 0,
            'g_factors': self.g_dim * self.g_dim * b if self._has_g else 0,
            'a_batch': self.a_dim * self.a_dim * b if self._a_count else 0,
            'g_batch': self.g_dim * self.g_dim * b if self._g_count else 0,
            'a_inverses': 0, 'g_inverses': 0,
        }
        for k in self._inv_ready:
            bucket = 'a_inverses' if k in ('qa', 'da', 'a_inv') else 'g_inverses'
            out[bucket] += self._inv[k].numel() * b
        return out

    def reset_batch(self) -> None:
        if self._a_batch_view is not None:
            self._a_batch_view.zero_()
            self._g_batch_view.zero_()
        self._a_count = self._g_count = 0
        self._a_pending = self._g_pending = False


def any_match(query: str, patterns: list[str]) -> bool:
    return any(re.compile(p).search(query) for p in patterns)


def register_modules(model: torch.nn.Module, skip_layers: list[str],
                     **layer_kwargs: Any) -> dict[torch.nn.Module, tuple[str, KFACLayer]]:
    """Leaf-module discovery with regex skip on name and class name
    (kfac/layers/register.py:57-95)."""
    layers: dict[torch.nn.Module, tuple[str, KFACLayer]] = {}
    for name, module in model.named_modules():
        if len(list(module.children())) != 0:
            continue
        if any_match(name, skip_layers) or any_match(module.__class__.__name__, skip_layers):
            continue
        if not all(p.requires_grad for p in module.parameters()):
            continue
        helper = get_module_helper(module)
        if helper is None:
            continue
        assert module not in layers
        layer = KFACLayer(helper, **layer_kwargs)
        layer.index = len(layers)
        layers[module] = (name, layer)
    return layers
