"""Helpers shared by oracle/gen_golden.py and the fixture replay in tests/conftest.py.
TEST INFRASTRUCTURE (never imported by the product)."""
from __future__ import annotations

import torch


def materialize_kwargs(kw: dict) -> dict:
    """Fixture kwargs are stored in a picklable form: {'schedule': [v0, v1, ...]} stands for a
    callable hyper-parameter `lambda step: v[min(step, len(v) - 1)]` (the reference accepts
    `Callable[[int], float]` for damping / factor_decay / kl_clip / lr / *_update_steps,
    kfac/preconditioner.py:54-87)."""
    out = {}
    for k, v in kw.items():
        if isinstance(v, dict) and 'schedule' in v:
            vals = list(v['schedule'])
            out[k] = (lambda vals: (lambda step: vals[min(int(step), len(vals) - 1)]))(vals)
        else:
            out[k] = v
    return out


def loss_by_name(name: str):
    if name == 'mse_sum':
        return torch.nn.MSELoss(reduction='sum')
    if name == 'ce':
        return torch.nn.CrossEntropyLoss()
    raise KeyError(name)
