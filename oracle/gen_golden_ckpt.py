"""Generate tests/golden/ref_checkpoint.pt: a `state_dict()` written by the UNMODIFIED reference
(kfac/base_preconditioner.py:215-247) after three training steps of SmallConvNet, together with what the
reference computes on the step right after loading it (kfac/base_preconditioner.py:249-308).
TEST INFRASTRUCTURE -- build container only (needs /root/reference):   python oracle/gen_golden_ckpt.py"""
from __future__ import annotations

import copy
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference  # noqa: E402
from workloads import SmallConvNet  # noqa: E402

kfac = import_reference()
from kfac.layers.base import KFACBaseLayer  # noqa: E402
from kfac.preconditioner import KFACPreconditioner  # noqa: E402


def main():
    torch.manual_seed(0)
    model = SmallConvNet()
    kw = dict(damping=0.003, factor_decay=0.9, kl_clip=0.002, lr=0.05, factor_update_steps=1, inv_update_steps=2)
    pre = KFACPreconditioner(model, **kw)
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    crit = torch.nn.CrossEntropyLoss()
    torch.manual_seed(1)
    batches = [(torch.randn(4, 3, 12, 12), torch.randint(0, 5, (4,))) for _ in range(4)]
    for s in range(3):
        opt.zero_grad()
        crit(model(batches[s][0]), batches[s][1]).backward()
        pre.step()
        opt.step()
    sd = copy.deepcopy(pre.state_dict())
    weights = {k: v.clone() for k, v in model.state_dict().items()}
    # what a FRESH reference preconditioner does with this checkpoint on the next step
    model2 = SmallConvNet()
    model2.load_state_dict(weights)
    pre2 = KFACPreconditioner(model2, **kw)
    pre2.load_state_dict(copy.deepcopy(sd))
    captured = {}
    orig = KFACBaseLayer.update_grad

    def spy(self, scale=None):
        captured.setdefault('P', {})[id(self)] = self.grad.clone()
        captured['scale'] = scale
        return orig(self, scale)
    KFACBaseLayer.update_grad = spy
    try:
        model2.zero_grad()
        crit(model2(batches[3][0]), batches[3][1]).backward()
        pre2.step()
    finally:
        KFACBaseLayer.update_grad = orig
    out = {
        'kwargs': kw, 'weights': weights, 'state_dict': sd, 'batch': batches[3],
        'after': {'steps': pre2.steps, 'scale': captured['scale'],
                  'layers': {name: {'A': l.a_factor.clone(), 'G': l.g_factor.clone(), 'P': captured['P'][id(l)]}
                             for name, l in pre2._layers.values()},
                  'final_grads': {n: p.grad.clone() for n, p in model2.named_parameters()}},
    }
    path = os.path.join(ROOT, 'tests', 'golden', 'ref_checkpoint.pt')
    torch.save(out, path)
    print('wrote', path, 'state_dict keys:', sorted(sd), 'layers:', sorted(sd['layers']))


if __name__ == '__main__':
    main()
