"""Generate tests/golden/*.pt by running the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py

Every fixture stores the inputs (model state_dict, x, y) together with the
reference's per-step outputs so that the GPU box (where /root/reference does
not exist) can replay them through the CUDA path and the oracle.
"""
from __future__ import annotations

import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference  # noqa: E402
from oracle.fixture_util import loss_by_name, materialize_kwargs  # noqa: E402
from oracle.models import MODEL_ZOO, GeomConvNet, SeqModel, SmallConvNet, TinyModel  # noqa: E402

kfac = import_reference()
from kfac.assignment import KAISAAssignment  # noqa: E402
from kfac.layers.base import KFACBaseLayer  # noqa: E402
from kfac.preconditioner import KFACPreconditioner  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(GOLD, exist_ok=True)


def run_reference(model, batches, loss_fn, steps, micro=1, **kw):
    """Run `steps` optimisation steps; record per-step per-layer tensors."""
    pre = KFACPreconditioner(model, **kw)
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    record = []
    captured = {}
    orig_update = KFACBaseLayer.update_grad

    def spy(self, scale=None):
        captured.setdefault('P', {})[id(self)] = self.grad.clone()
        captured['scale'] = scale
        return orig_update(self, scale)

    KFACBaseLayer.update_grad = spy
    try:
        bi = 0
        for s in range(steps):
            opt.zero_grad()
            for _ in range(micro):
                x, y = batches[bi % len(batches)]
                bi += 1
                loss_fn(model(x), y).backward()
            raw = {n: p.grad.clone() for n, p in model.named_parameters()}
            pre.step()
            step_rec = {'scale': captured['scale'], 'raw_grads': raw,
                        'final_grads': {n: p.grad.clone() for n, p in model.named_parameters()},
                        'layers': {}}
            for name, layer in pre._layers.values():
                step_rec['layers'][name] = {
                    'A': layer.a_factor.clone(), 'G': layer.g_factor.clone(),
                    'P': captured['P'][id(layer)],
                }
            record.append(step_rec)
            opt.step()
    finally:
        KFACBaseLayer.update_grad = orig_update
    return record


def fixture(name, make_model, batches, loss_fn, steps, micro=1, **kw):
    torch.manual_seed(0)
    model = make_model()
    init = {k: v.clone() for k, v in model.state_dict().items()}
    loss_name = loss_fn if isinstance(loss_fn, str) else None
    if loss_name:
        loss_fn = loss_by_name(loss_name)
    rec = run_reference(model, batches, loss_fn, steps, micro=micro, **materialize_kwargs(kw))
    kw_ser = {k: (v if not hasattr(v, 'name') else v.name) for k, v in kw.items()}
    assert make_model.__name__ in MODEL_ZOO
    torch.save({'init': init, 'batches': batches, 'steps': steps, 'micro': micro,
                'kwargs': kw_ser, 'record': rec, 'model': make_model.__name__,
                'loss': loss_name}, os.path.join(GOLD, name + '.pt'))
    print('wrote', name, os.path.getsize(os.path.join(GOLD, name + '.pt')) // 1024, 'KiB')


def main():
    mse = torch.nn.MSELoss(reduction='sum')
    ce = torch.nn.CrossEntropyLoss()
    # ---- config 0 of BASELINE.json: TinyModel (testing/models.py:13-31),
    # inputs as tests/training_test.py:15-55
    torch.manual_seed(1)
    tiny_batches = [(torch.rand(4, 10), torch.rand(4, 10)) for _ in range(3)]
    fixture('tiny_eigen', TinyModel, tiny_batches, mse, steps=8,
            factor_update_steps=1, inv_update_steps=1, damping=0.001, lr=0.1)
    fixture('tiny_eigen_noprediv', TinyModel, tiny_batches, mse, steps=4,
            compute_eigenvalue_outer_product=False)
    fixture('tiny_inverse', TinyModel, tiny_batches, mse, steps=4,
            compute_method='inverse')
    fixture('tiny_sched', TinyModel, tiny_batches, mse, steps=7,
            factor_update_steps=2, inv_update_steps=4)
    # ---- conv + linear with every geometry the hot path has to cover
    torch.manual_seed(2)
    conv_batches = [(torch.randn(3, 3, 12, 12), torch.randint(0, 5, (3,)))
                    for _ in range(2)]
    fixture('conv_eigen', SmallConvNet, conv_batches, ce, steps=4,
            damping=0.003, factor_decay=0.9)
    fixture('conv_inverse', SmallConvNet, conv_batches, ce, steps=3,
            compute_method='inverse', damping=0.003)
    fixture('conv_accum', SmallConvNet, conv_batches, ce, steps=3, micro=2,
            accumulation_steps=2, damping=0.003)

    # ---- widened cases (round 1): per-axis conv geometry, (batch, seq, feature) Linear inputs,
    # callable hyper-parameters, skip_layers
    torch.manual_seed(3)
    geom_batches = [(torch.randn(2, 2, 11, 9), torch.randint(0, 4, (2,))) for _ in range(2)]
    fixture('geom_eigen', GeomConvNet, geom_batches, 'ce', steps=3, damping=0.003)
    fixture('geom_inverse', GeomConvNet, geom_batches, 'ce', steps=2, damping=0.003,
            compute_method='inverse')
    torch.manual_seed(4)
    seq_batches = [(torch.randn(2, 5, 12), torch.randn(2, 5, 7)) for _ in range(2)]
    fixture('seq_eigen', SeqModel, seq_batches, 'mse_sum', steps=3, damping=0.002)
    fixture('tiny_callable', TinyModel, tiny_batches, 'mse_sum', steps=6,
            damping={'schedule': [0.004, 0.003, 0.002, 0.001]},
            factor_decay={'schedule': [0.9, 0.95]},
            kl_clip={'schedule': [0.002, 0.001, 0.0005]},
            lr={'schedule': [0.2, 0.1, 0.05]},
            inv_update_steps={'schedule': [1, 1, 2]})
    fixture('conv_skip', SmallConvNet, conv_batches, 'ce', steps=3, damping=0.003,
            skip_layers=['conv3', 'Linear'])
    # factors updated in step() instead of in the hooks (base_preconditioner.py:323-333), with
    # gradient accumulation; and a conv net on a factor/inverse schedule
    fixture('conv_nohook', SmallConvNet, conv_batches, 'ce', steps=3, micro=2, accumulation_steps=2,
            damping=0.003, update_factors_in_hook=False)
    fixture('conv_sched', SmallConvNet, conv_batches, 'ce', steps=6, damping=0.003,
            factor_update_steps=2, inv_update_steps=4)

    # ---- KAISA placement (kfac/assignment.py:227-395)
    ident = lambda ranks: tuple(ranks)  # noqa: E731
    table = []
    r50 = [(147, 64), (64, 64)] + [(64, 256)] * 4 + [(256, 64)] * 2 + [(576, 64)] * 3 + \
          [(128, 512)] * 4 + [(256, 128), (256, 512)] + [(512, 128)] * 3 + [(1152, 128)] * 4 + \
          [(256, 1024)] * 6 + [(512, 256), (512, 1024)] + [(1024, 256)] * 5 + [(2304, 256)] * 6 + \
          [(512, 2048)] * 3 + [(1024, 512), (1024, 2048)] + [(2048, 512)] * 2 + \
          [(4608, 512)] * 3 + [(2049, 1000)]
    works = {
        'tiny': {'linear1': {'A': 10 ** 3, 'G': 20 ** 3}, 'linear2': {'A': 21 ** 3, 'G': 10 ** 3}},
        'ties': {f'l{i}': {'A': 8.0, 'G': 8.0} for i in range(7)},
        'r50': {f'layer{i:02d}': {'A': float(a) ** 3, 'G': float(g) ** 3}
                for i, (a, g) in enumerate(r50)},
    }
    for wname, work in works.items():
        for world in (1, 2, 4, 8, 16):
            for gw in sorted({1, 2, world // 2 or 1, world}):
                if world % gw:
                    continue
                frac = gw / world
                for colocate in (True, False):
                    for rank in sorted({0, world - 1}):
                        a = KAISAAssignment(work, local_rank=rank, world_size=world,
                                            grad_worker_fraction=frac, group_func=ident,
                                            colocate_factors=colocate)
                        table.append({
                            'work': wname, 'world': world, 'fraction': frac,
                            'colocate': colocate, 'rank': rank,
                            'inv': {l: {f: a.inv_worker(l, f) for f in a.get_factors(l)}
                                    for l in a.get_layers()},
                            'is_grad_worker': {l: a.is_grad_worker(l) for l in a.get_layers()},
                            'src_grad_worker': {l: a.src_grad_worker(l) for l in a.get_layers()},
                            'grad_worker_group': {l: sorted(a.grad_worker_group(l)) for l in a.get_layers()},
                            'grad_receiver_group': {l: sorted(a.grad_receiver_group(l)) for l in a.get_layers()},
                            'bcast_grads': a.broadcast_gradients(),
                            'bcast_invs': a.broadcast_inverses(),
                        })
    with open(os.path.join(GOLD, 'kaisa_assignment.json'), 'w') as f:
        json.dump({'works': works, 'table': table}, f)
    print('wrote kaisa_assignment.json', len(table), 'cases')


if __name__ == '__main__':
    main()
