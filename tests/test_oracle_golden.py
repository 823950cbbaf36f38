"""Pins the CPU oracle against the reference-generated golden fixtures and
the reference's own exact vectors (tests/layers/utils_test.py:25-72)."""
import json
import os

import pytest
import torch

from conftest import GOLDEN, GOLDEN_CPU_ONLY, GOLDEN_NAMES, rel_fro, replay
from oracle import kfac_oracle as O

NAMES = list(GOLDEN_NAMES) + list(GOLDEN_CPU_ONLY)


def test_get_cov_reference_vectors():
    # exact vectors held by the reference: tests/layers/utils_test.py:25-72
    a = torch.tensor([[1., 2, 3], [4, 5, 6], [7, 8, 9]])
    assert torch.equal(O.get_cov(a, scale=1.0),
                       torch.tensor([[66., 78, 90], [78, 93, 108], [90, 108, 126]]))
    assert torch.equal(O.get_cov(a), torch.tensor([[22., 26, 30], [26, 31, 36], [30, 36, 42]]))
    assert torch.equal(O.get_cov(torch.ones(2, 2)), torch.ones(2, 2))
    c = O.get_cov(torch.randn(7, 5))
    assert torch.equal(c, c.t())
    with pytest.raises(ValueError):
        O.get_cov(torch.ones(3))


def _mk(model, **kw):
    method = str(kw.pop('compute_method', 'eigen')).lower()
    prediv = kw.pop('compute_eigenvalue_outer_product', True)
    return O.OraclePreconditioner(model, compute_method=method, prediv=prediv, **kw)


@pytest.mark.parametrize('name', NAMES)
def test_oracle_matches_reference(name):
    worst = 0.0
    for s, gold, model, pre in replay(name, _mk):
        by_name = {L.name: L for L in pre.layers.values()}
        for lname, g in gold['layers'].items():
            L = by_name[lname]
            assert rel_fro(L.A, g['A']) < 1e-6, (s, lname, 'A')
            assert rel_fro(L.G, g['G']) < 1e-6, (s, lname, 'G')
            e = rel_fro(L.P, g['P'])
            worst = max(worst, e)
            assert e < 2e-4, (s, lname, 'P', e)
        assert abs(pre.last_scale - gold['scale']) <= 1e-4 * abs(gold['scale'])
        for n, p in model.named_parameters():
            assert rel_fro(p.grad, gold['final_grads'][n]) < 2e-4, (s, n)
    print(name, 'worst P rel-fro', worst)


def test_golden_assignment_file_is_sane():
    with open(os.path.join(GOLDEN, 'kaisa_assignment.json')) as f:
        d = json.load(f)
    assert len(d['table']) > 100


@pytest.mark.parametrize('name', ['dist_w2_comm', 'dist_w2_mem', 'dist_w4_hybrid', 'dist_w4_mem_inverse'])
def test_oracle_matches_distributed_reference(name):
    """The multi-GPU parity tests (tests/dist_parity.py) compare every rank with the single-process
    oracle on the CONCATENATED batch.  That equivalence is pinned here against the UNMODIFIED
    reference run data-parallel under gloo (oracle/gen_golden_dist.py: 2 and 4 ranks, COMM-OPT /
    HYBRID-OPT / MEM-OPT placements): every rank back-propagates the MEAN loss of its own shard, so the
    G statistics see 1/per_rank-scaled grad-outputs -> world * mean-loss on the concatenated batch,
    then undo the factor on the parameter gradients (DDP averages them)."""
    from oracle.models import SmallConvNet
    fx = torch.load(os.path.join(GOLDEN, name + '.pt'), weights_only=False)
    model = SmallConvNet()
    model.load_state_dict(fx['init'])
    pre = O.OraclePreconditioner(model, damping=fx['damping'], compute_method=fx['method'])
    crit = torch.nn.CrossEntropyLoss()
    world = fx['world']
    worst = 0.0
    for s in range(len(fx['record'])):
        model.zero_grad()
        (crit(model(fx['x']), fx['y']) * world).backward()
        for p in model.parameters():
            p.grad /= world
        pre.step()
        for n, p in model.named_parameters():
            e = rel_fro(p.grad, fx['record'][s][n])
            worst = max(worst, e)
            assert e < 2e-4, (name, s, n, e)
        with torch.no_grad():
            for p in model.parameters():
                p -= fx['lr_sgd'] * p.grad
    print(name, 'worst rel-fro oracle(concatenated batch) vs distributed reference', worst)
