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


def test_oracle_operators_match_their_definitions():
    """Independent of the reference: the two K-FAC operators of the path as dense linear algebra
    (SURVEY.md App. A.11).  eigen = (G (x) A + damping I)^-1 vec(grad); inverse =
    (G + damping I)^-1 grad (A + damping I)^-1; conv factor scaling A = P^T P / (B S^3)."""
    torch.manual_seed(5)
    g, a, damping = 4, 6, 0.05
    xa, xg = torch.randn(20, a, dtype=torch.float64), torch.randn(20, g, dtype=torch.float64)
    A, G = (xa.t() @ xa / 20).float(), (xg.t() @ xg / 20).float()
    grad = torch.randn(g, a)
    da, qa = O.eigen_decompose(A)
    dg, qg = O.eigen_decompose(G)
    # row-major vec(grad): (G kron A) vec(X) = vec(G X A^T)
    K = torch.kron(G.double(), A.double()) + damping * torch.eye(g * a, dtype=torch.float64)
    want = torch.linalg.solve(K, grad.double().reshape(-1)).reshape(g, a)
    got = O.precondition_eigen(grad, qa, qg, dgda=O.eigen_dgda(dg, da, damping))
    assert rel_fro(got, want) < 1e-5
    got2 = O.precondition_eigen(grad, qa, qg, da=da, dg=dg, damping=damping)
    assert rel_fro(got2, want) < 1e-5
    want_inv = torch.linalg.inv(G.double() + damping * torch.eye(g, dtype=torch.float64)) @ grad.double() @ \
        torch.linalg.inv(A.double() + damping * torch.eye(a, dtype=torch.float64))
    got_inv = O.precondition_inverse(grad, O.damped_inverse(A, damping), O.damped_inverse(G, damping))
    assert rel_fro(got_inv, want_inv) < 1e-5
    # conv A-factor: patches (B, Ho, Wo, C kh kw) with a ones column, scaled 1 / S before the covariance
    x = torch.randn(2, 3, 5, 5)
    pt = torch.nn.functional.unfold(x, (3, 3), padding=1, stride=2)        # (B, C kh kw, S)
    B_, S = 2, pt.shape[-1]
    P = torch.cat([pt.transpose(1, 2).reshape(B_ * S, -1), torch.ones(B_ * S, 1)], 1)
    want_a = P.t() @ P / (B_ * S ** 3)
    got_a = O.conv2d_a_factor(x, (3, 3), (2, 2), (1, 1), True)
    assert rel_fro(got_a, want_a) < 1e-5
    # kl-clip scale
    P1 = [torch.randn(3, 5)]
    G1 = [torch.randn(3, 5)]
    vg = float((P1[0] * G1[0]).sum()) * 0.1 ** 2
    assert abs(O.grad_scale(P1, G1, 0.1, 0.001) - min(1.0, (0.001 / abs(vg)) ** 0.5)) < 1e-6
