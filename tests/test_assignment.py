"""KAISA placement parity: reference-generated tables + the reference's own
expected partitions (tests/assignment_test.py:61-142)."""
import json
import os

import pytest

from conftest import GOLDEN
from kfac_b200.assignment import KAISAAssignment, WorkAssignment


def fs(*groups):
    return {frozenset(g) for g in groups}


def test_partitions_reference_vectors():
    # tests/assignment_test.py:61-142 expected values
    P = KAISAAssignment.partition_grad_workers
    R = KAISAAssignment.partition_grad_receivers
    assert P(1, 1) == fs([0])
    assert P(2, 1) == fs([0], [1]) and P(2, 2) == fs([0, 1])
    assert P(8, 2) == fs([0, 4], [1, 5], [2, 6], [3, 7])
    assert P(8, 4) == fs([0, 2, 4, 6], [1, 3, 5, 7])
    assert P(8, 8) == fs(range(8))
    assert R(8, 2) == fs([0, 1, 2, 3], [4, 5, 6, 7])
    assert R(8, 4) == fs([0, 1], [2, 3], [4, 5], [6, 7])
    assert R(8, 1) == fs(range(8))
    assert R(16, 16) == fs(*[[i] for i in range(16)])
    for f in (P, R):
        with pytest.raises(ValueError):
            f(0, 1)
        with pytest.raises(ValueError):
            f(8, 3)


def test_matches_reference_tables():
    with open(os.path.join(GOLDEN, 'kaisa_assignment.json')) as f:
        gold = json.load(f)
    for case in gold['table']:
        work = gold['works'][case['work']]
        a = KAISAAssignment(work, local_rank=case['rank'], world_size=case['world'],
                            grad_worker_fraction=case['fraction'],
                            group_func=lambda r: tuple(r), colocate_factors=case['colocate'])
        key = (case['work'], case['world'], case['fraction'], case['colocate'], case['rank'])
        assert {l: {f: a.inv_worker(l, f) for f in a.get_factors(l)} for l in a.get_layers()} == case['inv'], key
        for l in a.get_layers():
            assert a.is_grad_worker(l) == case['is_grad_worker'][l], key
            assert a.src_grad_worker(l) == case['src_grad_worker'][l], key
            assert sorted(a.grad_worker_group(l)) == case['grad_worker_group'][l], key
            assert sorted(a.grad_receiver_group(l)) == case['grad_receiver_group'][l], key
        assert a.broadcast_gradients() == case['bcast_grads']
        assert a.broadcast_inverses() == case['bcast_invs']


def test_validation_errors():
    work = {'l': {'A': 1.0, 'G': 1.0}}
    kw = dict(group_func=lambda r: None)
    with pytest.raises(ValueError):
        KAISAAssignment(work, local_rank=0, world_size=4, grad_worker_fraction=1.5, **kw)
    with pytest.raises(ValueError):
        KAISAAssignment(work, local_rank=-1, world_size=4, grad_worker_fraction=1.0, **kw)
    with pytest.raises(ValueError):
        KAISAAssignment(work, local_rank=4, world_size=4, grad_worker_fraction=1.0, **kw)
    with pytest.raises(ValueError):
        KAISAAssignment(work, local_rank=0, world_size=8, grad_worker_fraction=0.33, **kw)
    assert issubclass(KAISAAssignment, WorkAssignment)
    a = KAISAAssignment(work, local_rank=0, world_size=1, grad_worker_fraction=1.0, **kw)
    assert 'layer="l"' in repr(a)
    assert a.factor_group('l', 'A') is None
