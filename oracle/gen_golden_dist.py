"""Generate tests/golden/dist_*.pt: the UNMODIFIED reference run data-parallel (gloo, several
ranks, KAISA grad_worker_fraction placements) on per-rank shards of a global batch.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python oracle/gen_golden_dist.py

The fixtures pin the claim the multi-GPU parity tests rest on (tests/dist_parity.py): for a model
without BatchNorm the distributed step equals the single-process step on the CONCATENATED batch once
the per-rank mean-loss scaling is reproduced (tests/test_oracle_golden.py::test_oracle_matches_
distributed_reference replays the oracle against them on the CPU).
"""
from __future__ import annotations

import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, 'tests', 'golden')
PER_RANK = 4
STEPS = 3


def global_batch(world):
    torch.manual_seed(1)
    return torch.randn(world * PER_RANK, 3, 12, 12), torch.randint(0, 5, (world * PER_RANK,))


def worker(rank, world, port, frac, method, out_path):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle.ref_import import import_reference
    import_reference()
    from kfac.preconditioner import KFACPreconditioner
    from oracle.models import SmallConvNet
    torch.manual_seed(0)
    model = SmallConvNet()
    init = {k: v.clone() for k, v in model.state_dict().items()}
    gx, gy = global_batch(world)
    x, y = gx[rank * PER_RANK:(rank + 1) * PER_RANK], gy[rank * PER_RANK:(rank + 1) * PER_RANK]
    pre = KFACPreconditioner(model, damping=0.003, grad_worker_fraction=frac, compute_method=method)
    crit = torch.nn.CrossEntropyLoss()
    record = []
    for _ in range(STEPS):
        model.zero_grad()
        crit(model(x), y).backward()
        for p in model.parameters():          # what DDP does
            dist.all_reduce(p.grad)
            p.grad /= world
        pre.step()
        grads = {n: p.grad.clone() for n, p in model.named_parameters()}
        # every rank must hold the same preconditioned gradients
        for n, g in grads.items():
            ref = g.clone()
            dist.broadcast(ref, src=0)
            assert torch.allclose(ref, g, rtol=1e-5, atol=1e-7), (rank, n)
        record.append(grads)
        with torch.no_grad():
            for p in model.parameters():
                p -= 0.05 * p.grad
    if rank == 0:
        torch.save({'init': init, 'world': world, 'per_rank': PER_RANK, 'fraction': frac, 'method': method,
                    'x': gx, 'y': gy, 'lr_sgd': 0.05, 'damping': 0.003, 'record': record}, out_path)
        print('wrote', os.path.basename(out_path), os.path.getsize(out_path) // 1024, 'KiB', flush=True)
    dist.barrier()
    dist.destroy_process_group()


def main():
    os.makedirs(GOLD, exist_ok=True)
    port = 29611
    for world, frac, method, tag in ((2, 1.0, 'eigen', 'w2_comm'), (2, 0.5, 'eigen', 'w2_mem'),
                                     (4, 0.5, 'eigen', 'w4_hybrid'), (4, 0.25, 'inverse', 'w4_mem_inverse')):
        port += 1
        mp.spawn(worker, args=(world, port, frac, method, os.path.join(GOLD, f'dist_{tag}.pt')), nprocs=world, join=True)


if __name__ == '__main__':
    main()
