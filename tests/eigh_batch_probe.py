"""Diagnostic (not a pytest): one kfac_eigh_batched call over the ResNet-50 factor dimensions
on K-FAC-like matrices; prints the CUDA-event time of the call.  Used under `ncu` for launch lists.
    python tests/eigh_batch_probe.py [reps] [dims: r50 | comma list]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kfac_b200 import _cabi  # noqa: E402

R50 = [(147, 64, 1), (64, 64, 1), (64, 256, 4), (256, 64, 2), (576, 64, 3), (128, 512, 4), (256, 128, 1), (256, 512, 1),
       (512, 128, 3), (1152, 128, 4), (256, 1024, 6), (512, 256, 1), (512, 1024, 1), (1024, 256, 5), (2304, 256, 6),
       (512, 2048, 3), (1024, 512, 1), (1024, 2048, 1), (2048, 512, 2), (4608, 512, 3), (2049, 1000, 1)]

lib = _cabi.load()
dev = torch.device('cuda:0')
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
spec = sys.argv[2] if len(sys.argv) > 2 else 'r50'
if spec == 'r50':
    dims = [d for a, g, c in R50 for d in [a, g] * c]
else:
    dims = [int(x) for x in spec.split(',')]
torch.manual_seed(0)
mats = []
for n in dims:
    m = max(8, n // 3)
    x = torch.relu(torch.randn(m, n, device=dev) + 0.3) * torch.logspace(0, -2, n, device=dev)
    F = 0.5 * torch.eye(n, device=dev) + 0.5 * (x.t() @ x) / m
    mats.append(((F + F.t()) / 2).contiguous())
items = (_cabi.EighItem * len(dims))()
ns = (C.c_int * len(dims))(*dims)
keep = []
for i, (n, F) in enumerate(zip(dims, mats)):
    ld = _cabi.ld4(n)
    Q, QT, d = torch.zeros(n, ld, device=dev), torch.zeros(n, ld, device=dev), torch.zeros(n, device=dev)
    keep.append((Q, QT, d))
    items[i] = _cabi.EighItem(F.data_ptr(), Q.data_ptr(), QT.data_ptr(), d.data_ptr(), n, ld, None)
need = lib.kfac_eigh_workspace_bytes(ns, len(dims))
ws = torch.empty(need, dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
print(f'{len(dims)} matrices, sum n^3 = {sum(n ** 3 for n in dims):.3e}, workspace {need / 1e9:.2f} GB', flush=True)
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.kfac_launch_count()
    e0.record()
    rc = lib.kfac_eigh_batched(items, len(dims), ws.data_ptr(), need, 0, 0.0, s)
    e1.record()
    torch.cuda.synchronize()
    assert rc == 0, lib.kfac_last_error()
    print(f'rep {r}: {e0.elapsed_time(e1):.2f} ms, {lib.kfac_launch_count() - l0} launches', flush=True)
worst = 0.0
for (Q, QT, d), F, n in zip(keep, mats, dims):
    if n > 1200:
        continue
    Q64 = Q[:, :n].double()
    res = float((F.double() @ Q64 - Q64 * d.double()).norm() / F.double().norm())
    worst = max(worst, res)
print(f'worst residual (n <= 1200): {worst:.2e}')
