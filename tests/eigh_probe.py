"""Diagnostic (not a pytest): accuracy / sweep counts of kfac_eigh_batched on hard spectra.
    KFAC_EIGH_DEBUG=1 python tests/eigh_probe.py [n ...]"""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gpu_kernels import make_psd  # noqa: E402
from kfac_b200 import _cabi  # noqa: E402

lib = _cabi.load()
dev = torch.device('cuda:0')
sizes = [int(a) for a in sys.argv[1:]] or [1024, 2304]
for n in sizes:
    for kind in ('geo', 'cov', 'cluster', 'lowrank'):
        F = make_psd(n, kind, n + 1).to(dev).contiguous()
        ld = _cabi.ld4(n)
        Q = torch.zeros(n, ld, device=dev)
        QT = torch.zeros(n, ld, device=dev)
        d = torch.empty(n, device=dev)
        ns = (C.c_int * 1)(n)
        need = lib.kfac_eigh_workspace_bytes(ns, 1)
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        items = (_cabi.EighItem * 1)(_cabi.EighItem(F.data_ptr(), Q.data_ptr(), QT.data_ptr(), d.data_ptr(), n, ld, None))
        torch.cuda.synchronize()
        t0 = time.time()
        rc = lib.kfac_eigh_batched(items, 1, ws.data_ptr(), need, 0, 0.0, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        dt = time.time() - t0
        assert rc == 0, lib.kfac_last_error()
        F64, Q64, d64 = F.double(), Q[:, :n].double(), d.double()
        w, V = torch.linalg.eigh(F64)
        sc = float(w.max())
        res = {}
        for damp in (1e-3, 1e-5):
            ref = (V / (w.clamp(min=0) + damp * sc)) @ V.t()
            got = (Q64 / (d64 + damp * sc)) @ Q64.t()
            res[damp] = float((got - ref).norm() / ref.norm())
        orth = float((Q64.t() @ Q64 - torch.eye(n, device=dev, dtype=torch.float64)).abs().max())
        print(f'n={n} {kind:8s} time={dt*1e3:7.1f} ms  f-err(damp1e-3)={res[1e-3]:.2e} f-err(1e-5)={res[1e-5]:.2e} '
              f'orth={orth:.1e} lam_err={float((torch.sort(d64).values - w.clamp(min=0)).abs().max()) / sc:.1e}', flush=True)
