"""Diagnostic (not a pytest): warm-started kfac_eigh_batched on a K-FAC-like factor sequence
F_{t+1} = 0.95 F_t + 0.05 X_t^T X_t / m (F_0 = I, X_t fresh samples, m < n => decaying identity
cluster + low-rank updates), for several sweep caps.  Reports the error of the damped inverse
against fp64 at the last step -- the number the stop rule of the solver has to protect.
    KFAC_EIGH_DEBUG=1 python tests/eigh_warm_probe.py n m [steps]"""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kfac_b200 import _cabi  # noqa: E402

lib = _cabi.load()
dev = torch.device('cuda:0')
n, m = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
caps = [int(c) for c in os.environ.get('PROBE_CAPS', '1,2,3,0').split(',')]


def batches():
    g = torch.Generator(device='cpu').manual_seed(7)
    scale = torch.logspace(0, -2, n).unsqueeze(0)            # graded feature scales
    mix = torch.randn(n, n, generator=g) / n ** 0.5
    for _ in range(steps):
        x = torch.relu(torch.randn(m, n, generator=g) @ mix + 0.3) * scale
        x[:, -1] = 1.0                                         # bias column
        yield x.to(dev)


for cap in caps:
    F = torch.eye(n, device=dev)
    ld = _cabi.ld4(n)
    Q = torch.zeros(n, ld, device=dev)
    QT = torch.zeros(n, ld, device=dev)
    QTprev = torch.zeros(n, ld, device=dev)
    d = torch.empty(n, device=dev)
    ns = (C.c_int * 1)(n)
    need = lib.kfac_eigh_workspace_bytes(ns, 1)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    times = []
    for t, x in enumerate(batches()):
        F = (0.95 * F + 0.05 * (x.t() @ x) / m).contiguous()
        F = ((F + F.t()) / 2).contiguous()
        warm = QTprev.data_ptr() if t > 0 else None
        items = (_cabi.EighItem * 1)(_cabi.EighItem(F.data_ptr(), Q.data_ptr(), QT.data_ptr(), d.data_ptr(), n, ld, warm))
        torch.cuda.synchronize()
        t0 = time.time()
        rc = lib.kfac_eigh_batched(items, 1, ws.data_ptr(), need, cap if t > 0 else 0, 0.0,
                                   torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        times.append((time.time() - t0) * 1e3)
        assert rc == 0, lib.kfac_last_error()
        QTprev.copy_(QT)
    F64, Q64, d64 = F.double(), Q[:, :n].double(), d.double()
    w, V = torch.linalg.eigh(F64)
    sc = float(w.max())
    res = {}
    for damp in (1e-3, 1e-5):
        ref = (V / (w.clamp(min=0) + damp * sc)) @ V.t()
        got = (Q64 / (d64 + damp * sc)) @ Q64.t()
        res[damp] = float((got - ref).norm() / ref.norm())
    orth = float((Q64.t() @ Q64 - torch.eye(n, device=dev, dtype=torch.float64)).abs().max())
    print(f'n={n} m={m} cap={cap} last-step time={times[-1]:7.1f} ms (all: {" ".join(f"{x:.0f}" for x in times)})  '
          f'f-err(1e-3)={res[1e-3]:.2e} f-err(1e-5)={res[1e-5]:.2e} orth={orth:.1e} lam_max={sc:.2e}', flush=True)
