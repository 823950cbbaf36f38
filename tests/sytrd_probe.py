"""Diagnostic: one tridiagonalisation (test entry) for `ncu --set full` captures.  python tests/sytrd_probe.py n ncta"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kfac_b200 import _cabi  # noqa: E402

lib = _cabi.load()
lib.kfac_experimental_sytrd.restype = C.c_int
lib.kfac_experimental_sytrd.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
lib.kfac_experimental_direct_workspace_bytes.restype = C.c_size_t
lib.kfac_experimental_direct_workspace_bytes.argtypes = [C.c_int]
n, ncta = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device('cuda:0')
torch.manual_seed(0)
A = torch.randn(n, n, device=dev)
F = ((A + A.t()) / 2).contiguous()
d, e, tau = torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
need = lib.kfac_experimental_direct_workspace_bytes(n)
ws = torch.empty(need, dtype=torch.uint8, device=dev)
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = lib.kfac_experimental_sytrd(F.data_ptr(), n, d.data_ptr(), e.data_ptr(), None, n, tau.data_ptr(), ws.data_ptr(), need,
                                     ncta, torch.cuda.current_stream().cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    assert rc == 0
    print(f'sytrd n={n} ncta={ncta}: {e0.elapsed_time(e1):.2f} ms')
