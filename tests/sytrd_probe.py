"""Diagnostic: one tridiagonalisation (test entry) for `ncu --set full` captures.  python tests/sytrd_probe.py n ncta"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kfac_b200 import _cabi  # noqa: E402

lib = _cabi.load()
lib.kfac_stage_sytrd.restype = C.c_int
lib.kfac_stage_sytrd.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
lib.kfac_stage_direct_workspace_bytes.restype = C.c_size_t
lib.kfac_stage_direct_workspace_bytes.argtypes = [C.c_int]
n, ncta = int(sys.argv[1]), int(sys.argv[2])
prof = len(sys.argv) > 3 and sys.argv[3] == 'prof'
lib.kfac_stage_sytrd_profile.restype = C.c_int
lib.kfac_stage_sytrd_profile.argtypes = [C.c_int, C.c_void_p]
NAMES = ['C scalars', 'VT + tiles', 'dots + reduce', 'barrier 1', 'B loads/sums/row s+1', 'gather smem', 'finish rows', 'barrier 2', 'update', 'barrier 3']
dev = torch.device('cuda:0')
torch.manual_seed(0)
A = torch.randn(n, n, device=dev)
F = ((A + A.t()) / 2).contiguous()
d, e, tau = torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
need = lib.kfac_stage_direct_workspace_bytes(n)
ws = torch.empty(need, dtype=torch.uint8, device=dev)
for it in range(2):
    if prof and it == 1:
        lib.kfac_stage_sytrd_profile(1, None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = lib.kfac_stage_sytrd(F.data_ptr(), n, d.data_ptr(), e.data_ptr(), None, n, tau.data_ptr(), ws.data_ptr(), need,
                                     ncta, torch.cuda.current_stream().cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    assert rc == 0
    print(f'sytrd n={n} ncta={ncta}: {e0.elapsed_time(e1):.2f} ms')
if prof:
    out = (C.c_ulonglong * 16)()
    lib.kfac_stage_sytrd_profile(0, out)
    tot = sum(out)
    for nm, v in zip(NAMES, out):
        print(f'   {nm:24s} {v / 1.965e3 / (n - 1):8.2f} us/col  {100.0 * v / max(tot, 1):5.1f} %')
    print(f'   total {tot / 1.965e3 / (n - 1):.2f} us/col (clock64 at 1.965 GHz)')
