"""Experimental kernels that are NOT on the product path (off unless KFAC_TEST_EXPERIMENTAL=1):
validated/timed here before they replace a product kernel."""
import ctypes as C
import os
import time

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get('KFAC_TEST_EXPERIMENTAL') != '1',
                                 reason='experimental kernels: set KFAC_TEST_EXPERIMENTAL=1')]


@pytest.mark.parametrize('flags', [0, 1, 4])
@pytest.mark.parametrize('n', [64, 100, 128])
def test_systolic_jacobi_matches_eigh(n, flags):
    from kfac_b200 import _cabi
    from test_gpu_kernels import make_psd
    lib = _cabi.load()
    fn = lib.kfac_experimental_jacobi_systolic
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    dev = torch.device('cuda:0')
    count = 72
    F = torch.stack([make_psd(n, kind, 11 * i + n) for i, kind in
                     zip(range(count), ['cov', 'geo', 'cluster', 'lowrank'] * (count // 4))]).to(dev).contiguous()
    Q = torch.empty_like(F)
    d = torch.empty(count, n, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        assert fn(F.data_ptr(), n, count, Q.data_ptr(), d.data_ptr(), 0, flags, s) == 0, lib.kfac_last_error()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(10):
        fn(F.data_ptr(), n, count, Q.data_ptr(), d.data_ptr(), 0, flags, s)
    torch.cuda.synchronize()
    print(f'systolic flags={flags} n={n} x{count}: {(time.time() - t0) * 100:.3f} ms per launch')
    F64, Q64, d64 = F.double(), Q.double(), d.double()
    eye = torch.eye(n, device=dev, dtype=torch.float64)
    assert (Q64.transpose(1, 2) @ Q64 - eye).abs().max() < 5e-5
    rec = Q64 @ torch.diag_embed(d64) @ Q64.transpose(1, 2)
    scale = F64.abs().amax(dim=(1, 2), keepdim=True)
    assert ((rec - F64).abs() / scale).max() < 2e-5


@pytest.mark.parametrize('jopt', ['1', '2', '4', '8', '9', '16', '17', '24', '25'])
def test_eigh_with_experimental_pair_solver(jopt):
    # KFAC_EIGH_JOPT is read once per process: run the block-solver tests in a child
    import subprocess
    import sys
    env = dict(os.environ, KFAC_EIGH_JOPT=jopt)
    env.pop('KFAC_TEST_EXPERIMENTAL', None)
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'test_gpu_kernels.py')
    out = subprocess.run([sys.executable, '-m', 'pytest', here, '-q', '-m', 'gpu', '-x', '-k',
                          'eigh and not wide', '-p', 'no:cacheprovider'], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
