"""Stage tests of the direct eigensolver (csrc/sytrd.cu, stedc.cu, eigh_direct.cu) through test-only C entries:
tridiagonalisation F = H T H^T and divide & conquer T = Z L Z^T, each against fp64 linear algebra."""
import ctypes as C

import pytest
import torch

from kfac_like import kfac_like_sequence

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def lib():
    from kfac_b200 import _cabi
    lib = _cabi.load()
    lib.kfac_stage_sytrd.restype = C.c_int
    lib.kfac_stage_sytrd.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    lib.kfac_stage_stedc.restype = C.c_int
    lib.kfac_stage_stedc.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                            C.c_void_p]
    lib.kfac_stage_direct_workspace_bytes.restype = C.c_size_t
    lib.kfac_stage_direct_workspace_bytes.argtypes = [C.c_int]
    return lib


def S():
    return torch.cuda.current_stream().cuda_stream


def sym(n, seed, kind):
    g = torch.Generator().manual_seed(seed)
    if kind == 'rand':
        A = torch.randn(n, n, generator=g)
        return ((A + A.t()) / 2).contiguous()
    if kind == 'kfac':
        return list(kfac_like_sequence(n, max(8, n // 3), 3, 'cpu', seed))[-1]
    if kind == 'diag':
        return torch.diag(torch.rand(n, generator=g))
    raise ValueError(kind)


def run_sytrd(lib, F, ncta):
    dev = torch.device('cuda:0')
    n = F.shape[0]
    Fd = F.to(dev).contiguous()
    d = torch.zeros(n, device=dev)
    e = torch.zeros(n, device=dev)
    tau = torch.zeros(n, device=dev)
    VT = torch.zeros(n, n, device=dev)
    need = lib.kfac_stage_direct_workspace_bytes(n)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        e0.record()
        rc = lib.kfac_stage_sytrd(Fd.data_ptr(), n, d.data_ptr(), e.data_ptr(), VT.data_ptr(), n, tau.data_ptr(),
                                         ws.data_ptr(), need, ncta, S())
        e1.record()
        assert rc == 0, lib.kfac_last_error()
        torch.cuda.synchronize()
    return d, e, tau, VT, e0.elapsed_time(e1)


@pytest.mark.parametrize('n,ncta,kind', [(2, 1, 'rand'), (3, 1, 'rand'), (33, 1, 'rand'), (64, 2, 'rand'), (65, 3, 'rand'),
                                         (130, 1, 'rand'), (200, 7, 'kfac'), (257, 5, 'rand'), (576, 0, 'kfac'),
                                         (1000, 0, 'rand'), (1024, 12, 'kfac'), (777, 148, 'diag'), (2049, 0, 'kfac'),
                                         (2304, 24, 'kfac'), (4608, 36, 'kfac'), (4608, 74, 'kfac'), (4608, 0, 'kfac')])
def test_sytrd(lib, n, ncta, kind):
    F = sym(n, n, kind)
    d, e, tau, VT, ms = run_sytrd(lib, F, ncta)
    dev = d.device
    F64 = F.double().to(dev)
    # Q_H = H_0 H_1 ... H_{n-2}, applied to the identity in fp64
    Q = torch.eye(n, dtype=torch.float64, device=dev)
    V = VT.double()
    t = tau.double()
    for j in range(n - 2, -1, -1):          # Q = H_0 (H_1 (... H_{n-2} I))
        v = V[j]
        Q -= t[j] * torch.outer(v, v @ Q)
    T = torch.diag(d.double()) + torch.diag(e.double()[:n - 1], 1) + torch.diag(e.double()[:n - 1], -1)
    rec = float((Q @ T @ Q.t() - F64).norm() / F64.norm())
    orth = float((Q.t() @ Q - torch.eye(n, dtype=torch.float64, device=dev)).abs().max())
    ev = float((torch.linalg.eigvalsh(T) - torch.linalg.eigvalsh(F64)).abs().max() / F64.norm())
    print(f'sytrd n={n} ncta={ncta} {kind}: {ms:.2f} ms  reconstruction {rec:.2e}  orth(H) {orth:.2e}  eig {ev:.2e}')
    assert orth < 1e-4
    assert rec < 2e-5
    assert ev < 1e-5


@pytest.mark.parametrize('n,kind', [(65, 'rand'), (100, 'rand'), (128, 'rand'), (200, 'kfac'), (576, 'kfac'), (1000, 'rand'),
                                    (2049, 'kfac'), (1024, 'equal'), (4608, 'kfac'), (300, 'zero_e'),
                                    # merges of more than 1024 rows take the structured (packed, two-halves) product
                                    (2304, 'rand'), (2500, 'equal'), (3000, 'zero_e'), (2100, 'glued'), (4608, 'rand')])
def test_stedc(lib, n, kind):
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(n)
    if kind in ('rand',):
        d = torch.randn(n, generator=g)
        e = torch.randn(n - 1, generator=g)
    elif kind == 'equal':
        d = torch.full((n,), 0.7)
        e = torch.zeros(n - 1)
        e[::7] = 1e-3
    elif kind == 'zero_e':
        d = torch.rand(n, generator=g)
        e = torch.zeros(n - 1)
    elif kind == 'glued':   # pairs of nearly equal eigenvalues on both sides of every cut: many deflating rotations
        d = (torch.arange(n) % 37).float() * 0.1 + 1e-7 * torch.randn(n, generator=g)
        e = 1e-4 * torch.rand(n - 1, generator=g)
    else:   # tridiagonal of a K-FAC-like factor (fp64 Householder on the host side of the test)
        F = sym(n, n, 'kfac').double()
        import scipy.linalg
        H = scipy.linalg.hessenberg(F.numpy())
        d = torch.tensor(H.diagonal().copy()).float()
        e = torch.tensor(H.diagonal(-1).copy()).float()
    dd, ed = d.to(dev), e.to(dev)
    ev = torch.zeros(n, device=dev)
    Q = torch.zeros(n, n, device=dev)
    need = lib.kfac_stage_direct_workspace_bytes(n)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        e0.record()
        rc = lib.kfac_stage_stedc(dd.data_ptr(), ed.data_ptr(), n, ev.data_ptr(), Q.data_ptr(), ws.data_ptr(), need, S())
        e1.record()
        assert rc == 0, lib.kfac_last_error()
        torch.cuda.synchronize()
    T = torch.diag(d.double()) + torch.diag(e.double(), 1) + torch.diag(e.double(), -1)
    T = T.to(dev)
    Q64, w = Q.double(), ev.double()
    orth = float((Q64.t() @ Q64 - torch.eye(n, dtype=torch.float64, device=dev)).abs().max())
    res = float((T @ Q64 - Q64 * w).norm() / T.norm())
    wr = torch.linalg.eigvalsh(T)
    eg = float((w - wr).abs().max() / wr.abs().max())
    print(f'stedc n={n} {kind}: {e0.elapsed_time(e1):.2f} ms  orth {orth:.2e}  residual {res:.2e}  eig {eg:.2e}')
    assert bool((w[1:] >= w[:-1]).all())
    assert orth < 5e-5 and res < 5e-5 and eg < 5e-6
