"""Parity at the sizes bench.py actually runs (BASELINE.json configs 3/4/5): the eigensolver at
n in {2048, 2049, 2304, 4608} on K-FAC-like factor sequences (cold and warm-started), the FULL-width
ResNet-50 at batch 32 against the CPU oracle run live on the box, and a GPT-NeoX-125M-width block.
Reference lines: kfac/layers/eigen.py:295-385."""
import copy
import ctypes as C

import pytest
import torch

from conftest import rel_fro
from kfac_like import kfac_like_sequence

pytestmark = pytest.mark.gpu

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _check_against_fp64(F, Q, d, damping=1e-3):
    """Damped-inverse error against an fp64 eigendecomposition (computed on the GPU by torch, test
    infrastructure) + orthogonality + residual."""
    n = F.shape[0]
    F64, Q64, d64 = F.double(), Q.double(), d.double()
    assert torch.isfinite(Q64).all() and torch.isfinite(d64).all()
    assert float(d64.min()) >= 0.0
    eye = torch.eye(n, device=F.device, dtype=torch.float64)
    orth = float((Q64.t() @ Q64 - eye).abs().max())
    w, V = torch.linalg.eigh(F64)
    sc = float(w.abs().max())
    res = float((F64 @ Q64 - Q64 * d64).norm() / F64.norm())
    ev = float((torch.sort(d64).values - w.clamp(min=0)).abs().max() / sc)
    ref = (V / (w.clamp(min=0) + damping * sc)) @ V.t()
    got = (Q64 / (d64 + damping * sc)) @ Q64.t()
    ferr = float((got - ref).norm() / ref.norm())
    return {'orth': orth, 'residual': res, 'eigval': ev, 'f_err': ferr}


@pytest.mark.parametrize('n,m', [(2048, 1568), (2049, 32), (2304, 6272), (4608, 1568)])
def test_eigh_bench_sizes_cold_and_warm(n, m):
    """Rows m as in ResNet-50 bs32: layer4 3x3 convs see 32*7*7 = 1568 patch rows, layer3 6272, the fc layer 32."""
    from kfac_b200 import _cabi
    lib = _cabi.load()
    dev = torch.device('cuda:0')
    ld = _cabi.ld4(n)
    Q = torch.zeros(n, ld, device=dev)
    QT = torch.zeros(n, ld, device=dev)
    QTprev = torch.zeros(n, ld, device=dev)
    d = torch.empty(n, device=dev)
    ns = (C.c_int * 1)(n)
    need = lib.kfac_eigh_workspace_bytes(ns, 1)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    for t, F in enumerate(kfac_like_sequence(n, min(m, n), 3, dev)):
        warm = QTprev.data_ptr() if t > 0 else None       # step 0 is the cold solve
        items = (_cabi.EighItem * 1)(_cabi.EighItem(F.data_ptr(), Q.data_ptr(), QT.data_ptr(), d.data_ptr(), n, ld, warm))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.kfac_eigh_batched(items, 1, ws.data_ptr(), need, 0, 0.0, s)
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0, lib.kfac_last_error()
        r = _check_against_fp64(F, Q[:, :n], d)
        print(f'n={n} m={m} step={t} ({"warm" if t else "cold"}) {e0.elapsed_time(e1):8.1f} ms  {r}')
        assert r['orth'] < 2e-4, r
        assert r['f_err'] < 1e-3, r
        assert r['eigval'] < 5e-5, r
        assert torch.equal(QT[:, :n], Q[:, :n].t())
        QTprev.copy_(QT)


def test_eigh_mixed_batch_one_call():
    """One call over a mixed list -- the two shared-memory Jacobi classes, direct-solver sizes around every internal
    boundary (leaf 64, dense/structured merges at 1024, 256-vector block reflectors) and a 6145-wide factor
    (larger than anything in ResNet-50, not a multiple of 64): every result is checked against fp64."""
    from kfac_b200 import _cabi
    lib = _cabi.load()
    dev = torch.device('cuda:0')
    dims = [6145, 40, 128, 129, 191, 256, 257, 1000, 1025, 1088, 2047, 64]
    mats = [next(iter(kfac_like_sequence(n, max(8, min(n, 512)), 1, dev))) for n in dims]
    items = (_cabi.EighItem * len(dims))()
    outs = []
    for i, (n, F) in enumerate(zip(dims, mats)):
        ld = _cabi.ld4(n)
        Q, QT, d = torch.zeros(n, ld, device=dev), torch.zeros(n, ld, device=dev), torch.empty(n, device=dev)
        outs.append((Q, QT, d))
        items[i] = _cabi.EighItem(F.data_ptr(), Q.data_ptr(), QT.data_ptr(), d.data_ptr(), n, ld, None)
    ns = (C.c_int * len(dims))(*dims)
    need = lib.kfac_eigh_workspace_bytes(ns, len(dims))
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    for rep in range(2):          # the second call reuses a dirty workspace
        assert lib.kfac_eigh_batched(items, len(dims), ws.data_ptr(), need, 0, 0.0, s) == 0, lib.kfac_last_error()
    host = torch.zeros(1, dtype=torch.int32).pin_memory()
    assert lib.kfac_eigh_status(ws.data_ptr(), host.data_ptr(), s) == 0
    torch.cuda.synchronize()
    assert int(host[0]) == 0
    for n, F, (Q, QT, d) in zip(dims, mats, outs):
        r = _check_against_fp64(F, Q[:, :n], d)
        print(f'mixed batch n={n}: {r}')
        assert r['orth'] < 2e-4 and r['f_err'] < 1e-3 and r['eigval'] < 5e-5, (n, r)
        assert torch.equal(QT[:, :n], Q[:, :n].t())


def _parity_full(make_model, batches, loss_fn, **kw):
    """Full-model parity vs the CPU oracle: same raw gradients fed to both, per-layer A, G, P compared."""
    from kfac_b200.preconditioner import KFACPreconditioner
    from oracle.kfac_oracle import OraclePreconditioner
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    ref_model = make_model()
    gpu_model = copy.deepcopy(ref_model).to(dev)
    ref = OraclePreconditioner(ref_model, **kw)
    pre = KFACPreconditioner(gpu_model, **kw)
    opt_r = torch.optim.SGD(ref_model.parameters(), lr=0.01)
    worst, worst_at = 0.0, None
    for s, (x, y) in enumerate(batches):
        ref_model.zero_grad()
        gpu_model.zero_grad()
        loss_fn(ref_model(x), y).backward()
        loss_fn(gpu_model(x.to(dev)), y.to(dev)).backward()
        for p, q in zip(ref_model.parameters(), gpu_model.parameters()):
            q.grad.copy_(p.grad.to(dev))
        ref.step()
        pre.step()
        torch.cuda.synchronize()
        ref_layers = {L.name: L for L in ref.layers.values()}
        for name, layer in pre._layers.values():
            R = ref_layers[name]
            ea, eg = rel_fro(layer.a_factor, R.A), rel_fro(layer.g_factor, R.G)
            assert ea < 1e-4 and eg < 1e-4, (s, name, ea, eg)
            e = rel_fro(layer._p_view, R.P)
            if e > worst:
                worst, worst_at = e, (s, name, layer.a_dim, layer.g_dim)
            assert e < 1e-3, (s, name, 'P', e)
        assert abs(pre._compute_grad_scale() - ref.last_scale) <= 1e-3 * ref.last_scale
        opt_r.step()
        for p, q in zip(ref_model.parameters(), gpu_model.parameters()):
            q.data.copy_(p.data.to(dev))
        for (_, b), (_, c) in zip(ref_model.named_buffers(), gpu_model.named_buffers()):
            c.data.copy_(b.data.to(dev))
    return worst, worst_at


def test_resnet50_full_width_parity():
    """BASELINE.json configs[2]/[3] model: torchvision-shaped ResNet-50, batch 32, 224x224, two steps on two
    different batches (the second step is warm-started and sees an EMA of two different statistics)."""
    from workloads import resnet50
    torch.manual_seed(5)
    batches = [(torch.randn(32, 3, 224, 224), torch.randint(0, 1000, (32,))) for _ in range(2)]
    worst, at = _parity_full(resnet50, batches, torch.nn.CrossEntropyLoss(),
                             damping=0.001, factor_decay=0.95, kl_clip=0.001, lr=0.1)
    print('resnet50 full width: worst P rel-fro vs oracle', worst, 'at', at)


def test_gpt_neox_125m_width_parity():
    """BASELINE.json configs[4] layer shapes at TP = 1 (a in {769, 3073}, g in {2304, 768, 3072}):
    one block, (batch 2, seq 256, hidden 768) token activations, two steps."""
    from workloads import NeoXStack
    torch.manual_seed(6)
    batches = [(torch.randn(2, 256, 768), torch.randn(2, 256, 768)) for _ in range(2)]
    worst, at = _parity_full(lambda: NeoXStack(1), batches, torch.nn.MSELoss(), damping=0.003)
    print('gpt-neox-125m width: worst P rel-fro vs oracle', worst, 'at', at)
