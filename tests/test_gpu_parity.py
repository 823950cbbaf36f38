"""End-to-end parity of KFACPreconditioner (CUDA path, through the C ABI)
against (i) the reference-generated golden fixtures and (ii) the CPU oracle on
the BASELINE.json workloads.  Bar: per-layer preconditioned gradient within
1e-3 relative Frobenius (BASELINE.json north_star)."""
import copy

import pytest
import torch

from conftest import GOLDEN_NAMES, rel_fro, replay

pytestmark = pytest.mark.gpu

# compare like with like: the CPU oracle's forward/backward is fp32, so keep the
# GPU model's convolutions/matmuls out of TF32 (the K-FAC path itself never uses it
# unsplit).
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

NAMES = list(GOLDEN_NAMES)


def _mk(model, **kw):
    from kfac_b200.preconditioner import KFACPreconditioner
    return KFACPreconditioner(model, **kw)


@pytest.mark.parametrize('name', NAMES)
def test_golden_replay(name):
    dev = torch.device('cuda:0')
    worst = 0.0
    for s, gold, model, pre in replay(name, _mk, device=dev):
        torch.cuda.synchronize()
        layers = {n: l for n, l in pre._layers.values()}
        for lname, g in gold['layers'].items():
            L = layers[lname]
            assert rel_fro(L.a_factor, g['A']) < 2e-5, (s, lname, 'A')
            assert rel_fro(L.g_factor, g['G']) < 2e-5, (s, lname, 'G')
            e = rel_fro(L._p_view, g['P'])
            worst = max(worst, e)
            assert e < 1e-3, (s, lname, 'P', e)
        scale = pre._compute_grad_scale()
        assert abs(scale - gold['scale']) <= 1e-3 * abs(gold['scale']), (s, scale, gold['scale'])
        for n, p in model.named_parameters():
            assert rel_fro(p.grad, gold['final_grads'][n]) < 1e-3, (s, n)
    print(name, 'worst P rel-fro vs reference', worst)


def _parity_vs_oracle(make_model, x, y, loss_fn, steps, **kw):
    from kfac_b200.preconditioner import KFACPreconditioner
    from oracle.kfac_oracle import OraclePreconditioner
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    ref_model = make_model()
    gpu_model = copy.deepcopy(ref_model).to(dev)
    okw = dict(kw)
    method = str(okw.pop('compute_method', 'eigen')).lower()
    ref = OraclePreconditioner(ref_model, compute_method=method, **okw)
    pre = KFACPreconditioner(gpu_model, **kw)
    opt_r = torch.optim.SGD(ref_model.parameters(), lr=0.01)
    opt_g = torch.optim.SGD(gpu_model.parameters(), lr=0.01)
    xd, yd = x.to(dev), y.to(dev)
    worst = 0.0
    for s in range(steps):
        opt_r.zero_grad()
        opt_g.zero_grad()
        loss_fn(ref_model(x), y).backward()
        loss_fn(gpu_model(xd), yd).backward()
        # feed the SAME raw gradients to both so only the K-FAC path is compared
        for p, q in zip(ref_model.parameters(), gpu_model.parameters()):
            q.grad.copy_(p.grad.to(dev))
        ref.step()
        pre.step()
        torch.cuda.synchronize()
        ref_layers = {L.name: L for L in ref.layers.values()}
        for name, layer in pre._layers.values():
            R = ref_layers[name]
            assert rel_fro(layer.a_factor, R.A) < 1e-4, (s, name, 'A', rel_fro(layer.a_factor, R.A))
            assert rel_fro(layer.g_factor, R.G) < 1e-4, (s, name, 'G', rel_fro(layer.g_factor, R.G))
            e = rel_fro(layer._p_view, R.P)
            worst = max(worst, e)
            assert e < 1e-3, (s, name, 'P', e)
        assert abs(pre._compute_grad_scale() - ref.last_scale) <= 1e-3 * ref.last_scale
        opt_r.step()
        # keep the two models in lock-step
        for p, q in zip(ref_model.parameters(), gpu_model.parameters()):
            q.data.copy_(p.data.to(dev))
        for (bn, b), (_, c) in zip(ref_model.named_buffers(), gpu_model.named_buffers()):
            c.data.copy_(b.data.to(dev))
    return worst


def test_resnet32_parity():
    """BASELINE.json configs[1]: ResNet-32, CIFAR-shaped synthetic batch."""
    from oracle.models import resnet32
    torch.manual_seed(1)
    x = torch.randn(32, 3, 32, 32)
    y = torch.randint(0, 10, (32,))
    worst = _parity_vs_oracle(resnet32, x, y, torch.nn.CrossEntropyLoss(), steps=3,
                              damping=0.003, factor_decay=0.95, kl_clip=0.001, lr=0.1)
    print('resnet32 worst P rel-fro', worst)


def test_resnet32_inverse_method_parity():
    from oracle.models import resnet32
    torch.manual_seed(2)
    x = torch.randn(16, 3, 32, 32)
    y = torch.randint(0, 10, (16,))
    worst = _parity_vs_oracle(resnet32, x, y, torch.nn.CrossEntropyLoss(), steps=2,
                              damping=0.003, compute_method='inverse')
    print('resnet32 inverse worst P rel-fro', worst)


def test_bottleneck_stack_parity():
    """ResNet-50-shaped layers at reduced width (the full model runs in bench.py)."""
    from oracle.models import ResNet50
    torch.manual_seed(3)
    x = torch.randn(4, 3, 64, 64)
    y = torch.randint(0, 10, (4,))
    worst = _parity_vs_oracle(lambda: ResNet50(num_classes=10, width=16, blocks=(1, 1, 1, 1)), x, y,
                              torch.nn.CrossEntropyLoss(), steps=2, damping=0.001)
    print('bottleneck worst P rel-fro', worst)


def test_state_dict_roundtrip_and_schedule():
    from kfac_b200.preconditioner import KFACPreconditioner
    from oracle.models import TinyModel
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    m = TinyModel().to(dev)
    p = KFACPreconditioner(m, factor_update_steps=2, inv_update_steps=4)
    x = torch.rand(4, 10, device=dev)
    for _ in range(3):
        m.zero_grad()
        m(x).sum().backward()
        p.step()
    sd = p.state_dict()
    assert sd['steps'] == 3 and set(sd['layers']) == {'linear1', 'linear2'}
    assert sd['layers']['linear1']['A'].shape == (10, 10) and sd['layers']['linear2']['A'].shape == (21, 21)
    m2 = TinyModel().to(dev)
    p2 = KFACPreconditioner(m2, factor_update_steps=2, inv_update_steps=4)
    p2.load_state_dict(sd)
    assert p2.steps == 3
    for (_, a), (_, b) in zip(p._layers.values(), p2._layers.values()):
        assert torch.equal(a.a_factor, b.a_factor) and torch.equal(a.g_factor, b.g_factor)
        assert b.qa is not None and b.qg is not None     # inverses recomputed on load
    mem = p.memory_usage()
    assert mem['total'] > 0 and set(mem) >= {'a_factors', 'g_factors', 'a_inverses', 'g_inverses', 'total'}
    # hooks are inert in eval mode
    m.eval()
    before = p._layers[m.linear1][1]._a_count
    m(x)
    assert p._layers[m.linear1][1]._a_count == before


def test_step_before_factors_raises():
    from kfac_b200.preconditioner import KFACPreconditioner
    from oracle.models import TinyModel
    dev = torch.device('cuda:0')
    m = TinyModel().to(dev)
    p = KFACPreconditioner(m)
    m.eval()
    m(torch.rand(2, 10, device=dev)).sum().backward()
    with pytest.raises(RuntimeError):
        p.step()


def test_training_loss_decreases():
    """tests/training_test.py:15-55 of the reference: losses[0] > losses[-1]."""
    from kfac_b200.preconditioner import KFACPreconditioner
    from oracle.models import TinyModel
    dev = torch.device('cuda:0')
    torch.manual_seed(42)
    m = TinyModel().to(dev)
    opt = torch.optim.SGD(m.parameters(), lr=0.001)
    p = KFACPreconditioner(m, lr=0.001)
    x, y = torch.rand(4, 10, device=dev), torch.rand(4, 10, device=dev)
    crit = torch.nn.MSELoss(reduction='sum')
    losses = []
    for _ in range(20):
        opt.zero_grad()
        loss = crit(m(x), y)
        loss.backward()
        p.step()
        opt.step()
        losses.append(loss.item())
    assert losses[0] > losses[-1]


def test_multi_gpu_kaisa_parity():
    """KAISA COMM/HYBRID/MEM-OPT on every visible GPU vs the oracle on the concatenated batch."""
    import os
    import socket
    import subprocess
    import sys
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs')
    n = 2 if n < 4 else (4 if n < 8 else 8)
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'dist_parity.py')
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}',
                          '--master-addr', '127.0.0.1', '--master-port', str(port), script],
                         capture_output=True, text=True, timeout=600)
    print(out.stdout[-3000:])
    assert out.returncode == 0 and 'DIST PARITY OK' in out.stdout, out.stderr[-3000:]


def test_linear_stack_parity():
    """BASELINE.json configs[4] shapes at reduced width: a GPT-NeoX-like stack of
    biased Linear layers fed (batch*seq, hidden) token activations (TP=1: the
    math of kfac/gpt_neox/layer.py:238-247 equals kfac/layers/eigen.py:374-385)."""
    class Stack(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.qkv = torch.nn.Linear(96, 288)
            self.proj = torch.nn.Linear(288, 96)
            self.up = torch.nn.Linear(96, 384)
            self.down = torch.nn.Linear(384, 96)

        def forward(self, x):          # x: (batch, seq, hidden) -- leading dims are flattened by K-FAC
            h = torch.tanh(self.proj(torch.tanh(self.qkv(x))))
            return self.down(torch.nn.functional.gelu(self.up(h)))

    torch.manual_seed(4)
    x = torch.randn(4, 64, 96)
    y = torch.randn(4, 64, 96)
    worst = _parity_vs_oracle(Stack, x, y, torch.nn.MSELoss(), steps=3, damping=0.003)
    print('linear stack worst P rel-fro', worst)


def test_grad_scaler_unscales_g_factor():
    """AMP-style loss scaling (kfac/layers/base.py:365-366): G statistics are divided by the scale."""
    from kfac_b200.preconditioner import KFACPreconditioner
    from oracle.models import TinyModel
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    base = TinyModel().to(dev)
    x, y = torch.rand(4, 10, device=dev), torch.rand(4, 10, device=dev)
    crit = torch.nn.MSELoss(reduction='sum')
    outs = []
    for scale in (1.0, 128.0):
        m = copy.deepcopy(base)
        p = KFACPreconditioner(m, grad_scaler=(lambda s=scale: s) if scale != 1.0 else None)
        (crit(m(x), y) * scale).backward()
        for q in m.parameters():
            q.grad /= scale               # what GradScaler.unscale_ does before step()
        p.step()
        torch.cuda.synchronize()
        outs.append([q.grad.clone() for q in m.parameters()])
    for a, b in zip(*outs):
        assert rel_fro(a, b) < 1e-4


def test_cross_load_reference_checkpoint():
    """A state_dict WRITTEN BY THE UNMODIFIED REFERENCE (tests/golden/ref_checkpoint.pt, generated by
    oracle/gen_golden_ckpt.py; schema of kfac/base_preconditioner.py:215-247) loads into the CUDA preconditioner,
    and the step right after loading reproduces what the reference itself computes after loading it
    (inverses recomputed from the loaded factors on load, kfac/base_preconditioner.py:296-308)."""
    from conftest import load_fixture
    from kfac_b200.preconditioner import KFACPreconditioner
    from workloads import SmallConvNet
    fx = load_fixture('ref_checkpoint')
    dev = torch.device('cuda:0')
    model = SmallConvNet()
    model.load_state_dict(fx['weights'])
    model.to(dev)
    pre = KFACPreconditioner(model, **fx['kwargs'])
    pre.load_state_dict(fx['state_dict'])
    assert pre.steps == fx['state_dict']['steps'] == 3
    # round trip: what we write back has the reference's schema and the same factors
    sd = pre.state_dict()
    assert set(sd) == set(fx['state_dict'])
    assert set(sd['layers']) == set(fx['state_dict']['layers'])
    for name, st in fx['state_dict']['layers'].items():
        assert set(sd['layers'][name]) == {'A', 'G'}
        assert rel_fro(sd['layers'][name]['A'], st['A']) < 1e-7 and rel_fro(sd['layers'][name]['G'], st['G']) < 1e-7
    x, y = fx['batch']
    model.zero_grad()
    torch.nn.CrossEntropyLoss()(model(x.to(dev)), y.to(dev)).backward()
    pre.step()
    torch.cuda.synchronize()
    layers = {n: l for n, l in pre._layers.values()}
    for name, g in fx['after']['layers'].items():
        L = layers[name]
        assert rel_fro(L.a_factor, g['A']) < 2e-5 and rel_fro(L.g_factor, g['G']) < 2e-5, name
        assert rel_fro(L._p_view, g['P']) < 1e-3, (name, rel_fro(L._p_view, g['P']))
    assert abs(pre._compute_grad_scale() - fx['after']['scale']) <= 1e-3 * abs(fx['after']['scale'])
    for n, p in model.named_parameters():
        assert rel_fro(p.grad, fx['after']['final_grads'][n]) < 1e-3, n
    # and a state dict written here is accepted by the oracle-side schema check: tensors, fp32, square
    for name, st in sd['layers'].items():
        assert st['A'].dtype == torch.float32 and st['A'].shape[0] == st['A'].shape[1]


def test_factor_and_inv_dtype_are_accepted():
    """factor_dtype / inv_dtype (kfac/layers/base.py:350,364; eigen.py:319-320): fp64 is cast at the boundary
    (fp32 storage and arithmetic), bf16 rounds the hook inputs like the reference does."""
    from kfac_b200.preconditioner import KFACPreconditioner
    from workloads import TinyModel
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    base = TinyModel().to(dev)
    x = torch.rand(8, 10, device=dev)
    grads = {}
    for fd, idt in ((None, torch.float32), (torch.float64, torch.float64), (torch.bfloat16, torch.float32)):
        m = copy.deepcopy(base)
        p = KFACPreconditioner(m, factor_dtype=fd, inv_dtype=idt)
        m(x).sum().backward()
        p.step()
        torch.cuda.synchronize()
        layer = p._layers[m.linear1][1]
        assert layer.a_factor.dtype == (fd or torch.float32)
        assert layer.qa.dtype == idt
        grads[fd] = [q.grad.clone() for q in m.parameters()]
    for a, b in zip(grads[None], grads[torch.float64]):
        assert torch.equal(a, b)                  # fp64 request = fp32 arithmetic, identical bits
    for a, b in zip(grads[None], grads[torch.bfloat16]):
        assert rel_fro(a, b) < 5e-2               # bf16-rounded statistics
    with pytest.raises(ValueError):
        KFACPreconditioner(copy.deepcopy(base), inv_dtype=torch.float16)
