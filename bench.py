#!/usr/bin/env python
"""bench.py -- K-FAC hot path on ResNet-50 (ImageNet-shaped synthetic data).

    python bench.py --gpus N --steps K --warmup W            (own arm, B200)
    python bench.py --impl reference --gpus N --steps K ...   (CPU reference arm)

A "step" is one full training iteration of the data-parallel K-FAC hot path:
forward + backward with the factor hooks (K1-K4), `preconditioner.step()`
(factor EMA/all-reduce, eigendecompositions, precondition, kl-clip, write-back)
and the SGD update, with `factor_update_steps = inv_update_steps = 1`, i.e.
EVERY step runs every stage of the path (SURVEY.md 8(d) "stress 1/1").
Metric: images/s (whole job, all ranks).  One JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import warnings  # noqa: E402

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

warnings.filterwarnings('ignore', message='Full backward hook is firing')

METRIC = 'images/sec ResNet-50 K-FAC training step (fwd+bwd+factor hooks+preconditioner.step()+SGD), factor=inv=1'


def metric_name(model):
    return METRIC if model == 'resnet50' else METRIC.replace('ResNet-50', 'ResNet-32')


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--model', default='resnet50', choices=['resnet50', 'resnet32'])
    ap.add_argument('--batch', type=int, default=0, help='per-GPU batch (default 32 / 128)')
    ap.add_argument('--grad-worker-fraction', type=float, default=-1.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--budget-s', type=float, default=240.0, help='reference arm wall-clock budget')
    return ap.parse_args()


def make_workload(name, batch):
    from oracle.models import resnet32, resnet50
    if name == 'resnet50':
        b = batch or 32
        return resnet50, (b, 3, 224, 224), 1000, dict(damping=0.001, factor_decay=0.95, kl_clip=0.001, lr=0.1), \
            f'ResNet-50 ImageNet-shape synthetic, bs{b}/GPU, factor_update_steps=1, inv_update_steps=1, damping 1e-3'
    b = batch or 128
    return resnet32, (b, 3, 32, 32), 10, dict(damping=0.003, factor_decay=0.95, kl_clip=0.001, lr=0.1), \
        f'ResNet-32 CIFAR-shape synthetic, bs{b}/GPU, factor_update_steps=1, inv_update_steps=1, damping 3e-3'


# ------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8): 'hw_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4): 'sw_power_cap',
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.1)

    def stop(self):
        self._stop_evt.set()
        if self.ok:
            self.join(timeout=2)
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(s)}


# ------------------------------------------------------------------ phases
class PhaseTimer:
    """CUDA-event timing of the native phases inside step() (current stream)."""

    def __init__(self, pre):
        self.pre = pre
        self.events = {}
        self.enabled = False
        for name in ('_flush_factor_updates', '_reduce_factors', '_compute_inverses', '_precondition',
                     '_scale_and_write_back'):
            self._wrap(pre, name)
        for _, layer in pre._layers.values():
            self._wrap(layer.module, 'accumulate_a', 'factor_a')
            self._wrap(layer.module, 'accumulate_g', 'factor_g')

    def _wrap(self, obj, name, key=None):
        key = key or name.strip('_')
        fn = getattr(obj, name)

        def timed(*a, **k):
            if not self.enabled:
                return fn(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            self.events.setdefault(key, []).append((e0, e1))
            return out
        setattr(obj, name, timed)

    def totals_ms(self):
        return {k: sum(a.elapsed_time(b) for a, b in v) for k, v in self.events.items()}


def algorithmic_work(pre):
    """SURVEY.md 8(d) conventions: eigh 9n^3, precondition 4ga(g+a)."""
    eig = prec = 0.0
    for _, l in pre._layers.values():
        a, g = l.a_dim, l.g_dim
        eig += 9.0 * (a ** 3 + g ** 3)
        prec += 4.0 * g * a * (g + a)
    return eig, prec


def tc_gemm_microbench(lib, dev, n=4096, iters=10):
    """Isolated timing of the tcgen05 GEMM engine (the kernel under K3/K5/K8):
    D = A B^T, n^3 problem, fp32 operands (3 x 64 MB > 126 MB L2), CUDA events."""
    A = torch.randn(n, n, device=dev)
    B = torch.randn(n, n, device=dev)
    D = torch.empty(n, n, device=dev)
    s = torch.cuda.current_stream().cuda_stream

    def run():
        rc = lib.kfac_gemm_tn_tc(A.data_ptr(), n, B.data_ptr(), n, D.data_ptr(), n, n, n, n, 1.0, 0, 1, s)
        assert rc == 0, lib.kfac_last_error()
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return 2.0 * n ** 3 / (ms / 1e3) / 1e12, ms


# ------------------------------------------------------------------ own arm
def run_b200(args):
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py (own arm) needs a CUDA device'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from kfac_b200 import _cabi
    from kfac_b200.preconditioner import KFACPreconditioner
    lib = _cabi.load()

    make_model, shape, classes, hp, wl = make_workload(args.model, args.batch)
    torch.manual_seed(0)
    model = make_model().to(dev)
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
    frac = args.grad_worker_fraction
    if frac < 0:
        frac = 1.0 if world == 1 else 0.5
    pre = KFACPreconditioner(model, factor_update_steps=1, inv_update_steps=1, grad_worker_fraction=frac, **hp)
    opt = torch.optim.SGD(model.parameters(), lr=hp['lr'], momentum=0.9)
    crit = torch.nn.CrossEntropyLoss()
    timer = PhaseTimer(pre)

    torch.manual_seed(1 + rank)
    B = shape[0]
    host_x = torch.randn(*shape).pin_memory()
    host_y = torch.randint(0, classes, (B,)).pin_memory()
    dev_x, dev_y = host_x.to(dev), host_y.to(dev)
    loss_host = torch.zeros(1).pin_memory()

    def one_step(e2e):
        if e2e:
            x = host_x.to(dev, non_blocking=True)
            y = host_y.to(dev, non_blocking=True)
        else:
            x, y = dev_x, dev_y
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x), y)
        loss.backward()
        pre.step()
        opt.step()
        if e2e:
            loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(n, e2e):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.kfac_launch_count()
        e0.record()
        for _ in range(n):
            one_step(e2e)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), lib.kfac_launch_count() - l0

    for _ in range(max(args.warmup, 3)):
        one_step(False)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    timer.enabled = True
    ms_dev, launches = timed_loop(args.steps, False)
    timer.enabled = False
    phase_ms = timer.totals_ms()
    ms_e2e, _ = timed_loop(args.steps, True)
    clocks = sampler.stop()

    K = args.steps
    imgs = B * world * K
    value = imgs / (ms_dev / 1e3)
    e2e_value = imgs / (ms_e2e / 1e3)
    eig_flops, prec_flops = algorithmic_work(pre)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:  # noqa: BLE001
        pass
    tensor_peak = float(peaks.get('bf16_tflops_sustained', 1400.0))
    peak_src = 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)' if peaks else 'fallback'
    # dominant phase: the batched eigensolver; at N>1 each rank solves only its
    # KAISA share, the roofline line is reported for rank 0's local phase time.
    inv_ms = phase_ms.get('compute_inverses', 0.0) / K
    my_eig = 0.0
    r = rank
    for name, l in pre._layers.values():
        if pre._assignment.inv_worker(name, 'A') == r:
            my_eig += 9.0 * l.a_dim ** 3
        if pre._assignment.inv_worker(name, 'G') == r:
            my_eig += 9.0 * l.g_dim ** 3
    achieved = (my_eig / (inv_ms / 1e3) / 1e12) if inv_ms > 0 else 0.0
    gemm_tf, gemm_ms = tc_gemm_microbench(lib, dev) if rank == 0 else (0.0, 0.0)
    burst_peak = float(peaks.get('bf16_tflops', 1590.0))
    out = {
        'metric': metric_name(args.model), 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': K,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms_dev / K, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 (model fwd/bwd fp32; K-FAC path fp32)',
        'data': 'synthetic',
        'config': {'workload': wl, 'global_batch': B * world, 'grad_worker_fraction': frac,
                   'parallelism': f'dp{world}', 'l2': 'inputs+factors >> 126 MB L2 (activations 1.4 GB, factors 615 MB)'},
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': 'images/s', 'ms_per_step': ms_e2e / K,
                'h2d_bytes_per_step': host_x.numel() * 4 + host_y.numel() * 8, 'd2h_bytes_per_step': 4},
        'gpu_launches': int(launches),
        'kfac_phase_ms_per_step': {k: v / K for k, v in sorted(phase_ms.items())},
        # preconditioner.step() itself (BASELINE.json: "K-FAC step() ms"): the phases step() runs, without the
        # factor hooks that run inside forward/backward
        'kfac_step_ms': sum(v for k, v in phase_ms.items() if k not in ('factor_a', 'factor_g')) / K,
        'kfac_hooks_ms': sum(v for k, v in phase_ms.items() if k in ('factor_a', 'factor_g')) / K,
        # dominant launch of the step: ONE kfac_eigh_batched call (all factors this rank owns), timed
        # live in the timed region with CUDA events on the launching stream
        'roofline': {'kernel': 'kfac_eigh_batched (block one-sided Jacobi rounds: tcgen05 Gram -> smem Jacobi '
                               '-> tcgen05 apply), one call per step over all factors of this rank',
                     'bound': 'tensor', 'achieved': achieved, 'peak': tensor_peak, 'unit': 'TFLOP/s',
                     'frac': achieved / tensor_peak if tensor_peak else None, 'traffic': None,
                     'ms_per_launch': inv_ms,
                     'convention': 'algorithmic 9 n^3 flop per eigendecomposition (SURVEY.md 8d) summed over the '
                                   "rank's factors / CUDA-event time of the call; the Jacobi rounds execute ~12 n^3 "
                                   'flop per sweep as 3xTF32 MMAs and are latency/HBM bound (profiles/r01_ncu_full_pipeline.md: '
                                   'Gram launch 88.9 MB DRAM traffic, tensor pipe 53 % active)',
                     'peak_source': peak_src},
        # the GEMM engine all tensor-core kernels instantiate, timed alone on a 4096^3 problem
        'roofline_engine': {'kernel': 'tc::pipeline_kernel<GemmPolicy> (tcgen05 3xTF32 GEMM engine), 4096^3, isolated',
                            'bound': 'tensor', 'achieved': gemm_tf, 'peak': burst_peak, 'unit': 'TFLOP/s',
                            'frac': gemm_tf / burst_peak if burst_peak else None, 'traffic': None,
                            'ms_per_launch': gemm_ms,
                            'convention': 'algorithmic fp32 flops 2*M*N*K per launch / CUDA-event time; every product is '
                                          'issued as 3 TF32 MMAs, so the tensor pipe executes 3x these flops at the TF32 '
                                          'rate (= half the bf16 rate the peak was measured with)',
                            'peak_source': 'measured (MEASURED_PEAKS.json bf16_tflops, burst)' if peaks else 'fallback 1590'},
        'algorithmic_flops_per_step': {'eigh_9n3': eig_flops, 'precondition_4ga(g+a)': prec_flops},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out['cpu_baseline'] = cpu_baseline(args, steps=1, warmup=0, budget=120.0)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------ CPU arm
def usable_cores():
    """Host threads this process may really use: CPU affinity capped by the cgroup quota (a
    container that sees 128 logical CPUs but owns a fraction of them would otherwise
    oversubscribe MKL/OpenMP catastrophically) and by 32 (LAPACK eigh does not scale further)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        n = os.cpu_count() or 1
    try:
        q, p = open('/sys/fs/cgroup/cpu.max').read().split()[:2]
        if q != 'max':
            n = min(n, max(1, int(float(q) / float(p))))
    except Exception:  # noqa: BLE001
        try:
            q = int(open('/sys/fs/cgroup/cpu/cpu.cfs_quota_us').read())
            p = int(open('/sys/fs/cgroup/cpu/cpu.cfs_period_us').read())
            if q > 0:
                n = min(n, max(1, q // p))
        except Exception:  # noqa: BLE001
            pass
    env = os.environ.get('KFAC_BENCH_CPU_THREADS')
    if env:
        return max(1, int(env))
    return max(1, min(n, 32))


def cpu_reference_loop(args, steps, warmup, budget):
    """The oracle port (same torch CPU ops as the reference) on the host cores."""
    from oracle.kfac_oracle import OraclePreconditioner
    make_model, shape, classes, hp, wl = make_workload(args.model, args.batch)
    cores = usable_cores()
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = make_model()
    pre = OraclePreconditioner(model, factor_update_steps=1, inv_update_steps=1, **hp)
    opt = torch.optim.SGD(model.parameters(), lr=hp['lr'], momentum=0.9)
    crit = torch.nn.CrossEntropyLoss()
    x = torch.randn(*shape)
    y = torch.randint(0, classes, (shape[0],))

    def one():
        opt.zero_grad(set_to_none=True)
        crit(model(x), y).backward()
        t0 = time.perf_counter()
        pre.step()
        t1 = time.perf_counter()
        opt.step()
        return t1 - t0
    # every step is checked against the wall-clock budget; if not even the warm-up fits, the
    # last completed step is the sample
    t_start = time.perf_counter()
    walls, kfacs = [], []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        k = one()
        walls.append(time.perf_counter() - t0)
        kfacs.append(k)
        if time.perf_counter() - t_start + walls[-1] > budget:
            break
    if len(walls) > warmup:
        walls, kfacs = walls[warmup:], kfacs[warmup:]
    else:
        walls, kfacs = walls[-1:], kfacs[-1:]
    done, total, step_s = len(walls), sum(walls), sum(kfacs)
    return {'images_per_s': shape[0] * done / total, 'ms_per_step': total / done * 1e3,
            'kfac_step_ms': step_s / done * 1e3, 'steps_done': done, 'cores': cores, 'workload': wl, 'batch': int(shape[0])}


def cpu_baseline(args, steps, warmup, budget):
    """Runs the CPU arm in a child process under a hard wall-clock limit so that the default
    bench run always ends within minutes, whatever the host does."""
    import subprocess
    cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--model', args.model,
           '--steps', str(steps), '--warmup', str(warmup), '--budget-s', str(budget)]
    if args.batch:
        cmd += ['--batch', str(args.batch)]
    env = dict(os.environ)
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK', 'OMP_NUM_THREADS'):
        env.pop(k, None)
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=budget + 90, env=env)
        line = [l for l in out.stdout.splitlines() if l.startswith('{')][-1]
        d = json.loads(line)
        cb = d['cpu_baseline']
        cb['kfac_step_ms'] = d.get('kfac_step_ms')
        cb['ms_per_step'] = d.get('ms_per_step')
        return cb
    except Exception as e:  # noqa: BLE001
        return {'value': None, 'unit': 'images/s', 'cores': usable_cores(), 'kind': 'port',
                'sample': f'CPU oracle did not finish one step within {budget + 90:.0f} s ({type(e).__name__})'}


def _cpu_baseline_inline(args, steps, warmup, budget):
    r = cpu_reference_loop(args, steps, warmup, budget)
    return {'value': r['images_per_s'], 'unit': 'images/s', 'cores': r['cores'], 'kind': 'port',
            'sample': f"{r['steps_done']} full step(s) of the same workload, no warm-up "
                      f"(oracle port = the reference's torch CPU ops); kfac step() {r['kfac_step_ms']:.0f} ms",
            'kfac_step_ms': r['kfac_step_ms'], 'ms_per_step': r['ms_per_step']}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    r = cpu_reference_loop(args, args.steps, min(args.warmup, 1), args.budget_s)
    out = {
        'impl': 'reference', 'metric': metric_name(args.model), 'value': r['images_per_s'], 'unit': 'images/s',
        'n_gpus': args.gpus, 'steps': r['steps_done'], 'warmup': min(args.warmup, 1),
        'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': r['workload'], 'global_batch': r['batch'], 'parallelism': 'cpu, world_size 1'},
        'cpu_baseline': {'value': r['images_per_s'], 'unit': 'images/s', 'cores': r['cores'], 'kind': 'port',
                         'sample': f"{r['steps_done']} of {args.steps} requested steps within a "
                                   f"{args.budget_s:.0f} s budget; kfac step() {r['kfac_step_ms']:.0f} ms"},
        'e2e': {'value': r['images_per_s'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'kfac_step_ms': r['kfac_step_ms'],
    }
    print(json.dumps(out))


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_b200(a)
