"""DRAM traffic of one kfac_eigh_batched call from an ncu CSV
(`--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`) of tests/eigh_batch_probe.py:
sums the launches of this library's kernels (kfac::*, tc::pipeline_kernel) and writes profiles/eigh_traffic.json.

    python profiles/summarize_traffic.py gpurun_out/r2_traffic.csv resnet50 > profiles/r02_eigh_traffic.md
"""
import collections
import csv
import json
import os
import sys

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0}


def main(path, model):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    per = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])   # launches, read, write, seconds
    ids = collections.defaultdict(set)
    for row in csv.DictReader(lines):
        k = row['Kernel Name']
        if not (k.startswith('kfac::') or 'kfac::' in k.split('(')[0] or 'pipeline_kernel' in k):
            continue
        name = k.split('(')[0][-80:]
        v = float(row['Metric Value'].replace(',', '')) * UNIT.get(row['Metric Unit'], 1.0)
        m = row['Metric Name']
        ids[name].add(row['ID'])
        if m == 'dram__bytes_read.sum':
            per[name][1] += v
        elif m == 'dram__bytes_write.sum':
            per[name][2] += v
        elif m == 'gpu__time_duration.sum':
            per[name][3] += v
    for name in per:
        per[name][0] = len(ids[name])
    rd = sum(v[1] for v in per.values())
    wr = sum(v[2] for v in per.values())
    t = sum(v[3] for v in per.values())
    print(f'one kfac_eigh_batched call ({model} factor dimensions): DRAM read {rd / 1e9:.2f} GB, write {wr / 1e9:.2f} GB, '
          f'{sum(v[0] for v in per.values())} launches, {t * 1e3:.1f} ms of serialised device time\n')
    print('| kernel | launches | read GB | write GB | ms |')
    print('|---|---:|---:|---:|---:|')
    for name, v in sorted(per.items(), key=lambda kv: -(kv[1][1] + kv[1][2])):
        print(f'| `{name}` | {v[0]} | {v[1] / 1e9:.3f} | {v[2] / 1e9:.3f} | {v[3] * 1e3:.2f} |')
    out = {'model': model, 'dram_bytes_per_call': rd + wr, 'dram_read_bytes': rd, 'dram_write_bytes': wr,
           'source': os.path.basename(path), 'how': 'ncu dram__bytes_read.sum + dram__bytes_write.sum summed over every kernel of one '
           'kfac_eigh_batched call on the ResNet-50 factor dimensions (tests/eigh_batch_probe.py)'}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'eigh_traffic.json'), 'w') as f:
        json.dump(out, f, indent=1)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else 'resnet50')
