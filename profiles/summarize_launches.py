"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.

    python profiles/summarize_launches.py gpurun_out/launches_r50.csv > profiles/r01_launches_r50.md
"""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v *= {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(unit, 1.0)
        k = row['Kernel Name']
        k = k.split('(')[0][-90:]
        agg[k][0] += 1
        agg[k][1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    print(f'launches captured: {n}; total device time {tot / 1e6:.2f} ms (cold-cache, serialised: compare SHARES)\n')
    print('| ms | share | launches | avg us | kernel |')
    print('|---:|---:|---:|---:|---|')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f'| {v[1] / 1e6:.2f} | {100 * v[1] / tot:.1f}% | {v[0]} | {v[1] / v[0] / 1e3:.1f} | `{k}` |')


if __name__ == '__main__':
    main(sys.argv[1])
