"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.

    python profiles/summarize_launches.py gpurun_out/launches_r50.csv > profiles/r01_launches_r50.md
    python profiles/summarize_launches.py launches.csv --period sytrd_kernel 2   # only the launches from the 2nd to the
                                                                                 # 3rd occurrence of that kernel (one step)
"""
import collections
import csv
import sys


def main(path, period=None, which=1):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    rows = [r for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
    if period:
        marks = [i for i, r in enumerate(rows) if period in r['Kernel Name']]
        rows = rows[marks[which - 1]:marks[which]]
        print(f'window: launches from occurrence {which} to {which + 1} of `{period}` (one bench step)\n')
    for row in rows:
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
    if len(sys.argv) > 3 and sys.argv[2] == '--period':
        main(sys.argv[1], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 1)
    else:
        main(sys.argv[1])
