"""Count the SASS instructions that prove the hardware path, per kernel of the shipped library.

    python profiles/sass_evidence.py kfac-pytorch_b200/csrc/libkfac_b200.so > profiles/r02_sass_evidence.md
"""
import collections
import re
import subprocess
import sys

KEYS = ['UTCHMMA', 'UTMALDG', 'LDTM', 'UTCBAR', 'SYNCS', 'FFMA2', 'FMUL2', 'REDG', 'ATOM', 'MEMBAR']


def main(so):
    sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
    cnt = collections.defaultdict(collections.Counter)
    f = None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            f = m.group(1)
            continue
        for k in KEYS:
            if k in line:
                cnt[f][k] += 1
    print('# SASS evidence (round 2, shipped `libkfac_b200.so`, `cuobjdump -sass`)\n')
    print('Counts of the instructions that prove the hardware path, per kernel (kernels without any of them omitted).')
    print('`UTCHMMA` = tcgen05.mma, `UTMALDG` = TMA tensor load, `LDTM` = tcgen05.ld (TMEM -> registers), `UTCBAR` = '
          'tcgen05.commit, `SYNCS` = mbarrier, `FFMA2`/`FMUL2` = packed fp32 pairs (sm_100), `REDG`/`ATOM` = global '
          'reductions / atomics.\n')
    print('| kernel | ' + ' | '.join(KEYS) + ' |')
    print('|---|' + '---:|' * len(KEYS))
    for f, c in cnt.items():
        name = subprocess.run(['c++filt', f], capture_output=True, text=True).stdout.strip()
        name = name.replace('(anonymous namespace)::', '').replace('kfac::', '')
        name = re.sub(r'\((?!anonymous).*', '', name)[:70]
        print(f'| `{name}` | ' + ' | '.join(str(c.get(k, 0)) for k in KEYS) + ' |')
    print('\n`tc::pipeline_kernel<GemmPolicy>` (factor SYRK, single GEMMs) and `tc::pipeline_kernel<GroupedPolicy>` '
          '(precondition stages, D&C merges, back-transformation) are the two tcgen05 kernels: 12 UTCHMMA each = 3 TF32 '
          'split terms x 4 MMAs (K = 8) per 32-wide k-block. `sytrd_kernel` is SIMT by design (DESIGN.md 3): FFMA2 in its '
          'tile products and rank-64 update, barriers as `red.release.gpu` (REDG) + `ld.acquire.gpu` polling.')


if __name__ == '__main__':
    main(sys.argv[1])
