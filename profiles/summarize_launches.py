"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of ONE steady-state update step
(the launches between the last two prep_kernel launches).  ncu times are cold-cache and serialised: compare SHARES."""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = []
    for x in csv.DictReader(lines):
        if x.get('Metric Name') == 'gpu__time_duration.sum':
            rows.append((int(x['ID']), x['Kernel Name'], float(x['Metric Value'].replace(',', '')), x.get('Grid Size'), x.get('Block Size')))
    return rows


def main(path, verbose=False):
    rows = load(path)
    idx = [i for i, r in enumerate(rows) if 'prep_kernel' in r[1]]
    step = rows[idx[-2]:idx[-1]]
    agg = collections.OrderedDict()
    for _, n, t, g, b in step:
        k = re.sub(r'\(.*', '', n).replace('<unnamed>::', '').replace('void ', '')
        d = agg.setdefault(k, [0, 0.0])
        d[0] += 1
        d[1] += t
    tot = sum(v[1] for v in agg.values())
    print(f'{path}: {len(step)} launches in one step, sum of kernel durations {tot / 1e3:.1f} us (serialised, cold L2)')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{v[1] / 1e3:9.1f} us {100 * v[1] / tot:5.1f}%  x{v[0]:3d}  avg {v[1] / v[0] / 1e3:7.2f} us  {k[:80]}')
    if verbose:
        for _, n, t, g, b in step:
            print(f'{t / 1e3:8.1f}', g, b, re.sub(r'\(.*', '', n).replace('<unnamed>::', '')[:70])


if __name__ == '__main__':
    main(sys.argv[1], verbose=len(sys.argv) > 2)
