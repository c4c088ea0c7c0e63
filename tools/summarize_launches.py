"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import csv
import collections
import re
import sys


def main(path):
    rows = []
    with open(path, newline='') as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = r.get('Kernel Name', '?')
        name = re.sub(r'\(.*', '', name)
        name = re.sub(r'^void ', '', name)
        name = re.sub(r'\(anonymous namespace\)::', '', name)
        try:
            val = float(r['Metric Value'].replace(',', ''))
        except Exception:
            continue
        unit = r.get('Metric Unit', 'ns')
        scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(unit, 1e-3)
        rows.append((name, val * scale))
    tot = sum(v for _, v in rows)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v in rows:
        agg[n][0] += 1
        agg[n][1] += v
    print(f'total launches {len(rows)}  total device time {tot / 1e3:.3f} ms')
    print(f'{"kernel":70s} {"count":>6s} {"total us":>12s} {"avg us":>10s} {"share":>7s}')
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{n[:70]:70s} {c:6d} {v:12.1f} {v / c:10.1f} {100 * v / tot:6.1f}%')


if __name__ == '__main__':
    main(sys.argv[1])
