"""Summarise `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_tc --csv` of one bench
pass into the per-launch DRAM traffic record bench.py quotes as roofline.traffic.  python tools/gemm_traffic.py in.csv out.json"""
import collections
import csv
import json
import sys


def main(src, dst):
    with open(src, newline='') as f:
        lines = [l for l in f if not l.startswith('==')]
    per = collections.defaultdict(dict)
    for r in csv.DictReader(lines):
        try:
            v = float(r['Metric Value'].replace(',', ''))
        except Exception:
            continue
        u = r.get('Metric Unit', '')
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(u, 1)
        per[r['ID']][r['Metric Name']] = v * scale
    n = len(per)
    rd = sum(d.get('dram__bytes_read.sum', 0) for d in per.values())
    wr = sum(d.get('dram__bytes_write.sum', 0) for d in per.values())
    us = sum(d.get('gpu__time_duration.sum', 0) for d in per.values())
    json.dump(dict(launches=n, dram_read_bytes_per_launch=rd / max(n, 1), dram_write_bytes_per_launch=wr / max(n, 1),
                   ncu_us_per_launch=us / max(n, 1),
                   note='ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the tcgen05 GEMM launches of one training pass '
                        '(bench.py --batch 256: the same shape mix as the full step, which is 16 such passes)'), open(dst, 'w'), indent=1)
    print(open(dst).read())


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
