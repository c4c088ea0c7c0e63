"""Per-kernel durations of the decode step from an ncu launch list (csv from --metrics gpu__time_duration.sum)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
seq = [(r[ki].split('(')[0].split('::')[-1][:34], r[gi], float(r[vi].replace(',', ''))) for r in rows[1:]]
per_step = int(sys.argv[2]) if len(sys.argv) > 2 else 27
tot = sum(v for _, _, v in seq)
print(f'{len(seq)} launches, {tot / 1e3:.1f} us, {tot / 1e3 / (len(seq) / per_step):.1f} us of kernel time per token step')
a = collections.defaultdict(lambda: [0, 0])
for n, g, v in seq:
    a[(n, g)][0] += v
    a[(n, g)][1] += 1
for k, (v, c) in sorted(a.items(), key=lambda kv: -kv[1][0]):
    print(f'{k[0]:36s} grid {k[1]:14s} x{c:3d}  avg {v / c / 1e3:7.2f} us  {100 * v / tot:5.1f} %')
