"""Top-N hottest SASS instructions of an .ncu-rep (stall samples), with a few neighbours for context.  python tools/ncu_hot.py rep [N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
isrc, ismp = h.index('Source'), h.index('Warp Stall Sampling (All Samples)')
iex = h.index('Instructions Executed')
data = []
for i, r in enumerate(rows[2:]):
    try:
        data.append((int(r[ismp]), int(r[iex]), r[isrc].strip()))
    except Exception:
        data.append((0, 0, r[isrc].strip() if len(r) > isrc else ''))
tot = sum(d[0] for d in data) or 1
order = sorted(range(len(data)), key=lambda i: -data[i][0])[:N]
print(f'total samples {tot}; instructions {len(data)}')
for i in sorted(order):
    s, e, t = data[i]
    prev = data[i - 1][2][:60] if i > 0 else ''
    print(f'{i:5d} {100 * s / tot:5.1f}%  exec {e:8d}  {t[:90]}   <- {prev}')
