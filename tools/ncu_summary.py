"""Summarise an .ncu-rep (ncu --set full) into a small text file for profiles/: key raw metrics + hottest SASS lines."""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__cluster_size', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
        'launch__shared_mem_per_block_dynamic', 'sm__inst_executed_pipe_tensor_subpipe_hmma.sum']


def main(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print(f'# {rep}')
    name_i = hdr.index('Kernel Name') if 'Kernel Name' in hdr else None
    if name_i is not None:
        print('kernel:', vals[name_i][:160])
    for k in KEYS:
        for h, u, v in zip(hdr, units, vals):
            if h == k:
                print(f'{k} = {v} {u}')
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    isrc, ismp = h.index('Source'), h.index('Warp Stall Sampling (All Samples)')
    data = []
    for r in rows[2:]:
        try:
            data.append((int(r[ismp]), r[isrc]))
        except Exception:
            pass
    tot = sum(d[0] for d in data) or 1
    print(f'warp-stall samples: {tot}; hottest SASS instructions (share of samples):')
    for s, t in sorted(data, reverse=True)[:14]:
        print(f'  {100 * s / tot:5.1f}%  {t.strip()[:110]}')
    mn = [t for _, t in data]
    for m in ('UTCHMMA', 'UTMALDG', 'LDTM', 'UTCBAR', 'SYNCS.PHASECHK'):
        print(f'  SASS {m}: {sum(1 for t in mn if m in t)} sites')


if __name__ == '__main__':
    main(sys.argv[1])
