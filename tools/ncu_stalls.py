"""Stall-reason and opcode breakdown of one kernel in an .ncu-rep (ncu --set full --import-source on)."""
import collections
import csv
import io
import subprocess
import sys


def main(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, v = rows[0], rows[1], rows[2]
    st = []
    for a, b, c in zip(h, u, v):
        if 'pcsamp_warps_issue_stalled' in a and 'not_issued' not in a:
            st.append((float(c.replace(',', '')), a.replace('smsp__pcsamp_warps_issue_stalled_', '')))
        if a in ('gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'launch__registers_per_thread',
                 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
                 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
                 'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
                 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__inst_executed_pipe_tensor.sum', 'lts__t_sector_hit_rate.pct'):
            print(f'{a} = {c} {b}')
    tot = sum(x[0] for x in st) or 1
    print('stall samples:', ', '.join(f'{n} {100 * x / tot:.1f}%' for x, n in sorted(st, reverse=True)[:9]))
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    isrc, iex = h.index('Source'), h.index('Instructions Executed')
    c = collections.Counter()
    tot = 0
    for r in rows[2:]:
        try:
            n = int(r[iex])
        except Exception:
            continue
        t = r[isrc].strip().split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        c[op] += n
        tot += n
    print('opcode mix:', ', '.join(f'{k} {100 * x / tot:.1f}%' for k, x in c.most_common(14)), f'(total {tot})')


if __name__ == '__main__':
    main(sys.argv[1])
