#!/usr/bin/env bash
# per-kernel device time / instruction counts of one pass over the index-path kernels (ncu launch list; serialised, cold cache:
# compare shares, never quote as bench values)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__cycles_elapsed.max,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum \
    --clock-control none --csv --log-file gpurun_out/r02_idx_launches.csv -k regex:"idx|vq_tc|reduce_partials" \
    python tools/idx_debug.py 4096 15 8 8 512 3 > gpurun_out/r02_idx_launches.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r02_idx_launches.csv')) if len(r) > 5]
h = rows[0]
ik, im, iv = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value')
for r in rows[1:]:
    print(r[ik][:58].ljust(58), r[im][:48].ljust(48), r[iv])
PY
