#!/usr/bin/env bash
# ncu launch list (per-kernel device time, serialized; compare SHARES) of one small bench invocation.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B=${1:-128}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b${B}.csv \
   python bench.py --batch $B --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_b${B}.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_b${B}.csv > gpurun_out/launches_b${B}_summary.txt 2>&1
tail -n 40 gpurun_out/launches_b${B}_summary.txt
