#!/usr/bin/env bash
# ncu --set full of one GEMM shape of tools/gemm_bench.py (case index $1, default 0 = fwd qkv), summarised into gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
c=${1:-0}
rep=gpurun_out/r02_gemm_case$c
timeout 300 ncu --set full --import-source on --clock-control none -k regex:gemm_tc -s 3 -c 1 -f -o $rep python tools/gemm_bench.py --cfgs=-1 --only $c --iters 2 > gpurun_out/ncu_gemm_case$c.log 2>&1
{ python tools/ncu_summary.py $rep.ncu-rep; python tools/ncu_stalls.py $rep.ncu-rep; python tools/ncu_hot.py $rep.ncu-rep 30; } > gpurun_out/r02_ncu_gemm_case$c.txt 2>&1
grep -E "kernel:|time_duration|tensor_cycles_active.avg.pct_of_peak_sustained_active|stall samples|warps_active|dram__bytes|registers" gpurun_out/r02_ncu_gemm_case$c.txt
