#!/usr/bin/env bash
# ncu --set full capture of each kernel of the index path at the BASELINE size (one launch each, via tools/idx_debug.py),
# summarised with tools/ncu_summary.py + tools/ncu_stalls.py into gpurun_out/r02_ncu_<kernel>.txt
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KERNELS=${*:-"vq_tc_fwd_kernel nce_idx_fwd_kernel nce_idx_bwd_tc_kernel reduce_partials_kernel vq_idx_dz_kernel vq_idx_de_kernel"}
for k in $KERNELS; do
  rep=gpurun_out/r02_$k
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -f -o $rep \
      python tools/idx_debug.py 4096 15 8 8 512 3 > gpurun_out/ncu_$k.log 2>&1
  { python tools/ncu_summary.py $rep.ncu-rep; python tools/ncu_stalls.py $rep.ncu-rep; } > gpurun_out/r02_ncu_$k.txt 2>&1
  echo "== $k"; grep -E "gpu__time_duration|stall samples|warps_active|issue_active|dram__bytes|registers|bank_conflicts" gpurun_out/r02_ncu_$k.txt
done
