cd /root/repo
for k in nce_idx_fwd_kernel vq_idx_bwd_kernel nce_idx_bwd_tc_kernel; do
  rep=gpurun_out/r02b_$k
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -f -o $rep python tools/idx_debug.py 4096 15 8 8 512 3 > gpurun_out/ncu_$k.log 2>&1
  { python tools/ncu_summary.py $rep.ncu-rep; python tools/ncu_stalls.py $rep.ncu-rep; python tools/ncu_hot.py $rep.ncu-rep 45; } > gpurun_out/r02b_ncu_$k.txt 2>&1
done
