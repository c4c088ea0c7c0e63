cd /root/repo
k=vq_idx_dz_kernel
rep=gpurun_out/r02c_$k
timeout 300 ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -f -o $rep python tools/idx_debug.py 4096 15 8 8 512 3 > gpurun_out/ncu_$k.log 2>&1
{ python tools/ncu_stalls.py $rep.ncu-rep; python tools/ncu_hot.py $rep.ncu-rep 40; } > gpurun_out/r02c_ncu_$k.txt 2>&1
ncu -i $rep.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,v=rows[0],rows[1],rows[2]
for a,b,c in zip(h,u,v):
    if any(k in a for k in ('l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','lts__t_sectors_op_read.sum','l1tex__t_sector_hit_rate','smsp__warps_eligible.avg.per_cycle_active','smsp__issue_active.avg.per_cycle_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts.sum','smsp__average_warp_latency','lts__t_sectors_srcunit_tex_op_read.sum','l1tex__lsu_writeback_active','l1tex__throughput','lsu_mem_global_op_st.sum')): print(a,'=',c,b)
" >> gpurun_out/r02c_ncu_$k.txt
