import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','lts__t_sectors_op_read.sum','lts__t_sectors_op_write.sum','lts__t_sector_hit_rate.pct','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__waves_per_multiprocessor','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum','smsp__inst_executed.sum']
stall=[h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    print('-----')
    for w in want:
        for i,h in enumerate(hdr):
            if h==w: print(w, '=', r[i], units[i])
    st=sorted(((float(r[hdr.index(h)]),h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')) for h in stall), reverse=True)[:6]
    print('top stalls:', ', '.join(f"{n} {v:.2f}" for v,n in st))
