import csv,sys
rows=list(csv.reader(sys.stdin)); hdr,units,vals=rows[0],rows[1],rows[2]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__lsu_writeback_active.sum.pct_of_peak_sustained_elapsed','lts__t_sectors_srcunit_tex_op_write.sum','lts__t_sectors_srcunit_tex_op_read.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct']
for k in keys:
    if k in hdr: i=hdr.index(k); print(f'{k:80s} {vals[i]:>16s} {units[i]}')
for h,u,v in zip(hdr,units,vals):
    if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h and float(v)>0.1: print(f'{h[34:]:60s} {float(v):.2f}')
