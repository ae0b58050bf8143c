import csv, sys
want = ["gpu__time_duration.sum","launch__grid_size","launch__block_size","launch__registers_per_thread","launch__shared_mem_per_block_dynamic",
"dram__bytes_read.sum","dram__bytes_read.sum.per_second","dram__bytes_write.sum","lts__t_sector_hit_rate.pct","lts__t_sectors.sum","lts__throughput.avg.pct_of_peak_sustained_elapsed",
"l1tex__t_sector_hit_rate.pct","l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
"l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum","l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum","l1tex__throughput.avg.pct_of_peak_sustained_active","sm__cycles_elapsed.max","smsp__cycles_active.avg",
"smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__thread_inst_executed_per_inst_executed.ratio",
"smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","sm__throughput.avg.pct_of_peak_sustained_elapsed",
"l1tex__m_xbar2l1tex_read_bytes.sum","l1tex__m_xbar2l1tex_read_bytes.sum.per_second","lts__t_bytes.sum.per_second","l1tex__m_xbar2l1tex_read_sectors.sum","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]; vals=rows[2]
print("Kernel", vals[hdr.index("Kernel Name")])
for w in want:
    if w in hdr:
        i=hdr.index(w); print(f"{w:90s} {vals[i]:>20s} {units[i]}")
