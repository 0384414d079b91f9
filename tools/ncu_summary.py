import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
h=rows[0]; u=rows[1]; v=rows[2]
keys=['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','TPC.TriageCompute.sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_lsu.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__inst_executed_op_shared_ld.sum', 'sm__inst_executed_pipe_cbu.sum','sm__inst_executed_pipe_adu.sum']
for i,n in enumerate(h):
    if n in keys or n.startswith('smsp__average_warps_issue_stalled') or ('inst_executed_pipe' in n and n.endswith('.sum')):
        print(f"{n:95s} {u[i]:16s} {v[i]}")
