import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines())); hdr,units,vals=rows[0],rows[1],rows[2]
d={h:(u,v) for h,u,v in zip(hdr,units,vals)}
keys=['gpu__time_duration.sum','dram__bytes_read.sum','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem',
'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active']
for k in keys:
    if k in d: print(k, d[k])
for h in hdr:
    if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h: print(h.replace('smsp__average_warps_issue_stalled_','stall_').replace('_per_issue_active.ratio',''), d[h][1])
for h in hdr:
    if ('pipe' in h and 'pct_of_peak_sustained_active' in h and 'inst_executed' in h): print(h, d[h][1])
