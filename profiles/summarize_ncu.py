import csv,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit','smsp__inst_executed.sum','smsp__issue_active.avg.pct','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__average_warps_issue_stalled','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','launch__grid_size','launch__shared_mem_per_block_dynamic','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__inst_executed_op_shared','lts__t_bytes.sum ','l1tex__m_xbar2l1tex_read_bytes.sum ']
for h,u,v in zip(hdr,units,vals):
    if any(h.startswith(k) or h==k.strip() for k in keys):
        if 'stalled' in h and float(v or 0)<0.15: continue
        print(f"{h:85s} {u:12s} {v}")
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; data=rows[2:]
ia=hdr.index('Source'); ie=hdr.index('Instructions Executed'); it=hdr.index('Avg. Threads Executed'); isamp=hdr.index('# Samples')
tot=sum(int(r[ie]) for r in data); tots=sum(int(r[isamp]) for r in data)
print('instr',tot,'samples',tots)
n=int(sys.argv[2]) if len(sys.argv)>2 else 25
for r in sorted(data,key=lambda r:-int(r[isamp]))[:n]: print(f"{int(r[isamp])/tots*100:5.1f}% ex={r[ie]:>11s} thr={r[it]:>3s} {r[ia][:90]}")
