import csv,sys,subprocess,collections
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__waves_per_multiprocessor','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for r in rows[2:]:
    print('=====',r[hdr.index('Kernel Name')][:60])
    for w in want:
        if w in hdr: print('  ',w,'=',r[hdr.index(w)],units[hdr.index(w)])
    items=[]
    for i,h in enumerate(hdr):
        if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued'):
            try: items.append((float(r[i]),h.replace('smsp__pcsamp_warps_issue_stalled_','')))
            except: pass
    tot=sum(v for v,_ in items)
    print('   stalls:',', '.join('%s %.0f%%'%(h,100*v/tot) for v,h in sorted(items,reverse=True)[:8]))
