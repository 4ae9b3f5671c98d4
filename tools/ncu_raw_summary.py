import csv,sys,subprocess
out=subprocess.run(['ncu','-i',sys.argv[1],'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','smsp__inst_executed.sum','smsp__cycles_active.avg','sm__cycles_elapsed.max','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','smsp__warps_eligible.avg.per_cycle_active','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sectors_srcunit_tex_op_write.sum','l1tex__m_xbar2l1tex_read_bytes.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('=====', r[hdr.index('Kernel Name')][:70])
    for w in want:
        if w in hdr: print('  %-70s %s %s'%(w, r[hdr.index(w)], units[hdr.index(w)]))
    items=[]
    for i,h in enumerate(hdr):
        if 'average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio'):
            try: items.append((float(r[i]),h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
            except: pass
    print('  stalls:', ', '.join('%s=%.2f'%(h,v) for v,h in sorted(items,reverse=True)[:8]))
