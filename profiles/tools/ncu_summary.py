import csv, subprocess, sys, io, collections
rep = sys.argv[1]
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys=['Kernel Name','gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__shared_mem_per_block_dynamic','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','dram__bytes_read.sum','dram__bytes_write.sum','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed']
d=dict(zip(hdr,zip(units,vals)))
for k in keys:
    if k in d: print(f"{k:85s} {d[k][0]:10s} {d[k][1]}")
print('--- stall reasons (warps per issue)')
for h in hdr:
    if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
        v=float(d[h][1]); 
        if v>0.05: print(f"   {h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):30s} {v:.2f}")
# source page
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
# find header row
hi=None
for i,r in enumerate(rows):
    if 'Source' in r and any('Samples' in c for c in r): hi=i; break
if hi is not None:
    h=rows[hi]; si=h.index('Source'); 
    sc=[i for i,c in enumerate(h) if c.strip()=='# Samples' or c.strip()=='Warp Stall Sampling (All Samples)' or 'Sampling (All' in c]
    ic=[i for i,c in enumerate(h) if c.strip()=='Instructions Executed']
    col=sc[0] if sc else None
    print('columns:', [c for c in h][:12])
    if col is not None:
        data=[]
        for n,r in enumerate(rows[hi+1:]):
            try: data.append((float(r[col] or 0), n+1, r[si].strip(), float(r[ic[0]] or 0) if ic else 0))
            except: pass
        tot=sum(x[0] for x in data)
        print('total samples',tot)
        for s,n,t,ie in sorted(data,reverse=True)[:40]:
            print(f"{100*s/tot:5.1f}%  inst {ie:12.0f}  L{n:4d}: {t[:110]}")
