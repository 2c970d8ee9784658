import csv, subprocess, sys, io
rep=sys.argv[1]
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys=['gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_uniform.sum','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed']
idx={k:hdr.index(k) for k in keys if k in hdr}
ki=hdr.index('Kernel Name')
tens=[h for h in hdr if 'tensor' in h and 'pct' in h][:6]
print('tensor metrics available:',tens)
for r in rows[2:]:
    print('==',r[ki][:100])
    for k,i in idx.items(): print(f"   {k:75s} {units[i]:10s} {r[i]}")
    for h in tens[:3]: print(f"   {h:75s} {units[hdr.index(h)]:10s} {r[hdr.index(h)]}")
