import csv, subprocess, sys, io
rep=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 60
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
hi=[i for i,r in enumerate(rows) if len(r)>5 and r[0]=='Line No' and '# Samples' in r][0]
h=rows[hi]; ci=h.index('# Samples'); ii=h.index('Instructions Executed')
agg={}
for r in rows[hi+1:]:
    if len(r)<=ii: continue
    if r[0]=='Line No': break
    if r[0].strip().isdigit() and r[2] in ('','-'):
        try: agg[int(r[0])]=(float(r[ci] or 0), float(r[ii] or 0), r[1])
        except: pass
tot_s=sum(v[0] for v in agg.values()); tot_i=sum(v[1] for v in agg.values())
print('total warp-instructions %.4g'%tot_i)
cum=0
for ln,(s,ie,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:top]:
    cum+=ie
    print(f"{100*ie/tot_i:5.1f}% inst (cum {100*cum/tot_i:5.1f}%) {100*s/tot_s:5.1f}% smp  L{ln:4d}: {t.strip()[:100]}")
