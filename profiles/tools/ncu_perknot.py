"""per-source-line warp-instructions per LARS knot from an ncu report: python scratch/ncu_perknot.py rep knots [min]"""
import csv, subprocess, sys, io
rep=sys.argv[1]; knots=float(sys.argv[2]); mn=float(sys.argv[3]) if len(sys.argv)>3 else 3.0
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
hi=[i for i,r in enumerate(rows) if len(r)>5 and r[0]=='Line No' and '# Samples' in r][0]
h=rows[hi]; ci=h.index('# Samples'); ii=h.index('Instructions Executed')
agg={}
for r in rows[hi+1:]:
    if len(r)<=ii: continue
    if r[0]=='Line No': break
    if r[0].strip().isdigit():
        ln=int(r[0])
        try: s=float(r[ci] or 0); ie=float(r[ii] or 0)
        except Exception: continue
        if ln not in agg: agg[ln]=[0,0,r[1]]
        if r[2] in ('','-'):
            agg[ln][0]=s; agg[ln][1]=ie
tot_s=sum(v[0] for v in agg.values()); tot_i=sum(v[1] for v in agg.values())
print('total samples %d, total warp-instructions %.4g, per knot %.1f'%(tot_s,tot_i,tot_i/knots))
for ln,(s,ie,t) in sorted(agg.items()):
    if ie/knots>=mn: print(f"L{ln:4d} {ie/knots:7.1f} i/knot {100*ie/tot_i:5.1f}% inst {100*s/tot_s:5.1f}% smp : {t.strip()[:100]}")
