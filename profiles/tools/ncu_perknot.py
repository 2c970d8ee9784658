"""per-source-line warp-instructions per LARS knot from an ncu report (all source files): python profiles/tools/ncu_perknot.py rep knots [min]"""
import csv, subprocess, sys, io, os
rep=sys.argv[1]; knots=float(sys.argv[2]); mn=float(sys.argv[3]) if len(sys.argv)>3 else 3.0
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
agg={}; fname=None; h=None
for r in rows:
    if len(r)>=2 and r[0]=='File Path': fname=os.path.basename(r[1]); h=None; continue
    if len(r)>5 and r[0]=='Line No': h=r; ci=h.index('# Samples'); ii=h.index('Instructions Executed'); wi=h.index('L1 Wavefronts Shared'); gi=h.index('L1 Tag Requests Global'); continue
    if h is None or len(r)<=ii: continue
    if r[0].strip().isdigit() and r[2] in ('','-'):
        try: s=float(r[ci] or 0); ie=float(r[ii] or 0); wf=float(r[wi] or 0); tg=float(r[gi] or 0)
        except Exception: continue
        agg[(fname,int(r[0]))]=[s,ie,r[1],wf,tg]
tot_s=sum(v[0] for v in agg.values()); tot_i=sum(v[1] for v in agg.values())
print('total samples %d, total warp-instructions %.4g, per knot %.1f; shared wavefronts per knot %.1f, global tag requests per knot %.1f'%(tot_s,tot_i,tot_i/knots,sum(v[3] for v in agg.values())/knots,sum(v[4] for v in agg.values())/knots))
for (f,ln),v in sorted(agg.items()):
    if v[1]/knots>=mn or v[0]/max(tot_s,1)>0.01:
        print('%-14s L%4d %6.1f i/knot %5.1f%% inst %5.1f%% smp  wf %5.1f : %s'%(f[:14],ln,v[1]/knots,100*v[1]/tot_i,100*v[0]/max(tot_s,1),v[3]/knots,v[2].strip()[:110]))
