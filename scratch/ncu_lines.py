"""per-source-line instruction / stall-sample totals from an ncu report: python scratch/ncu_lines.py rep [by=inst|smp] [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; by = sys.argv[2] if len(sys.argv) > 2 else "inst"; top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
hi=None
for i,r in enumerate(rows):
    if 'Source' in r and any('Samples' in c for c in r): hi=i; break
h=rows[hi]; si=h.index('Source')
sc=[i for i,c in enumerate(h) if 'Sampling (All' in c][0]
ic=[i for i,c in enumerate(h) if c.strip()=='Instructions Executed'][0]
data=[]
for n,r in enumerate(rows[hi+1:]):
    try: data.append((float(r[ic] or 0), float(r[sc] or 0), n+1, r[si].strip()))
    except Exception: pass
ti=sum(x[0] for x in data); ts=sum(x[1] for x in data)
print("total inst %.3e samples %d" % (ti, ts))
key = (lambda x: x[0]) if by == "inst" else (lambda x: x[1])
for ie,s,n,t in sorted(data,key=key,reverse=True)[:top]:
    print("%5.1f%% inst %5.1f%% smp  L%4d: %s" % (100*ie/ti, 100*s/ts, n, t[:120]))
