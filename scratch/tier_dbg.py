import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from onmf_ontf_ndl_b200 import OnmfEngine
from oracle import c_oracle
dev=torch.device('cuda:0')
def tt(x,dt): return torch.from_numpy(np.ascontiguousarray(x)).to(dev,dt)
def rel(a,b): return float(np.linalg.norm(a-b)/np.linalg.norm(b))
rng=np.random.default_rng(5)
d,k,n=1024,256,12
W=rng.random((d,k)); W/=np.linalg.norm(W,axis=0); X=rng.random((d,n))
Href=c_oracle.sparse_code(X,W,0.0)
print('ref nnz per col', (Href>0).sum(0))
for dt in (torch.float64, torch.float32):
    eng=OnmfEngine(d,k,alpha=0.0,dtype=dt,device=dev,collect_stats=True)
    H=eng.sparse_code(tt(X.T,dt),tt(W,dt)).cpu().numpy().T.astype(np.float64)
    print(dt, eng.read_stats(), 'rel', rel(H,Href), 'per-col', np.linalg.norm(H-Href,axis=0)/np.linalg.norm(Href,axis=0))
d,k,n=96,300,40
W=rng.random((d,k)); W/=np.linalg.norm(W,axis=0); X=rng.random((d,n))
Href=c_oracle.sparse_code(X,W,0.2)
for dt in (torch.float64, torch.float32):
    eng=OnmfEngine(d,k,alpha=0.2,dtype=dt,device=dev,collect_stats=True)
    H=eng.sparse_code(tt(X.T,dt),tt(W,dt)).cpu().numpy().T.astype(np.float64)
    print(dt, eng.read_stats(), 'rel', rel(H,Href))
