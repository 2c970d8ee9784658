import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scratch')
import numpy as np, warnings; warnings.filterwarnings('ignore')
from oracle import onmf_oracle as O, c_oracle
import devmodel2 as D2
g=np.load('/root/repo/tests/golden/cfg2_renoir_color_tensor.npz'); i=2
W=g['W_%d'%(i-1)]; Xb=g['X'][:,g['idx'][i]][:, :100]; Href=g['H_%d'%i][:, :100]
print('cond G', np.linalg.cond(W.T@W))
f32=np.float32
G64=W.T@W; C64=W.T@Xb
G32=(W.astype(f32).T@W.astype(f32)); C32=(W.astype(f32).T@Xb.astype(f32))
def rel(a,b): return np.linalg.norm(a-b)/np.linalg.norm(b)
def solve(G,C,**kw): return np.stack([D2.lars_var(np.asarray(G,np.float64),np.asarray(C[:,j],np.float64),1.0,300,**kw) for j in range(C.shape[1])],1)
print('exact G,c -> f64 solver', rel(solve(G64,C64,T=np.float64,TM=np.float64),Href))
print('G32,c32 (fp32 gemm) -> f64 solver', rel(solve(G32,C32,T=np.float64,TM=np.float64),Href))
print('G64 rounded, c32 -> f64 solver', rel(solve(G64.astype(f32),C32,T=np.float64,TM=np.float64),Href))
print('G32, c64 rounded -> f64 solver', rel(solve(G32,C64.astype(f32),T=np.float64,TM=np.float64),Href))
print('G64r,c64r -> f64 solver', rel(solve(G64.astype(f32),C64.astype(f32),T=np.float64,TM=np.float64),Href))
print('G64r,c64r -> f32 solver refine1', rel(solve(G64.astype(f32),C64.astype(f32),refine=1),Href))
print('G32,c32 -> f32 solver refine1', rel(solve(G32,C32,refine=1),Href))
print('G32,c32 -> f32 solver refine0', rel(solve(G32,C32,refine=0),Href))
# input X rounding alone
Xr=Xb.astype(f32).astype(np.float64); Wr=W.astype(f32).astype(np.float64)
print('only X,W rounded to fp32, exact after', rel(c_oracle.sparse_code(Xr,Wr,1.0),Href))
