import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scratch')
import numpy as np, warnings; warnings.filterwarnings('ignore')
from oracle import onmf_oracle as O
TINY=np.float32(np.finfo(np.float32).tiny)
def lars_mp(G, c, reg, d, refine=1, resid64=True, TM=np.float32, covfix=False, max_iter=1000):
    T=np.float32
    G=G.astype(T); c=c.astype(T); k=len(c)
    cov=c.copy(); coef=np.zeros(k,T); prev=np.zeros(k,T)
    active=[]; inactive=np.ones(k,bool)
    M=np.zeros((0,0),TM)
    amin=T(reg)/T(d); eps32=T(np.finfo(np.float32).eps)
    a_cur=T(0); a_prev=T(0); n_iter=0; drop=False
    while True:
        if inactive.any():
            vals=np.where(inactive,cov,-np.inf); j=int(np.argmax(vals)); C=vals[j]
        else: C=T(0); j=-1
        a_cur=C/T(d)
        if a_cur<=amin+eps32:
            if abs(a_cur-amin)>eps32 and n_iter>0:
                ss=(a_prev-amin)/(a_prev-a_cur); coef=prev+ss*(coef-prev)
            break
        if n_iter>=max_iter or len(active)>=k: break
        if not drop:
            g=G[active,j].astype(TM); u=M@g if len(active) else np.zeros(0,TM)
            sigma=TM(G[j,j])-g@u
            s=len(active)
            Mn=np.zeros((s+1,s+1),TM); inv=TM(1)/sigma
            Mn[:s,:s]=M+np.outer(u,u)*inv; Mn[:s,s]=-u*inv; Mn[s,:s]=-u*inv; Mn[s,s]=inv
            M=Mn; active.append(j); inactive[j]=False
        RT=np.float64 if resid64 else np.float32
        w=M.sum(1,dtype=TM).astype(RT)
        GA=G[np.ix_(active,active)].astype(RT)
        for _ in range(refine):
            r=np.ones(len(active),RT)-GA@w
            w=w+(M@r.astype(TM)).astype(RT)
        sw=w.sum()
        AAr=RT(1)/np.sqrt(sw); w=(w*AAr).astype(T); AA=T(AAr)
        corr=(G[:,active]@w).astype(T)
        with np.errstate(all='ignore'):
            r=(C-cov)/(AA-corr+TINY)
        r=np.where(inactive&(r>0),r,np.inf); g1=r.min()
        gamma=min(g1,C/AA)
        z=-coef[active]/(w+TINY); zp=np.where(z>0,z,np.inf)
        drop=False
        if zp.min()<gamma:
            gamma=zp.min(); p=int(np.argmin(zp)); drop=True
        n_iter+=1; prev=coef; a_prev=a_cur
        coef=np.zeros(k,T); coef[active]=prev[active]+gamma*w
        if covfix:
            # recompute cov exactly from coefficients: c - G coef (fp32 data, fp64 accumulate)
            cc=(c.astype(np.float64)-G.astype(np.float64)@coef.astype(np.float64)).astype(T)
            cov=np.where(inactive,cc,cov)
        else:
            cov=np.where(inactive,cov-gamma*corr,cov)
        if drop:
            mi=M[:,p].copy(); M=M-np.outer(mi,mi)/mi[p]
            M=np.delete(np.delete(M,p,0),p,1)
            jd=active.pop(p); inactive[jd]=True
            cov[jd]=c[jd]-G[jd]@coef
    return coef
def rel(a,b): return np.linalg.norm(a-b)/np.linalg.norm(b)
for name,i,n in [('cfg2_renoir_color_tensor',2,100),('cfg1_renoir_gray',2,150),('cfg1_renoir_gray',7,150)]:
    g=np.load('/root/repo/tests/golden/%s.npz'%name)
    W=g['W_%d'%(i-1)]; Xb=g['X'][:,g['idx'][i]][:, :n]; Href=g['H_%d'%i][:, :n]; d=W.shape[0]
    G64=W.T@W; C64=W.T@Xb
    G32=(W.astype(np.float32).T@W.astype(np.float32)); C32=(W.astype(np.float32).T@Xb.astype(np.float32))
    print(name,i,'cond %.1e'%np.linalg.cond(G64))
    for label,(G,C) in [('exact-rounded G,c',(G64,C64)),('fp32-gemm G,c',(G32,C32)),('G exact-rounded, c fp32',(G64,C32))]:
        for kw in [dict(refine=1,resid64=False),dict(refine=1,resid64=True),dict(refine=2,resid64=True),dict(refine=2,resid64=True,covfix=True),dict(refine=0,TM=np.float64)]:
            H=np.stack([lars_mp(G,C[:,j],1.0,d,**kw) for j in range(n)],1)
            print('   %-26s %-45s rel %.2e'%(label,kw,rel(H,Href)))
