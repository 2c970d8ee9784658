import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scratch')
import numpy as np
from oracle import onmf_oracle as O
import devmodel as D
TINY=D.TINY
def lars_var(G, c, reg, d, T=np.float32, TM=np.float32, refine=0, max_iter=1000):
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
            Mn=np.zeros((s+1,s+1),TM)
            inv=TM(1)/sigma
            Mn[:s,:s]=M+np.outer(u,u)*inv; Mn[:s,s]=-u*inv; Mn[s,:s]=-u*inv; Mn[s,s]=inv
            M=Mn; active.append(j); inactive[j]=False
        w=M.sum(1,dtype=TM)
        for _ in range(refine):
            r=np.ones(len(active),TM)-G[np.ix_(active,active)].astype(TM)@w
            w=w+M@r
        w=w.astype(T)
        AA=T(1)/np.sqrt(w.sum(dtype=T)); w=w*AA
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
        cov=np.where(inactive,cov-gamma*corr,cov)
        if drop:
            mi=M[:,p].copy(); M=M-np.outer(mi,mi)/mi[p]
            M=np.delete(np.delete(M,p,0),p,1)
            jd=active.pop(p); inactive[jd]=True
            cov[jd]=c[jd]-G[jd]@coef
    return coef
g=np.load('/root/repo/tests/golden/cfg1_renoir_gray.npz'); X=g['X']; W=g['W_7']; Xb=X[:,:300]
Href=O.sparse_code_sklearn(Xb,W,1.0)
G=W.T@W; C=W.T@Xb
for name,kw in [('fp32 all',dict()),('fp32 + refine1',dict(refine=1)),('fp32, M fp64',dict(TM=np.float64)),('fp32 + refine2',dict(refine=2))]:
    H=np.stack([lars_var(G,C[:,j],1.0,100,**kw) for j in range(300)],1)
    A=Href@Href.T; A2=H@H.T
    print(name,'rel %.2e  A rel %.2e'%(np.linalg.norm(H-Href)/np.linalg.norm(Href), np.linalg.norm(A-A2)/np.linalg.norm(A)))
