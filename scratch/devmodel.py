import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import onmf_oracle as O
TINY=np.float32(np.finfo(np.float32).tiny)
def lars_inv(G, c, reg, d, T=np.float32, max_iter=1000, recompute_w=True):
    G=G.astype(T); c=c.astype(T); k=len(c)
    cov=c.copy(); coef=np.zeros(k,T); prev=np.zeros(k,T)
    active=[]; inactive=np.ones(k,bool)
    M=np.zeros((0,0),T)
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
            g=G[active,j]; u=M@g if len(active) else np.zeros(0,T)
            sigma=G[j,j]-g@u
            piv=max(np.sqrt(abs(sigma)),T(np.finfo(T).eps))
            if piv<1e-7: cov[j]=0; continue
            s=len(active)
            Mn=np.zeros((s+1,s+1),T)
            inv=T(1)/sigma
            Mn[:s,:s]=M+np.outer(u,u)*inv; Mn[:s,s]=-u*inv; Mn[s,:s]=-u*inv; Mn[s,s]=inv
            M=Mn; active.append(j); inactive[j]=False
        if n_iter>0 and a_prev<a_cur: break
        w=M.sum(1,dtype=T)
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
def run(X,W,alpha,T):
    G=W.T@W; Cv=W.T@X
    return np.stack([lars_inv(G,Cv[:,j],alpha,X.shape[0],T) for j in range(X.shape[1])],1)
if __name__=='__main__':
    g=np.load('/root/repo/tests/golden/cfg1_renoir_gray.npz')
    X=g['X']; 
    for name,W in [('W0 rand',g['W0']),('W step1',g['W_0']),('W step8',g['W_7'])]:
        Xb=X[:,:300]; Href=O.sparse_code_sklearn(Xb,W,1.0)
        for T in (np.float64,np.float32):
            H=run(Xb,W,1.0,T)
            e=np.linalg.norm(H-Href)/np.linalg.norm(Href); ce=np.linalg.norm(H-Href,axis=0)/np.maximum(np.linalg.norm(Href,axis=0),1e-30)
            print(name, T.__name__, 'cond %.1e'%np.linalg.cond(W.T@W), 'rel %.2e worstcol %.2e  supp-mismatch cols %d'%(e,ce.max(), ((H>0)!=(Href>0)).any(0).sum()))
    rng=np.random.default_rng(1)
    W=rng.random((1024,256)); Xb=rng.random((1024,24)); Href=O.sparse_code_sklearn(Xb,W,1.0)
    for T in (np.float64,np.float32):
        H=run(Xb,W,1.0,T); e=np.linalg.norm(H-Href)/np.linalg.norm(Href)
        print('cfg5 rand W0',T.__name__,'rel %.2e'%e, 'supp mismatch', ((H>0)!=(Href>0)).any(0).sum())
    W/=np.linalg.norm(W,axis=0); Href=O.sparse_code_sklearn(Xb,W,1.0)
    for T in (np.float64,np.float32):
        H=run(Xb,W,1.0,T); e=np.linalg.norm(H-Href)/np.linalg.norm(Href)
        print('cfg5 norm W',T.__name__,'rel %.2e'%e, 'supp mismatch', ((H>0)!=(Href>0)).any(0).sum())
    print('--- intrinsic test')
    g=np.load('/root/repo/tests/golden/cfg1_renoir_gray.npz'); X=g['X']; W=g['W_7']; Xb=X[:,:300]
    Href=O.sparse_code_sklearn(Xb,W,1.0)
    G32=(W.astype(np.float32).T@W.astype(np.float32)).astype(np.float64); C32=(W.astype(np.float32).T@Xb.astype(np.float32)).astype(np.float64)
    H=np.stack([O.lars_lasso_positive(G32,C32[:,j],1.0,100) for j in range(300)],1)
    print('fp32 G,c + fp64 chol LARS: rel %.2e'%(np.linalg.norm(H-Href)/np.linalg.norm(Href)))
    H=np.stack([lars_inv(G32,C32[:,j],1.0,100,np.float64) for j in range(300)],1)
    print('fp32 G,c + fp64 inv LARS: rel %.2e'%(np.linalg.norm(H-Href)/np.linalg.norm(Href)))
    G64=W.T@W; C64=W.T@Xb
    H=np.stack([lars_inv(G64.astype(np.float32),C64[:,j].astype(np.float32),1.0,100,np.float32) for j in range(300)],1)
    print('fp64-accurate G,c rounded to fp32 + fp32 inv LARS: rel %.2e'%(np.linalg.norm(H-Href)/np.linalg.norm(Href)))
    # effect on aggregates
    H32=run(Xb,W,1.0,np.float32)
    A=Href@Href.T; A2=H32@H32.T; B=Href@Xb.T; B2=H32@Xb.T
    print('A rel %.2e  B rel %.2e  recon %.6f vs %.6f'%(np.linalg.norm(A-A2)/np.linalg.norm(A), np.linalg.norm(B-B2)/np.linalg.norm(B), np.linalg.norm(Xb-W@Href)/np.linalg.norm(Xb), np.linalg.norm(Xb-W@H32)/np.linalg.norm(Xb)))
