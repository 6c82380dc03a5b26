"""Numerics prototype (CPU): blocked Cholesky with fp64 vs int8-digit (Ozaki, 8 digits) trailing updates on ill-conditioned and
badly scaled SPD matrices: normwise residual max|A - L L^T| / max|A| and scaled residual max |A - L L^T|_ij / sqrt(a_ii a_jj)."""
import sys; sys.path.insert(0,'tools')
import numpy as np, scipy.linalg as sl
import proto_ozaki_i8 as P
rng=np.random.default_rng(0)
def blocked(a0, nb, S):
    a=a0.copy(); n=a.shape[0]; nt=n//nb
    for kk in range(nt):
        s=slice(kk*nb,(kk+1)*nb)
        a[s,s]=np.linalg.cholesky(a[s,s])
        if kk+1<nt:
            r=slice((kk+1)*nb,n)
            a[r,s]=sl.solve_triangular(a[s,s],a[r,s].T,lower=True).T
            p=a[r,s]
            a[r,r]-= (p@p.T) if S==0 else P.ozaki_gemm(p,p,S)
    return np.tril(a)
n,nb=1024,128
cases={}
q,_=np.linalg.qr(rng.standard_normal((n,n)))
for cond in (1e4,1e10,1e14):
    d=np.logspace(0,-np.log10(cond),n)
    a=(q*d)@q.T; a=(a+a.T)/2
    cases[f"random SPD cond {cond:.0e}"]=a
i=np.arange(1,n+1); cases["Lehmer (min/max)"]=np.minimum.outer(i,i)/np.maximum.outer(i,i)
dsc=np.ldexp(1.0,rng.integers(-30,30,n)); b=cases["random SPD cond 1e+04"]; cases["badly scaled D A D (60 binades), cond(A)=1e4"]=b*dsc[:,None]*dsc[None,:]
for name,a in cases.items():
    out=[]
    for S in (0,8):
        try:
            L=blocked(a,nb,S)
        except np.linalg.LinAlgError:
            out.append("not SPD numerically"); continue
        res=np.abs(np.tril(L@L.T-a)).max()/np.abs(a).max()
        dd=np.sqrt(np.diag(a)); sres=(np.abs(np.tril(L@L.T-a))/np.outer(dd,dd)).max()
        out.append(f"resid {res:.2e} scaled-resid {sres:.2e}")
    print(f"{name:48s} fp64 bulk: {out[0]:44s} int8-digit bulk: {out[1]}")
