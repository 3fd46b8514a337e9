"""Accuracy of block cyclic reduction against the reference's sequential LDL' for the background system, on the
CPU (a numpy emulation of the algorithm of csrc/background_kernels.cu, the oracle, and an extended-precision
LDL' as yardstick): typical systems, long stretches of zero weight, one to three refinement steps.

    python tools/cr_accuracy_probe.py
"""
import sys, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import oracle as O
from golden.make_background_golden import background_inputs

def second_diag(n,lam):
    d=np.full(n,6.0*lam)
    if n<3 or lam<=0: return np.zeros(n)
    if n==3: return np.array([lam,4*lam,lam])
    d[[0,-1]]=lam; d[[1,-2]]=5*lam; return d
def second_off1(n,lam):
    if n<3 or lam<=0: return np.zeros(n-1)
    if n==3: return np.full(2,-2*lam)
    o=np.full(n-1,-4.0*lam); o[[0,-1]]=-2*lam; return o
def build(w,rhs,lam,lam1):
    n=len(w)
    a=w.astype(float).copy()+second_diag(n,lam)
    if lam1>0 and n>=2:
        fd=np.full(n,2*lam1); fd[[0,-1]]=lam1; a+=fd
    b=second_off1(n,lam)+(-lam1 if (lam1>0 and n>=2) else 0.0)
    c=np.full(max(n-2,0),lam)
    N=(n+1)//2
    ap=np.concatenate([a,[1.0]]) if n%2 else a
    bp=np.concatenate([b,np.zeros(2*N-n+1)])[:2*N]
    cp=np.concatenate([c,np.zeros(2*N)])[:2*N]
    rp=np.concatenate([rhs,[0.0]]) if n%2 else rhs
    op=np.concatenate([np.ones(n),[0.0]]) if n%2 else np.ones(n)
    D=np.zeros((N,2,2)); U=np.zeros((N,2,2)); B=np.zeros((N,2,2))
    D[:,0,0]=ap[0::2]; D[:,1,1]=ap[1::2]; D[:,0,1]=D[:,1,0]=bp[0::2]
    U[:,0,0]=cp[0::2]; U[:,1,0]=bp[1::2]; U[:,1,1]=cp[1::2]
    U[-1]=0
    B[:,0,0]=rp[0::2]; B[:,1,0]=rp[1::2]; B[:,0,1]=op[0::2]; B[:,1,1]=op[1::2]
    return D,U,B
def solve_cr(D,U,B):
    N=len(D)
    if N==1: return np.linalg.solve(D[0],B[0])[None]
    ev=np.arange(0,N,2); od=np.arange(1,N,2)
    Dn=D[ev].copy(); Bn=B[ev].copy(); Un=np.zeros_like(Dn)
    inv=np.linalg.inv(D[od])
    # left neighbours: even i>0 has odd i-1
    hasl=ev>0; il=(ev[hasl]-1)//2
    alpha=np.einsum('nij,njk->nik', np.transpose(U[ev[hasl]-1],(0,2,1)), inv[il])   # L_i = U_{i-1}^T
    Dn[hasl]-=np.einsum('nij,njk->nik',alpha,U[ev[hasl]-1]); Bn[hasl]-=np.einsum('nij,njk->nik',alpha,B[ev[hasl]-1])
    hasr=ev+1<N; ir=(ev[hasr]+1)//2
    gamma=np.einsum('nij,njk->nik',U[ev[hasr]],inv[ir])
    Dn[hasr]-=np.einsum('nij,njk->nik',gamma,np.transpose(U[ev[hasr]],(0,2,1))); Bn[hasr]-=np.einsum('nij,njk->nik',gamma,B[ev[hasr]+1])
    Un[hasr]=-np.einsum('nij,njk->nik',gamma,U[ev[hasr]+1])
    Xe=solve_cr(Dn,Un,Bn)
    X=np.zeros_like(B); X[ev]=Xe
    T=B[od]-np.einsum('nij,njk->nik',np.transpose(U[od-1],(0,2,1)),X[od-1])
    hr=od+1<N
    T[hr]-=np.einsum('nij,njk->nik',U[od[hr]],X[od[hr]+1])
    X[od]=np.einsum('nij,njk->nik',inv,T)
    return X
def cr(w,rhs,lam,lam1,zc):
    n=len(w); D,U,B=build(w,rhs,lam,lam1); X=solve_cr(D,U,B).reshape(-1,2)[:n]
    if not zc: return X[:,0]
    s0,s1=X[:,0].sum(),X[:,1].sum(); mu=s0/s1 if abs(s1)>1e-12 else s0/n
    return X[:,0]-mu*X[:,1]
rng=np.random.default_rng(0)
print('case | rel diff CR vs LDL (max/max)')
for name,n,lam,lam1,gap in (('typical',200001,128.0,0.0,20),('long gap',200001,128.0,0.0,5000),('huge lam',200001,1e6,0.0,20),('huge lam+gap',200001,1e6,0.0,5000),('tiny weights',200001,128.0,0.0,20),('first only',100001,0.0,1e4,2000)):
    w,rhs=background_inputs(rng,n)
    s=n//2; w[s:s+gap]=0.0
    if name=='tiny weights': w*=1e-6; rhs*=1e-6
    for zc in (True,False):
        a=O.csolveZeroCenteredBackground(w,rhs,lam,zc,lamFirst=lam1); b=cr(w,rhs,lam,lam1,zc)
        print(f'{name:14s} zc={zc!s:5s} {np.abs(a-b).max()/np.abs(a).max():.2e}')

def ldl_ext(w,rhs,lam,lam1):
    # the reference's LDL' recurrences in extended precision (np.longdouble) as a yardstick
    L=np.longdouble
    n=len(w); d=(w.astype(L)+second_diag(n,lam).astype(L))
    if lam1>0:
        fd=np.full(n,2*lam1); fd[[0,-1]]=lam1; d=d+fd.astype(L)
    off=(second_off1(n,lam)+(-lam1 if lam1>0 else 0.0)).astype(L)
    lamL=L(lam); low=np.zeros(n,dtype=L); r=rhs.astype(L).copy()
    low[1]=off[0]/d[0]; d[1]=d[1]-low[1]*low[1]*d[0]
    for i in range(2,n):
        low[i]=(off[i-1]-lamL*low[i-1])/d[i-1]
        d[i]=d[i]-low[i]*low[i]*d[i-1]-(lamL*lamL)/d[i-2]
    r[1]=r[1]-low[1]*r[0]
    for i in range(2,n): r[i]=r[i]-low[i]*r[i-1]-(lamL/d[i-2])*r[i-2]
    r=r/d
    r[n-2]=r[n-2]-low[n-1]*r[n-1]
    for i in range(n-3,-1,-1): r[i]=r[i]-low[i+1]*r[i+1]-(lamL/d[i])*r[i+2]
    return r
print('--- against an extended-precision LDL (zeroCenter off), n=30001')
for gap in (20, 1000, 5000):
    n=30001; w,rhs=background_inputs(rng,n); s=n//2; w[s:s+gap]=0.0
    t=ldl_ext(w,rhs,128.0,0.0)
    a=O.csolveZeroCenteredBackground(w,rhs,128.0,False); b=cr(w,rhs,128.0,0.0,False)
    sc=float(np.abs(t).max())
    print(f'gap {gap:5d}: reference LDL err {float(np.abs(a-t).max())/sc:.2e}   block CR err {float(np.abs(b-t).max())/sc:.2e}   LDL vs CR {np.abs(a-b).max()/sc:.2e}')

def matvec(w,lam,lam1,x):
    n=len(w); a=w.astype(float)+second_diag(n,lam)
    if lam1>0:
        fd=np.full(n,2*lam1); fd[[0,-1]]=lam1; a=a+fd
    b=second_off1(n,lam)+(-lam1 if lam1>0 else 0.0)
    y=a*x; y[:-1]+=b*x[1:]; y[1:]+=b*x[:-1]
    if lam>0 and n>=3: y[:-2]+=lam*x[2:]; y[2:]+=lam*x[:-2]
    return y
def cr_cols(w,r0,r1,lam,lam1):
    n=len(w); D,U,B=build(w,r0,lam,lam1)
    col=np.concatenate([r1,[0.0]]) if n%2 else r1
    B[:,0,1]=col[0::2]; B[:,1,1]=col[1::2]
    return solve_cr(D,U,B).reshape(-1,2)[:n]
print('--- iterative refinement of block CR (zeroCenter off), n=30001')
for gap in (20,1000,5000,12000):
    n=30001; w,rhs=background_inputs(rng,n); s=n//2; w[s:s+gap]=0.0
    t=ldl_ext(w,rhs,128.0,0.0); sc=float(np.abs(t).max())
    a=O.csolveZeroCenteredBackground(w,rhs,128.0,False)
    X=cr_cols(w,rhs,np.ones(n),128.0,0.0)
    errs=[float(np.abs(X[:,0]-t).max())/sc]
    for it in range(3):
        R0=rhs-matvec(w,128.0,0.0,X[:,0]); R1=np.ones(n)-matvec(w,128.0,0.0,X[:,1])
        X=X+cr_cols(w,R0,R1,128.0,0.0)
        errs.append(float(np.abs(X[:,0]-t).max())/sc)
    print(f'gap {gap:5d}: LDL err {float(np.abs(a-t).max())/sc:.2e} | CR err after 0..3 refinements', ' '.join(f'{e:.2e}' for e in errs), '| refined CR vs LDL', f'{np.abs(X[:,0]-a).max()/sc:.2e}')
