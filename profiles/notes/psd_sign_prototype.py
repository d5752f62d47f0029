"""Numerical prototype (NumPy, CPU) for the round-2 PSD projection: Pi_+(X) = (X + sign(X) X) / 2 with the matrix sign
function computed by GEMM-only iterations instead of an eigendecomposition.  Matrices: the w_s iterates of the
closest-correlation SDP (config C4, N = 300 here) at ADMM iterations 1, 2, 10, 30, 60.

Output on this container (relative Frobenius error against eigh):
  iter  1  min|lam|/|X| 2.5e-04 | Newton-Schulz 26 its err 1.2e-15 | QDWH 4 its err 1.7e-14
  iter  2               1.1e-03 |               23     1.2e-15 |      4     4.1e-15
  iter 10               1.5e-04 |               28     9.2e-16 |      4     3.0e-14
  iter 30               1.5e-05 |               33     9.2e-16 |      5     2.0e-13
  iter 60               1.9e-03 |               21     1.3e-15 |      4     2.6e-15
"""
import numpy as np, time, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import cosmo_b200
from oracle import cosmo_oracle as O
from oracle.bridge import to_oracle_cones
# closest correlation problem: take w_s iterates (the matrices that get projected)
N=300
P,q,A,b,sets=cosmo_b200.problems.closest_correlation_sdp(N=N,seed=3)
cones=to_oracle_cones(sets)
mats=[]
def cb(it,ws):
    if it in (1,2,10,30,60):
        n=ws.n
        for rng,c in zip(O.row_ranges(ws.cones),ws.cones):
            if isinstance(c,O.PsdConeTriangle):
                mats.append((it,O.populate_upper_triangle(ws.w[n:][rng].copy(),c.sqrt_dim,1/np.sqrt(2))))
ws=O.Workspace(P,q,A,b,cones,O.Settings(scaling=0,max_iter=60)); ws.optimize(iter_callback=cb)
def proj_ref(X):
    w,V=np.linalg.eigh(X); return (V*np.maximum(w,0))@V.T, w
def newton_schulz_sign(X, tol=1e-13, maxit=100):
    nrm=np.linalg.norm(X,2)  # spectral norm (power iteration on GPU)
    S=X/nrm; k=0
    while k<maxit:
        S2=S@S
        Sn=0.5*S@(3*np.eye(len(X))-S2)
        k+=1
        d=np.linalg.norm(Sn-S,'fro')/np.sqrt(len(X))
        S=Sn
        if d<tol: break
    return S,k
def qdwh_sign(X, maxit=20):
    # dynamically weighted Halley (Nakatsukasa-Higham), Cholesky variant
    n=len(X); alpha=np.linalg.norm(X,2); S=X/alpha
    l=1e-16/1.0  # crude lower bound on smin/alpha (would be estimated)
    smin=np.min(np.abs(np.linalg.eigvalsh(X)))/alpha; l=max(smin*0.9,1e-17)
    k=0
    while abs(1-l)>1e-15 and k<maxit:
        l2=l*l; dd=(4*(1-l2)/(l2*l2))**(1/3); sq=np.sqrt(1+dd)
        a=sq+0.5*np.sqrt(8-4*dd+8*(2-l2)/(l2*sq)); b=(a-1)**2/4; c=a+b-1
        Z=np.eye(n)+c*(S.T@S); Lc=np.linalg.cholesky(Z)
        W=np.linalg.solve(Lc, S.T).T   # S Lc^-T
        W=np.linalg.solve(Lc.T, W.T).T # (S Z^-1)
        S=(b/c)*S+(a-b/c)*W
        l=l*(a+b*l2)/(1+c*l2); k+=1
    return S,k
for it,Xu in mats:
    X=np.triu(Xu)+np.triu(Xu,1).T
    ref,w=proj_ref(X)
    S,k=newton_schulz_sign(X)
    Pp=0.5*(X+S@X); Pp=0.5*(Pp+Pp.T)
    S2,k2=qdwh_sign(X)
    Pq=0.5*(X+S2@X); Pq=0.5*(Pq+Pq.T)
    print('iter',it,'N',len(X),'min|lam|/norm %.1e'%(np.min(np.abs(w))/np.max(np.abs(w))),'npos',int((w>0).sum()),
          '| NS its',k,'err %.2e'%(np.linalg.norm(Pp-ref)/np.linalg.norm(X)),'| QDWH its',k2,'err %.2e'%(np.linalg.norm(Pq-ref)/np.linalg.norm(X)))
