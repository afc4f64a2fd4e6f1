import numpy as np, scipy.linalg as sl, sys
sys.path.insert(0,'/root/repo/oracle')
import cokrig_oracle as orc

def split(A, s, first_bits=6, bits=7):
    # row-wise scaling; returns slices (float arrays of small ints), row exponents
    amax = np.abs(A).max(axis=1)
    e = np.where(amax>0, np.floor(np.log2(np.where(amax>0,amax,1)))+1, 0)
    r = A / np.exp2(e)[:,None]          # |r|<1
    out=[]; sh = first_bits
    scale_tot = 0
    for p in range(s):
        scale_tot += (first_bits if p==0 else bits)
        d = np.rint(r*np.exp2(scale_tot))
        r = r - d*np.exp2(-scale_tot)
        out.append(d)
    return out, e

def oz_gemm_nt(A,B,s,first_bits=6,bits=7, full=False):
    As,ea = split(A,s,first_bits,bits); Bs,eb = split(B,s,first_bits,bits)
    # weight of slice p (0-idx): 2^-(first_bits + bits*p)
    C = np.zeros((A.shape[0],B.shape[0]))
    for g in range(2*s-1 if full else s-1, -1, -1):
        acc = np.zeros_like(C)
        for p in range(s):
            q=g-p
            if 0<=q<s: acc += As[p]@Bs[q].T
        assert np.abs(acc).max() < 2**31
        C += acc*np.exp2(-(2*first_bits+bits*g))
    return C*np.exp2(ea)[:,None]*np.exp2(eb)[None,:]

def chol_blocked(S, nb, gemm):
    A = S.copy(); n=A.shape[0]
    for k in range(0,n,nb):
        k1=min(k+nb,n)
        A[k:k1,k:k1] = sl.cholesky(A[k:k1,k:k1],lower=True)
        if k1<n:
            A[k1:,k:k1] = sl.solve_triangular(A[k:k1,k:k1], A[k1:,k:k1].T, lower=True).T
            A[k1:,k1:] -= gemm(A[k1:,k:k1],A[k1:,k:k1])
    return np.tril(A)

def solve_blocked(L, R, nb, gemm):
    # R: (m, n) target-major; V = R L^-T
    R=R.copy(); n=L.shape[0]
    for k in range(0,n,nb):
        k1=min(k+nb,n)
        R[:,k:k1] = sl.solve_triangular(L[k:k1,k:k1], R[:,k:k1].T, lower=True).T
        if k1<n: R[:,k1:] -= gemm(R[:,k:k1], L[k1:,k:k1])
    return R

if __name__=="__main__":
    nx=int(sys.argv[1]) if len(sys.argv)>1 else 32
    nug=float(sys.argv[2]) if len(sys.argv)>2 else 0.01
    params=[1,1,1.5,1.5,1.5,.2,.2,.2,nug,nug,-.6]
    P=orc.Params(params); grid=orc.expand_grid(xcount=nx,ycount=nx)
    _,_,z=orc.sim_fields(P,grid,seed=1)
    pc=np.random.default_rng(7).uniform(0,1,(300,2))
    S=orc.joint_cov(P,[grid,grid],"euclidean"); 
    C=orc.pred_cross_cov(P,1,[grid,grid],pc,"euclidean")
    print("N",S.shape, "cond", np.linalg.cond(S))
    zz=np.hstack(z)
    Lref=sl.cholesky(S,lower=True)
    Vref=sl.solve_triangular(Lref, np.c_[C, zz], lower=True).T
    pr=Vref[:-1]@Vref[-1]; vr=(1+nug)-(Vref[:-1]**2).sum(1)
    for name,gemm in [("fp64",lambda a,b:a@b.T)]+[(f"oz{s}",(lambda s:(lambda a,b:oz_gemm_nt(a,b,s)))(s)) for s in (6,7,8,9)]:
        L=chol_blocked(S,512,gemm)
        V=solve_blocked(L,np.c_[C,zz].T,512,gemm)
        p=V[:-1]@V[-1]; v=(1+nug)-(V[:-1]**2).sum(1)
        print(name, "L err", np.abs(L-Lref).max(), "pred rel", np.max(np.abs(p-pr)/np.abs(pr)), "pred rel-to-scale", np.max(np.abs(p-pr))/np.abs(pr).max(), "var abs", np.abs(v-vr).max())
