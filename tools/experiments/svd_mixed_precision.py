"""numpy experiment (TEST INFRASTRUCTURE, not product code): FP32 Jacobi + FP64 polish on thetas collected from oracle
trajectories (TFI D=4, AFH D=8).  Result: for cond(theta) >= 1e6 the FP64 phase needs as many sweeps as
without the FP32 phase - the accumulated FP32 rotations mix 1e-7 of the large columns into the small ones.
    python tools/experiments/svd_mixed_precision.py"""
import sys, numpy as np
_R = __file__.rsplit('/', 3)[0]; sys.path.insert(0, _R + '/oracle'); sys.path.insert(0, _R + '/tensor-networks-simple-update_b200')
import tnsu_oracle as orc, device_model as dm
from tnsu_b200 import smc
from scipy import linalg

def collect(D, model, n_sweeps_list, dts, seed=3):
    smat = smc.infinite_structure_matrix_dict("peps")
    np.random.seed(seed)
    t,w = orc.random_network(smat,2,D)
    thetas=[]
    real_svd = linalg.svd
    rec = {'on':False}
    def spy(a,*k,**kw):
        if rec['on']: thetas.append(np.array(np.real(a)))
        return real_svd(a,*k,**kw)
    orc.linalg.svd = spy
    sw=0
    for dt,ns in zip(dts,n_sweeps_list):
        for i in range(ns):
            rec['on'] = (i>=ns-1)
            orc.sweep(t,w,smat,model,dt,D)
    orc.linalg.svd = real_svd
    return thetas

def jacobi_oe(x, acc, tol, max_sweeps, dtype):
    x=np.array(x,dtype=dtype); acc=np.array(acc,dtype=dtype)
    n=x.shape[1]
    nrot=0
    for sweep in range(max_sweeps):
        rotated=False
        for step in range(n+(n&1)):
            for p in range(step%2,n-1,2):
                q=p+1
                al=dtype(x[:,p]@x[:,p]); be=dtype(x[:,q]@x[:,q]); ga=dtype(x[:,p]@x[:,q])
                c,s=dtype(1),dtype(0)
                if ga*ga > dtype(tol*tol)*al*be and ga!=0:
                    rotated=True; nrot+=1
                    delta=be-al
                    inv_r=dtype(1)/np.sqrt(delta*delta+dtype(4)*ga*ga)
                    c2=dtype(0.5)+dtype(0.5)*abs(delta)*inv_r
                    c=np.sqrt(c2)
                    s=(-ga if delta<0 else ga)*inv_r/c
                for m in (x,acc):
                    pc,qc=m[:,p].copy(),m[:,q].copy()
                    m[:,p]=s*pc+c*qc
                    m[:,q]=c*pc-s*qc
        if not rotated: return x,acc,sweep+1
    return x,acc,max_sweeps

def mixed(theta, precond=True, tol32=2e-7, ns_steps=2):
    A = theta.T.copy()   # like kernel: A' = theta^T
    n = A.shape[1]
    if precond:
        q,r = dm.qr_column_pivoting(A.astype(np.float32).astype(float))  # emulate FP32 QR roughly by rounding input (cheap emulation)
        x0 = r.T.astype(np.float32); acc0 = q.astype(np.float32)
        # jacobi on X = R~^T with accumulator Q:  X W = Y ; we want W (n x n orthogonal): accumulate from identity
        x32,w32,sw32 = jacobi_oe(x0, np.eye(x0.shape[1],dtype=np.float32), tol32, 30, np.float32)
        # total right transform on A'^T...: A' = Q R~ -> X = R~^T = A'^T Q ; X W = Y.  Left vectors of A' = Q W (columns), so
        # define Z = Q W  (rows x steps): orthonormal-ish; then B = A'^T Z has orthogonal columns = V*Sigma of A'
        Z = q @ w32.astype(float)
    else:
        x32,w32,sw32 = jacobi_oe(A.astype(np.float32), np.eye(n,dtype=np.float32), tol32, 30, np.float32)
        Z = w32.astype(float)
    # Newton-Schulz orthonormalisation in FP64
    for _ in range(ns_steps):
        Z = Z @ (1.5*np.eye(Z.shape[1]) - 0.5*(Z.T@Z))
    orth = np.abs(Z.T@Z-np.eye(Z.shape[1])).max()
    if precond:
        B = A.T @ Z      # columns ~ orthogonal
        acc = Z
    else:
        B = A @ Z; acc = Z
    G = B.T@B; d=np.sqrt(np.diag(G)); cos0 = np.abs(G/np.outer(d,d)-np.eye(len(d))).max()
    x,acc,sw64 = dm.jacobi_odd_even(B, acc)
    sv = np.sort(np.sqrt(np.sum(x*x,axis=0)))[::-1]
    return sw32, sw64, orth, cos0, sv

if __name__=="__main__":
    pz, px = np.array([[1.0, 0], [0, -1.0]]), np.array([[0, 1.0], [1.0, 0]])
    sx,sy,sz = orc.spin_operators(0.5)
    for name,D,model in (("tfi D4",4,orc.Model([-1.0]*8,-3.0,[pz],[pz],[px])),("afh D8",8,orc.Model([1.0]*8,0.0,[sx,sy,sz],[sx,sy,sz],[]))):
        ths = collect(D, model, [1,6,40,60], [0.1,0.1,0.1,0.01])
        print(name, len(ths))
        for k,th in enumerate(ths):
            if k%4: continue
            u,s,vt,sw = dm.svd_preconditioned(th)
            ref = np.linalg.svd(th,compute_uv=False)
            for pre in (True,False):
                sw32,sw64,orth,cos0,sv = mixed(th,precond=pre)
                print(f"  theta {k:2d} cond {ref[0]/max(ref[-1],1e-300):.1e} base sweeps {sw} | mixed precond={pre}: fp32 sweeps {sw32} fp64 sweeps {sw64} orth {orth:.1e} cos0 {cos0:.1e} sv err {np.abs(sv-ref).max()/ref[0]:.1e}")
