"""numpy experiment (TEST INFRASTRUCTURE): Jacobi sweeps after pivoted QR, after pivoted QR + one / two more unpivoted QRs.
Result: 6.0 -> 4.25 -> 3.9 sweeps at 16 x 16, 6.9 -> 5.75 -> 5.25 at 32 x 32.
    python tools/experiments/svd_second_qr.py"""
import sys, numpy as np
sys.path.insert(0, __file__.rsplit('/', 1)[0])
from svd_mixed_precision import *
pz, px = np.array([[1.0, 0], [0, -1.0]]), np.array([[0, 1.0], [1.0, 0]])
sx,sy,sz = orc.spin_operators(0.5)
for name,D,model in (("tfi D4",4,orc.Model([-1.0]*8,-3.0,[pz],[pz],[px])),("afh D8",8,orc.Model([1.0]*8,0.0,[sx,sy,sz],[sx,sy,sz],[]))):
    ths = collect(D, model, [1,6,40,60], [0.1,0.1,0.1,0.01])
    tot=[0,0,0,0]
    for th in ths:
        u,s,vt,sw = dm.svd_preconditioned(th); tot[0]+=sw
        q,r = dm.qr_column_pivoting(th.T)
        # second QR (no pivoting) of X = r^T
        q2,r2 = np.linalg.qr(r.T)
        x,acc,sw2 = dm.jacobi_odd_even(r2.T, np.eye(r2.shape[0])); tot[1]+=sw2
        # third
        q3,r3 = np.linalg.qr(r2.T)
        x,acc,sw3 = dm.jacobi_odd_even(r3.T, np.eye(r3.shape[0])); tot[2]+=sw3
        # jacobi on rows instead (X = r)
        x,acc,sw4 = dm.jacobi_odd_even(r2, np.eye(r2.shape[1])); tot[3]+=sw4
    print(name, "avg sweeps: QRCP", tot[0]/len(ths), "QRCP+QR", tot[1]/len(ths), "QRCP+QR+QR", tot[2]/len(ths), "QRCP+QR rows", tot[3]/len(ths))
