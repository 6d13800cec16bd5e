"""numpy experiment (TEST INFRASTRUCTURE): warm-start the one-sided Jacobi of an edge with the right vectors of the same
edge's previous sweep, on the device-model trajectory (Householder LQ + preconditioned Jacobi as the kernels do it).
Result: cos(theta_new V_old) = 1.0, 11-12 sweeps instead of 6 - the LQ basis rotates between sweeps.
    python tools/experiments/svd_warm_start.py"""
import sys, numpy as np
_R = __file__.rsplit('/', 3)[0]; sys.path.insert(0, _R + '/oracle'); sys.path.insert(0, _R + '/tensor-networks-simple-update_b200')
import tnsu_oracle as orc, device_model as dm
from tnsu_oracle import edge_endpoints, tensor_legs
from tnsu_b200 import smc

def edge_update_warm(tensors, weights, smat, ek, gate, d_max, store, stats):
    ti, ai, tj, aj = edge_endpoints(smat, ek)
    lam = weights[ek]
    sides = []
    for t in (ti, tj):
        p, moved_shape = dm.bond_matrix(tensors[t], weights, smat, t, ek)
        l_mat, yh, tau, t_mat = dm.householder_lq_inplace(p)
        sides.append((l_mat, yh, t_mat, moved_shape))
    (l_i, yh_i, t_i, sh_i), (l_j, yh_j, t_j, sh_j) = sides
    d_i, d_j = sh_i[0], sh_j[0]
    k_i, k_j = l_i.shape[1], l_j.shape[1]
    dk = len(lam)
    r_i = np.reshape(l_i, (d_i, dk, k_i)); r_j = np.reshape(l_j, (d_j, dk, k_j))
    theta = np.real(np.einsum("ieq,e,jep,ijab->qabp", r_i, lam, r_j, gate))
    th2 = np.reshape(theta, (k_i * d_i, d_j * k_j))
    # baseline
    u0, s0, vt0, sw0 = dm.svd_preconditioned(th2)
    if ek in store:
        vprev = store[ek]
        x = th2 @ vprev
        g = x.T @ x; dd = np.sqrt(np.diag(g)); cos0 = np.abs(g/np.outer(dd,dd) - np.eye(len(dd))).max()
        x, acc, sw = dm.jacobi_odd_even(x, vprev)
        sv = np.sqrt(np.sum(x*x, axis=0)); order = np.argsort(-sv, kind="stable")
        sv, x, acc = sv[order], x[:, order], acc[:, order]
        u = x / np.where(sv > 0, sv, 1.0); vt = acc.T
        stats.append((sw0, sw, cos0))
    else:
        u, sv, vt = u0, s0, vt0
        acc = vt0.T
    store[ek] = acc.copy()
    d_new = min(d_max, len(sv))
    u, s, vh = u[:, :d_new], sv[:d_new], vt[:d_new, :]
    rt_i = np.reshape(np.transpose(np.reshape(u, (k_i, d_i, d_new)), (1, 2, 0)), (d_i * d_new, k_i))
    rt_j = np.reshape(np.transpose(np.reshape(vh, (d_new, d_j, k_j)), (1, 0, 2)), (d_j * d_new, k_j))
    for t, a, rt, yh, t_mat, sh in ((ti, ai, rt_i, yh_i, t_i, sh_i), (tj, aj, rt_j, yh_j, t_j, sh_j)):
        p_new = dm.recompose(rt, yh, t_mat)
        new_shape = list(sh); new_shape[1] = d_new
        out = np.moveaxis(np.reshape(p_new, new_shape), 1, a)
        edges, axes = tensor_legs(smat, t, ek)
        for e, ax in zip(edges, axes):
            shape = [1] * out.ndim; shape[ax] = len(weights[e])
            out = out * np.reshape(1.0 / weights[e], shape)
        tensors[t] = np.real(out / np.sqrt(np.real(np.sum(out * np.conj(out)))))
    weights[ek] = s / np.sum(s)

smat = smc.infinite_structure_matrix_dict("peps")
pz, px = np.array([[1.0, 0], [0, -1.0]]), np.array([[0, 1.0], [1.0, 0]])
for h in (-1.0, -3.0):
    model = orc.Model([-1.0]*8, h, [pz],[pz],[px])
    np.random.seed(5)
    t,w = orc.random_network(smat,2,4)
    store={}
    for dt, ns in ((0.1,40),(0.01,40),(0.001,20)):
        stats=[]
        for sw in range(ns):
            for ek in range(8):
                gate = orc.ite_gate(model.edge_hamiltonian(smat, ek), dt, 2, 2)
                edge_update_warm(t,w,smat,ek,gate,4,store,stats)
            if sw in (1,2,5,10,20,39,19) and len(stats)>=8:
                a=np.array(stats[-8:])
                print(f"h={h} dt={dt} sweep {sw}: baseline jacobi sweeps {a[:,0].mean():.1f} warm {a[:,1].mean():.1f} cos0 max {a[:,2].max():.1e}")
