"""The numpy model of the kernels' algorithm (Householder LQ kept as reflectors, compact-WY recompose,
one-sided Jacobi SVD) against the oracle: proves on the CPU that the route the CUDA code takes is
gauge-equivalent to the reference's RQ + gesdd route."""
import numpy as np
import pytest

import device_model as dm
import tnsu_oracle as orc
from conftest import build_case


@pytest.mark.parametrize("name", ["star_complex", "chain_growth", "obc4x3_D3", "peps_growth", "kagome_spin1"])
def test_device_model_matches_oracle(name):
    case = build_case(name)
    smat = case["smat"]
    np.random.seed(case["seed"])
    tensors, weights = orc.random_network(smat, case["d"], case["virtual_dim"])
    if case.get("complex_init"):
        tensors = [t + 1j * np.random.normal(size=t.shape) for t in tensors]
    t2, w2 = [t.copy() for t in tensors], [w.copy() for w in weights]
    model = orc.Model(case["j_ij"], case["h_k"], case["s_i"], case["s_j"], case["s_k"], case["hamiltonian"])
    di, dj = model.dims
    for _ in range(min(case["sweeps"], 3)):
        orc.sweep(tensors, weights, smat, model, case["dt"], case["d_max"])
        for ek in range(smat.shape[1]):
            gate = orc.ite_gate(model.edge_hamiltonian(smat, ek), case["dt"], di, dj)
            dm.edge_update(t2, w2, smat, ek, gate, case["d_max"])
        for a, b in zip(weights, w2):
            assert a.shape == b.shape and np.abs(a - b).max() < 1e-12
        e1 = orc.energy_per_site(tensors, weights, smat, model)
        e2 = orc.energy_per_site(t2, w2, smat, model)
        assert abs(e1 - e2) < 1e-12


def test_householder_lq_properties():
    rng = np.random.default_rng(0)
    for m, n in ((6, 40), (8, 8), (12, 5), (4, 4)):
        p = rng.normal(size=(m, n)) + 1j * rng.normal(size=(m, n))
        l_mat, yh, tau, t_mat = dm.householder_lq_inplace(p)
        k = min(m, n)
        q = dm.recompose(np.eye(k, dtype=complex), yh, t_mat)  # r~ = I  ->  Q itself
        assert np.abs(q @ q.conj().T - np.eye(k)).max() < 1e-13
        assert np.abs(l_mat @ q - p).max() < 1e-12
        assert np.abs(np.triu(l_mat[:k, :k], 1)).max() == 0.0


def test_jacobi_svd_properties():
    rng = np.random.default_rng(1)
    for shape in ((8, 8), (12, 24), (24, 12), (5, 5)):
        a = rng.normal(size=shape) + 1j * rng.normal(size=shape)
        u, s, vh = dm.jacobi_svd(a)
        assert np.abs((u * s) @ vh - a).max() < 1e-12
        assert np.abs(s - np.linalg.svd(a, compute_uv=False)).max() < 1e-12
        assert np.all(np.diff(s) <= 0)


def test_preconditioned_odd_even_svd():
    """The K2 route of svd_warp.cuh (pivoted QR, then odd-even one-sided Jacobi on R~^T started at Q): a valid SVD
    with orthonormal factors, LAPACK's singular values, and fewer sweeps than the plain iteration on a graded matrix."""
    rng = np.random.default_rng(3)
    for shape in ((32, 32), (16, 16), (24, 12), (12, 24), (9, 9)):
        a = rng.normal(size=shape) @ np.diag(np.logspace(0, -9, shape[1])) @ rng.normal(size=(shape[1], shape[1]))
        u, s, vt, sweeps = dm.svd_preconditioned(a)
        k = min(shape)
        assert u.shape == (shape[0], k) and vt.shape == (k, shape[1])
        assert np.abs(u.T @ u - np.eye(k)).max() < 1e-12 and np.abs(vt @ vt.T - np.eye(k)).max() < 1e-12
        assert np.abs((u * s) @ vt - a).max() < 1e-12 * np.abs(a).max()
        assert np.abs(s - np.linalg.svd(a, compute_uv=False)).max() < 1e-12 * s[0]
    a = rng.normal(size=(32, 32)) @ np.diag(np.logspace(0, -8, 32)) @ rng.normal(size=(32, 32))
    _, _, plain = dm.jacobi_odd_even(a, np.eye(32))
    assert dm.svd_preconditioned(a)[3] < plain


@pytest.mark.parametrize("name", ["chain_growth", "obc4x3_D3", "peps_growth", "tfi_D4", "pbc3_D2_dt0"])
def test_vfree_svd_route_matches_oracle(name):
    """The default K2 route of the real path iterates A alone and recovers the kept right singular vectors from
    theta afterwards (svd_warp_kernel<R, false>, device_model.svd_vfree).  Sweeps with that route agree with the
    oracle's LAPACK route to 1e-12 on weights and energy, including bond growth (exactly rank-deficient thetas)."""
    case = build_case(name)
    smat = case["smat"]
    np.random.seed(case["seed"])
    tensors, weights = orc.random_network(smat, case["d"], case["virtual_dim"])
    t2, w2 = [t.copy() for t in tensors], [w.copy() for w in weights]
    model = orc.Model(case["j_ij"], case["h_k"], case["s_i"], case["s_j"], case["s_k"], case["hamiltonian"])
    di, dj = model.dims
    for _ in range(min(case["sweeps"], 3)):
        orc.sweep(tensors, weights, smat, model, case["dt"], case["d_max"])
        for ek in range(smat.shape[1]):
            gate = orc.ite_gate(model.edge_hamiltonian(smat, ek), case["dt"], di, dj)
            dm.edge_update(t2, w2, smat, ek, np.real(gate), case["d_max"], vfree=True)
        for a, b in zip(weights, w2):
            assert a.shape == b.shape and np.abs(a - b).max() < 1e-12
        e1 = orc.energy_per_site(tensors, weights, smat, model)
        e2 = orc.energy_per_site(t2, w2, smat, model)
        assert abs(e1 - e2) < 1e-12


def test_tracked_norms_and_exit_rule_of_the_jacobi_iteration():
    """svdw_step_a carries the column norms along and skips the confirming sweep when the last sweep's rotations were
    small (device_model.jacobi_odd_even_tracked): never more sweeps than the plain iteration, at least one fewer on
    average, columns orthogonal to 1e-13 and LAPACK's singular values — on generic, graded and rank-deficient input."""
    rng = np.random.default_rng(11)
    saved = []
    for n, kind in ((32, "generic"), (32, "graded"), (32, "rank8"), (16, "generic"), (16, "graded")):
        for _ in range(3):
            u, _ = np.linalg.qr(rng.normal(size=(n, n)))
            v, _ = np.linalg.qr(rng.normal(size=(n, n)))
            sv = {"generic": rng.uniform(0.01, 1.0, n), "graded": np.logspace(0, -14, n),
                  "rank8": np.r_[rng.uniform(0.1, 1.0, 8), np.zeros(n - 8)]}[kind]
            a = (u * sv) @ v.T
            _, r_tilde = dm.qr_column_pivoting(a.T)
            _, _, plain = dm.jacobi_odd_even(r_tilde.T, np.zeros((1, n)))
            x, sweeps = dm.jacobi_odd_even_tracked(r_tilde.T)
            assert sweeps <= plain
            saved.append(plain - sweeps)
            got = np.sqrt(np.sum(x * x, axis=0))
            ref = np.linalg.svd(a, compute_uv=False)
            assert np.abs(np.sort(got)[::-1] - ref).max() < 1e-13 * ref[0]
            big = got > 1e-8 * got.max()
            g = (x[:, big] / got[big]).T @ (x[:, big] / got[big])
            assert np.abs(g - np.eye(int(big.sum()))).max() < 1e-13
    assert np.mean(saved) >= 0.5


def test_vfree_svd_properties():
    """svd_vfree on graded and on exactly rank-deficient matrices: LAPACK's leading singular values, orthonormal kept
    vectors on both sides, and theta w_k = sigma_k u_k wherever sigma_k is not negligible."""
    rng = np.random.default_rng(5)
    for shape, keep in (((32, 32), 8), ((16, 16), 4), ((24, 12), 6), ((12, 24), 6)):
        a = rng.normal(size=shape) @ np.diag(np.logspace(0, -9, shape[1])) @ rng.normal(size=(shape[1], shape[1]))
        u, s, wt, _ = dm.svd_vfree(a, keep)
        ref = np.linalg.svd(a, compute_uv=False)[:keep]
        assert np.abs(s - ref).max() < 1e-12 * ref[0]
        assert np.abs(u.T @ u - np.eye(keep)).max() < 1e-12 and np.abs(wt @ wt.T - np.eye(keep)).max() < 1e-12
        assert np.abs(a @ wt.T - u * s).max() < 1e-9 * ref[0]
    a = np.outer(rng.normal(size=16), rng.normal(size=16)) + np.outer(rng.normal(size=16), rng.normal(size=16))  # rank 2
    u, s, wt, _ = dm.svd_vfree(a, 6)
    assert np.abs(s[:2] - np.linalg.svd(a, compute_uv=False)[:2]).max() < 1e-12 and s[2:].max() < 1e-12
    assert np.abs(wt @ wt.T - np.eye(6)).max() < 1e-12
    assert np.abs((u[:, :2] * s[:2]) @ wt[:2] - a).max() < 1e-12


def test_cholqr2_lq_properties():
    """The streaming K1 route (lq_chol.cuh, device_model.cholqr2_lq): for cond(P) up to ~1e5 the two Cholesky passes give
    P = L Q with Q = L^-1 P orthonormal to a few 1e-13 and L lower triangular; beyond that (or for more rows than
    columns) the route declines and the Householder kernels are used."""
    rng = np.random.default_rng(11)
    for m, n, cond in ((16, 512, 1e1), (16, 512, 1e3), (16, 512, 5e4), (12, 1024, 1e2), (8, 64, 1e2)):
        u, _ = np.linalg.qr(rng.normal(size=(m, m)))
        v, _ = np.linalg.qr(rng.normal(size=(n, m)))
        p = (u * np.logspace(0, -np.log10(cond), m)) @ v.T
        out = dm.cholqr2_lq(p)
        assert out is not None, (m, n, cond)
        low, inv = out
        q = inv @ p
        assert np.abs(q @ q.T - np.eye(m)).max() < 5e-13
        assert np.abs(low @ q - p).max() < 1e-13 * np.abs(p).max()
        assert np.abs(np.triu(low, 1)).max() == 0.0
    u, _ = np.linalg.qr(rng.normal(size=(16, 16)))
    v, _ = np.linalg.qr(rng.normal(size=(512, 16)))
    assert dm.cholqr2_lq((u * np.logspace(0, -9, 16)) @ v.T) is None      # cond 1e9: flagged
    assert dm.cholqr2_lq(rng.normal(size=(12, 5))) is None                 # M > N: flagged


def test_cholqr2_route_matches_oracle_and_recompose_identity():
    """One edge update with the streaming route — L from CholeskyQR2, recompose T'[:, c] = (r~ L^-1) T[:, c] with the
    environment weights cancelling — equals the oracle's RQ route on weights and energy (peps D=4 state, 2 sweeps)."""
    case = build_case("tfi_D4")
    smat = case["smat"]
    np.random.seed(case["seed"])
    tensors, weights = orc.random_network(smat, case["d"], case["virtual_dim"])
    t2, w2 = [t.copy() for t in tensors], [w.copy() for w in weights]
    model = orc.Model(case["j_ij"], case["h_k"], case["s_i"], case["s_j"], case["s_k"], case["hamiltonian"])
    di, dj = model.dims
    from scipy import linalg
    for _ in range(2):
        orc.sweep(tensors, weights, smat, model, case["dt"], case["d_max"])
        for ek in range(smat.shape[1]):
            gate = np.real(orc.ite_gate(model.edge_hamiltonian(smat, ek), case["dt"], di, dj))
            ti, ai, tj, aj = orc.edge_endpoints(smat, ek)
            lam = w2[ek]
            facs = []
            for t in (ti, tj):
                p, shape = dm.bond_matrix(t2[t], w2, smat, t, ek)
                low, inv = dm.cholqr2_lq(np.real(p))
                facs.append((low, inv, shape))
            (l_i, inv_i, sh_i), (l_j, inv_j, sh_j) = facs
            dk = len(lam)
            theta = np.einsum("ieq,e,jep,ijab->qabp", np.reshape(l_i, (di, dk, -1)), lam, np.reshape(l_j, (dj, dk, -1)), gate)
            k_i, k_j = l_i.shape[1], l_j.shape[1]
            u, s, vh = linalg.svd(np.reshape(theta, (k_i * di, dj * k_j)), full_matrices=False)
            d_new = min(case["d_max"], len(s))
            u, s, vh = u[:, :d_new], s[:d_new], vh[:d_new]
            rt_i = np.reshape(np.transpose(np.reshape(u, (k_i, di, d_new)), (1, 2, 0)), (di * d_new, k_i))
            rt_j = np.reshape(np.transpose(np.reshape(vh, (d_new, dj, k_j)), (1, 0, 2)), (dj * d_new, k_j))
            for t, a, rt, inv, sh in ((ti, ai, rt_i, inv_i, sh_i), (tj, aj, rt_j, inv_j, sh_j)):
                raw = np.reshape(np.moveaxis(np.real(t2[t]), a, 1), (sh[0] * sh[1], -1))   # T itself: no weights
                new_shape = list(sh)
                new_shape[1] = d_new
                out = np.moveaxis(np.reshape((rt @ inv) @ raw, new_shape), 1, a)
                t2[t] = out / np.sqrt(np.sum(out * out))
            w2[ek] = s / np.sum(s)
        for a, b in zip(weights, w2):
            assert np.abs(a - b).max() < 1e-12
        assert abs(orc.energy_per_site(tensors, weights, smat, model) - orc.energy_per_site(t2, w2, smat, model)) < 1e-12
