"""numpy model of the ALGORITHM the CUDA kernels implement.  TEST INFRASTRUCTURE ONLY.

The reference (and `tnsu_oracle.py`) use LAPACK RQ + gesdd.  The sm_100a kernels use a different,
gauge-equivalent route ("Kernel math"):

  * in-place Householder LQ on the rows of the absorbed bond matrix P (reflector k is left in row k),
  * compact-WY form of the reflectors, so that P' = r~ . Q becomes one sync-free small GEMM
        P' = [r~ 0] - (r~ . Y1 . T^H) . Y^H,
  * one-sided (Hestenes) Jacobi SVD of theta with accumulated right vectors; for real thetas up to 32 x 32 in the
    odd-even ordering after a pivoted-QR preconditioner (`svd_preconditioned`, mirrors csrc/svd_warp.cuh).

This module states that route step by step in numpy with the same index conventions as
`csrc/edge_update.cuh`, so the math can be checked against the oracle on the CPU
(`tests/test_device_model.py`) before any GPU time is spent.  It is not a fallback: nothing in the
product package imports it.
"""
from __future__ import annotations

import numpy as np

from tnsu_oracle import edge_endpoints, tensor_legs


def bond_matrix(tensor, weights, smat, t, ek):
    """P[(s, x_k), c] = T[s, ...] * prod_{m != k} lam_m[x_m]; columns c enumerate the other legs in
    memory order (any fixed order is admissible: Q never leaves the kernel)."""
    edges, axes = tensor_legs(smat, t, ek)
    a = int(smat[t, ek])
    scaled = tensor.astype(complex) if np.iscomplexobj(tensor) else tensor.astype(float)
    for e, ax in zip(edges, axes):
        shape = [1] * tensor.ndim
        shape[ax] = len(weights[e])
        scaled = scaled * np.reshape(weights[e], shape)
    moved = np.moveaxis(scaled, a, 1)
    return np.reshape(moved, (moved.shape[0] * moved.shape[1], -1)), moved.shape


def householder_lq_inplace(p):
    """P (M x N) -> L (M x K), reflector rows Yh (K x N, row k = u_k^H, zero for c < k), tau (K).
    P . H_0 ... H_{K-1} = [L 0],  H_k = I - tau_k u_k u_k^H.  Mirrors the kernel: one Gram-row
    reduction g_i = sum_{c>k} P[i,c] conj(P[k,c]) over ALL rows per step (rows i<k give the
    reflector inner products needed for the compact-WY T factor)."""
    p = np.array(p)
    m, n = p.shape
    k_dim = min(m, n)
    l_mat = np.zeros((m, k_dim), dtype=p.dtype)
    tau = np.zeros(k_dim)
    t_mat = np.zeros((k_dim, k_dim), dtype=p.dtype)  # upper triangular
    for k in range(k_dim):
        g = p[:, k + 1:] @ np.conj(p[k, k + 1:])  # one block reduction, M values
        alpha = p[k, k]
        norm2 = np.real(g[k]) + abs(alpha) ** 2
        norm = np.sqrt(norm2)
        if norm == 0.0:
            tau[k] = 0.0
            l_mat[k:, k] = 0.0
            p[k, k] = 0.0
            continue
        absa = abs(alpha)
        phase_c = (np.conj(alpha) / absa) if absa > 0 else 1.0  # conj(alpha)/|alpha|
        # u_k (column vector): u[c] = conj(P[k,c]) for c>k, u[k] = conj(alpha) + phase_c*norm
        uk_head = np.conj(alpha) + phase_c * norm
        tau[k] = 1.0 / (norm * (norm + absa))
        # z_i = u_i^H u_k for i<k :  tail + stored head  (P[i,k] = conj(u_i[k]) for i<k)
        z = g[:k] + p[:k, k] * uk_head
        # dots for rows i>k: p_i . u_k = g_i + P[i,k]*uk_head
        w = tau[k] * (g[k + 1:] + p[k + 1:, k] * uk_head)
        # update rows i>k: P[i,c] -= w_i * conj(u_k[c])
        p[k + 1:, k + 1:] -= np.outer(w, p[k, k + 1:])
        p[k + 1:, k] -= w * np.conj(uk_head)
        l_mat[k, k] = -np.conj(phase_c) * norm  # = conj(beta'), beta' = -phase_c*norm
        l_mat[k + 1:, k] = p[k + 1:, k]
        p[k, k] = np.conj(uk_head)  # row k now stores u_k^H on c>=k
        p[k, :k] = 0.0
        # compact WY:  T[:k,k] = -tau_k T[:k,:k] z ;  T[k,k] = tau_k
        t_mat[:k, k] = -tau[k] * (t_mat[:k, :k] @ z)
        t_mat[k, k] = tau[k]
    return l_mat, p[:k_dim, :], tau, t_mat


def recompose(r_tilde, yh, t_mat):
    """P' = r~ . Q with Q = [I_K 0] (I - Y T^H Y^H):  P' = [r~ 0] - (r~ Y1 T^H) Y^H."""
    k_dim = yh.shape[0]
    y1 = np.conj(yh[:, :k_dim]).T  # K x K, lower triangular: Y1[c,k] = u_k[c]
    c_mat = r_tilde @ y1 @ np.conj(t_mat).T
    out = -c_mat @ yh
    out[:, :k_dim] += r_tilde
    return out


def jacobi_svd(a, tol=1e-15, max_sweeps=60):
    """One-sided Jacobi on the columns of A (rows >= cols): A V = U diag(s).  Round-robin pair
    order; rotation is the standard Hermitian 2x2 diagonalisation.  Returns u, s, vh sorted."""
    a = np.array(a)
    transposed = a.shape[0] < a.shape[1]
    if transposed:
        a = np.conj(a.T)
    n_rows, n = a.shape
    v = np.eye(n, dtype=a.dtype)
    for _ in range(max_sweeps):
        rotated = False
        for p in range(n - 1):
            for q in range(p + 1, n):
                alpha = np.real(np.vdot(a[:, p], a[:, p]))
                beta = np.real(np.vdot(a[:, q], a[:, q]))
                gamma = np.vdot(a[:, p], a[:, q])  # a_p^H a_q
                ag = abs(gamma)
                if ag <= tol * np.sqrt(alpha * beta) or ag == 0.0:
                    continue
                rotated = True
                zeta = (beta - alpha) / (2.0 * ag)
                t = np.sign(zeta) / (abs(zeta) + np.sqrt(1.0 + zeta * zeta)) if zeta != 0 else 1.0
                c = 1.0 / np.sqrt(1.0 + t * t)
                s = c * t
                ph = gamma / ag  # e^{i phi}
                for mat in (a, v):
                    cp_, cq_ = mat[:, p].copy(), mat[:, q].copy()
                    mat[:, p] = c * cp_ - s * np.conj(ph) * cq_
                    mat[:, q] = s * ph * cp_ + c * cq_
        if not rotated:
            break
    sv = np.sqrt(np.real(np.sum(a * np.conj(a), axis=0)))
    order = np.argsort(-sv, kind="stable")
    sv = sv[order]
    a = a[:, order]
    v = v[:, order]
    u = a / np.where(sv > 0, sv, 1.0)
    if transposed:
        return v, sv, np.conj(u.T)
    return u, sv, np.conj(v.T)


def qr_column_pivoting(a):
    """Householder QR with column pivoting in the form `svd_warp_kernel` uses it: the columns never move
    (the kernel keeps one column per lane), pivots are only CHOSEN by trailing norm, so the result is
    R~ = R P^T with A = Q R~.  Returns (q_thin, r_tilde), q_thin: rows x steps, r_tilde: steps x cols."""
    a = np.array(a, dtype=float)
    rows, cols = a.shape
    steps = min(rows, cols)
    z = np.eye(rows)  # accumulates Q^T = H_{steps-1} ... H_0
    done = np.zeros(cols, dtype=bool)
    for j in range(steps):
        norms = np.where(done, -1.0, np.sum(a[j:, :] ** 2, axis=0))
        piv = int(np.argmax(norms))
        done[piv] = True
        w = a[:, piv].copy()
        w[:j] = 0.0
        n2 = float(w @ w)
        if n2 > 0.0:
            norm = np.sqrt(n2)
            alpha = w[j]
            w[j] = alpha + (norm if alpha >= 0 else -norm)
            tau = 1.0 / (n2 + norm * abs(alpha))
            a -= np.outer(w, tau * (w @ a))
            z -= np.outer(w, tau * (w @ z))
    return z[:steps, :].T, a[:steps, :]


def jacobi_odd_even(x, acc, tol=1.5e-15, max_sweeps=60):
    """One-sided Jacobi on the columns of X with the accumulator `acc` rotated along, in the odd-even
    transposition ordering of `svd_warp_kernel`: step A pairs positions (2k, 2k+1), step B pairs (2k+1, 2k+2),
    every rotation writes its outputs swapped, so that all pairs meet once in n steps.  Returns the rotated
    (x, acc, sweeps)."""
    x, acc = np.array(x, dtype=float), np.array(acc, dtype=float)
    n = x.shape[1]
    for sweep in range(max_sweeps):
        rotated = False
        for step in range(n + (n & 1)):
            for p in range(step % 2, n - 1, 2):
                q = p + 1
                al, be, ga = x[:, p] @ x[:, p], x[:, q] @ x[:, q], x[:, p] @ x[:, q]
                c, s = 1.0, 0.0
                if ga * ga > tol * tol * al * be and ga != 0.0:
                    rotated = True
                    delta = be - al
                    inv_r = 1.0 / np.sqrt(delta * delta + 4.0 * ga * ga)
                    c2 = 0.5 + 0.5 * abs(delta) * inv_r
                    c = np.sqrt(c2)
                    s = (-ga if delta < 0 else ga) * inv_r / c
                for m in (x, acc):
                    pc, qc = m[:, p].copy(), m[:, q].copy()
                    m[:, p] = s * pc + c * qc
                    m[:, q] = c * pc - s * qc
        if not rotated:
            return x, acc, sweep + 1
    return x, acc, max_sweeps


def jacobi_odd_even_tracked(x, tol=1.5e-15, max_sweeps=60, exact_from=12, small=2.6e-30):
    """The iteration of the V-free kernel route (svdw_step_a): the odd-even ordering of `jacobi_odd_even` without an
    accumulator, with the squared column norms CARRIED ALONG (updated from each rotation, refreshed from the columns
    once per sweep, recomputed per pair from sweep `exact_from` on), and with the exit rule of the kernel: stop after a
    sweep without rotation, or after one whose (max cos^2 of a pair met) x (max sin^2 of a rotation applied) <= `small`.
    Returns (x, sweeps)."""
    x = np.array(x, dtype=float)
    n = x.shape[1]
    for sweep in range(max_sweeps):
        nrm = np.sum(x * x, axis=0)
        max_cos2 = max_sin2 = 0.0
        for step in range(n + (n & 1)):
            for p in range(step % 2, n - 1, 2):
                q = p + 1
                if sweep >= exact_from:
                    nrm[p], nrm[q] = x[:, p] @ x[:, p], x[:, q] @ x[:, q]
                al, be, ga = nrm[p], nrm[q], x[:, p] @ x[:, q]
                c, s = 1.0, 0.0
                if ga * ga > tol * tol * al * be and ga != 0.0:
                    delta = be - al
                    inv_r = 1.0 / np.sqrt(delta * delta + 4.0 * ga * ga)
                    c2 = 0.5 + 0.5 * abs(delta) * inv_r
                    c = np.sqrt(c2)
                    s = (-ga if delta < 0 else ga) * inv_r / c
                    max_cos2 = max(max_cos2, ga * ga / (al * be) if al * be > 0 else np.inf)
                    max_sin2 = max(max_sin2, s * s)
                pc, qc = x[:, p].copy(), x[:, q].copy()
                x[:, p] = s * pc + c * qc
                x[:, q] = c * pc - s * qc
                m = 2.0 * c * s * ga
                nrm[p], nrm[q] = s * s * al + c * c * be + m, c * c * al + s * s * be - m
        if max_cos2 * max_sin2 <= small:  # includes the sweep without rotation
            return x, sweep + 1
    return x, max_sweeps


def svd_preconditioned(theta):
    """theta = U diag(s) V^T the way `svd_warp_kernel` computes it: A' = theta^T, A' = Q R~ (pivoted QR),
    Jacobi on X = R~^T with the accumulator started at Q:  X W = Y  =>  A' = (Q W)(Y)^T, so the right vectors of
    theta are the final accumulator and the left vectors the normalised final X columns."""
    q_thin, r_tilde = qr_column_pivoting(np.array(theta, dtype=float).T)
    x, acc, sweeps = jacobi_odd_even(r_tilde.T, q_thin)
    sv = np.sqrt(np.sum(x * x, axis=0))
    order = np.argsort(-sv, kind="stable")
    sv, x, acc = sv[order], x[:, order], acc[:, order]
    u = x / np.where(sv > 0, sv, 1.0)
    return u, sv, acc.T, sweeps


def cholesky_lower_and_inverse(g, span=1e-5):
    """Right-looking Cholesky G = L L^T with the checks of lq_chol_kernel: returns (L, L^-1) or None on a non-positive
    pivot or when diag(L) spans more than 1/span (cond(P) too large for CholeskyQR2: the Householder kernels take it)."""
    g = np.array(g, dtype=float)
    m = g.shape[0]
    low = np.zeros_like(g)
    for k in range(m):
        d = g[k, k]
        if not (0.0 < d < 1e300):
            return None
        lkk = np.sqrt(d)
        low[k:, k] = g[k:, k] / lkk
        g[k + 1:, k + 1:] -= np.outer(low[k + 1:, k], low[k + 1:, k])
    diag = np.diag(low)
    if not diag.min() > span * diag.max():
        return None
    inv = np.zeros_like(low)
    for j in range(m):  # forward substitution, column by column
        for r in range(j, m):
            acc = (1.0 if r == j else 0.0) - low[r, j:r] @ inv[j:r, j]
            inv[r, j] = acc / low[r, r]
    return low, inv


def cholqr2_lq(p):
    """P = L Q the way lq_chol_kernel computes it: G1 = P P^T = L1 L1^T, Q1 = L1^-1 P, G2 = Q1 Q1^T = L2 L2^T,
    L = L1 L2, L^-1 = L2^-1 L1^-1.  Returns (L, L^-1) or None when the kernel would flag the matrix."""
    p = np.array(p, dtype=float)
    if p.shape[0] > p.shape[1]:
        return None
    first = cholesky_lower_and_inverse(p @ p.T)
    if first is None:
        return None
    l1, w1 = first
    q1 = w1 @ p
    second = cholesky_lower_and_inverse(q1 @ q1.T)
    if second is None:
        return None
    l2, w2 = second
    return l1 @ l2, w2 @ w1


def svd_vfree(theta, keep):
    """The default K2 route (svd_warp_kernel<R, false>): pivoted QR of theta^T, odd-even Jacobi on X = R~^T WITHOUT
    an accumulator, the `keep` largest columns give sigma and u = x / sigma; the matching right vectors are recovered
    from theta itself, w = theta^T u / ||theta^T u|| (a fixed dense start vector when that is zero), followed by two
    passes of modified Gram-Schmidt in rank order.  Returns u (n_i x keep), s (keep), w^T (keep x n_j), sweeps."""
    theta = np.array(theta, dtype=float)
    _, r_tilde = qr_column_pivoting(theta.T)
    x, sweeps = jacobi_odd_even_tracked(r_tilde.T)
    sv = np.sqrt(np.sum(x * x, axis=0))
    order = np.argsort(-sv, kind="stable")[:keep]
    sv, x = sv[order], x[:, order]
    u = x * np.where(sv > 0, 1.0 / np.where(sv > 0, sv, 1.0), 0.0)
    w = theta.T @ u
    for j in range(w.shape[1]):
        if not np.sum(w[:, j] ** 2) > 1e-280:
            c = np.arange(1, w.shape[0] + 1, dtype=np.uint64)
            h = ((c * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)) ^ np.uint64(((j + 1) * 40503) & 0xFFFFFFFF)
            w[:, j] = ((h >> np.uint64(8)) & np.uint64(0xFFFF)).astype(float) - 32767.5
    for _ in range(2):
        for j in range(w.shape[1]):
            n2 = np.sum(w[:, j] ** 2)
            if n2 > 0:
                w[:, j] /= np.sqrt(n2)
            for i in range(j + 1, w.shape[1]):
                w[:, i] -= (w[:, i] @ w[:, j]) * w[:, j]
    return u, sv, w.T, sweeps


def edge_update(tensors, weights, smat, ek, gate, d_max, vfree=False):
    """Same contract as tnsu_oracle.edge_update, computed the way the kernel does (`vfree`: the V-free SVD route
    of the real-float64 warp kernel instead of the accumulating Jacobi)."""
    ti, ai, tj, aj = edge_endpoints(smat, ek)
    lam = weights[ek]
    sides = []
    for t in (ti, tj):
        p, moved_shape = bond_matrix(tensors[t], weights, smat, t, ek)
        l_mat, yh, tau, t_mat = householder_lq_inplace(p)
        sides.append((l_mat, yh, t_mat, moved_shape))
    (l_i, yh_i, t_i, sh_i), (l_j, yh_j, t_j, sh_j) = sides
    d_i, d_j = sh_i[0], sh_j[0]
    k_i, k_j = l_i.shape[1], l_j.shape[1]
    dk = len(lam)
    r_i = np.reshape(l_i, (d_i, dk, k_i))
    r_j = np.reshape(l_j, (d_j, dk, k_j))
    theta = np.einsum("ieq,e,jep,ijab->qabp", r_i, lam, r_j, gate)
    th2 = np.reshape(theta, (k_i * d_i, d_j * k_j))
    if vfree:
        full = min(th2.shape)
        d_new = min(d_max, full) if d_max is not None else full
        u, s, vh, _ = svd_vfree(np.real(th2), d_new)
    else:
        u, s, vh = jacobi_svd(th2)
        d_new = min(d_max, len(s)) if d_max is not None else len(s)
        u, s, vh = u[:, :d_new], s[:d_new], vh[:d_new, :]
    rt_i = np.reshape(np.transpose(np.reshape(u, (k_i, d_i, d_new)), (1, 2, 0)), (d_i * d_new, k_i))
    rt_j = np.reshape(np.transpose(np.reshape(vh, (d_new, d_j, k_j)), (1, 0, 2)), (d_j * d_new, k_j))
    for t, a, rt, yh, t_mat, sh in ((ti, ai, rt_i, yh_i, t_i, sh_i), (tj, aj, rt_j, yh_j, t_j, sh_j)):
        p_new = recompose(rt, yh, t_mat)
        new_shape = list(sh)
        new_shape[1] = d_new
        out = np.moveaxis(np.reshape(p_new, new_shape), 1, a)
        edges, axes = tensor_legs(smat, t, ek)
        for e, ax in zip(edges, axes):
            shape = [1] * out.ndim
            shape[ax] = len(weights[e])
            out = out * np.reshape(1.0 / weights[e], shape)
        tensors[t] = out / np.sqrt(np.real(np.sum(out * np.conj(out))))
    weights[ek] = s / np.sum(s)
