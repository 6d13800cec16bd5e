"""numpy experiment (TEST INFRASTRUCTURE): one-sided Jacobi with the column norms carried along (updated from the rotation,
refreshed from the columns once per sweep) instead of recomputed for every pair — two of the three dot products of a
Jacobi step.  Compares sweeps and the accuracy of the kept singular values against recomputed norms on thetas collected
from oracle trajectories (TFI D=4, AFH D=8).
    python tools/experiments/svd_tracked_norms.py"""
import sys, numpy as np
sys.path.insert(0, __file__.rsplit('/', 1)[0])
from svd_mixed_precision import *

def jacobi_tracked(x, tol=1.5e-15, max_sweeps=60, refresh_every=1):
    x = np.array(x, dtype=float)
    n = x.shape[1]
    for sweep in range(max_sweeps):
        rotated = False
        if sweep % refresh_every == 0:
            nrm = np.sum(x * x, axis=0)
        for step in range(n + (n & 1)):
            for p in range(step % 2, n - 1, 2):
                q = p + 1
                al, be, ga = nrm[p], nrm[q], x[:, p] @ x[:, q]
                c, s = 1.0, 0.0
                if ga * ga > tol * tol * al * be and ga != 0.0:
                    rotated = True
                    delta = be - al
                    inv_r = 1.0 / np.sqrt(delta * delta + 4.0 * ga * ga)
                    c2 = 0.5 + 0.5 * abs(delta) * inv_r
                    c = np.sqrt(c2)
                    s = (-ga if delta < 0 else ga) * inv_r / c
                pc, qc = x[:, p].copy(), x[:, q].copy()
                x[:, p] = s * pc + c * qc
                x[:, q] = c * pc - s * qc
                # |s p + c q|^2 and |c p - s q|^2 from (al, be, ga)
                cs2 = 2.0 * c * s * ga
                nrm[p], nrm[q] = s * s * al + c * c * be + cs2, c * c * al + s * s * be - cs2
        if not rotated:
            return x, sweep + 1
    return x, max_sweeps

if __name__ == "__main__":
    pz, px = np.array([[1.0, 0], [0, -1.0]]), np.array([[0, 1.0], [1.0, 0]])
    sx, sy, sz = orc.spin_operators(0.5)
    for name, D, model in (("tfi D4", 4, orc.Model([-1.0] * 8, -3.0, [pz], [pz], [px])),
                           ("afh D8", 8, orc.Model([1.0] * 8, 0.0, [sx, sy, sz], [sx, sy, sz], []))):
        ths = collect(D, model, [1, 6, 40, 60], [0.1, 0.1, 0.1, 0.01])
        tot = [0, 0, 0]
        err = [0.0, 0.0, 0.0]
        for th in ths:
            ref = np.linalg.svd(th, compute_uv=False)
            q, r = dm.qr_column_pivoting(th.T)
            x0, _, sw0 = dm.jacobi_odd_even(r.T, np.eye(r.shape[0]))
            x1, sw1 = jacobi_tracked(r.T)
            x2, sw2 = jacobi_tracked(r.T, refresh_every=100)
            for i, (x, sw) in enumerate(((x0, sw0), (x1, sw1), (x2, sw2))):
                sv = np.sort(np.sqrt(np.sum(x * x, axis=0)))[::-1]
                tot[i] += sw
                err[i] = max(err[i], np.max(np.abs(sv[:D] - ref[:D]) / ref[0]))
        print(name, "thetas", len(ths), "avg sweeps: recomputed", tot[0] / len(ths), "tracked (refresh per sweep)", tot[1] / len(ths),
              "tracked (never refreshed)", tot[2] / len(ths), "max |d sigma| / sigma_1 of the kept values:", err)

def jacobi_tracked_guarded(x, tol=1.5e-15, max_sweeps=60, guard=1e-6):
    """tracked norms, recomputed from the column whenever the update cancelled more than `guard` (LAPACK dgesvj's rule)"""
    x = np.array(x, dtype=float)
    n = x.shape[1]
    recomputed = 0
    for sweep in range(max_sweeps):
        rotated = False
        nrm = np.sum(x * x, axis=0)
        for step in range(n + (n & 1)):
            for p in range(step % 2, n - 1, 2):
                q = p + 1
                al, be, ga = nrm[p], nrm[q], x[:, p] @ x[:, q]
                c, s = 1.0, 0.0
                rot = ga * ga > tol * tol * al * be and ga != 0.0
                if rot:
                    rotated = True
                    delta = be - al
                    inv_r = 1.0 / np.sqrt(delta * delta + 4.0 * ga * ga)
                    c2 = 0.5 + 0.5 * abs(delta) * inv_r
                    c = np.sqrt(c2)
                    s = (-ga if delta < 0 else ga) * inv_r / c
                pc, qc = x[:, p].copy(), x[:, q].copy()
                x[:, p] = s * pc + c * qc
                x[:, q] = c * pc - s * qc
                cs2 = 2.0 * c * s * ga
                hp, hq = max(s * s * al, c * c * be), max(c * c * al, s * s * be)
                nrm[p], nrm[q] = s * s * al + c * c * be + cs2, c * c * al + s * s * be - cs2
                if rot and (nrm[p] <= guard * hp or nrm[q] <= guard * hq):
                    recomputed += 1
                    nrm[p], nrm[q] = x[:, p] @ x[:, p], x[:, q] @ x[:, q]
        if not rotated:
            return x, sweep + 1, recomputed
    return x, max_sweeps, recomputed

if __name__ == "__main__":
    rng = np.random.default_rng(5)
    for label, n, expo in (("graded 1e-16", 32, 16), ("graded 1e-30", 32, 30), ("rank 8 of 32", 32, None), ("graded 16x16 1e-16", 16, 16)):
        res = []
        for trial in range(6):
            u, _ = np.linalg.qr(rng.standard_normal((n, n)))
            v, _ = np.linalg.qr(rng.standard_normal((n, n)))
            sv = 10.0 ** (-np.linspace(0, expo, n)) if expo else np.r_[np.linspace(1, 0.1, 8), np.zeros(n - 8)]
            th = (u * sv) @ v.T
            q, r = dm.qr_column_pivoting(th.T)
            _, _, sw0 = dm.jacobi_odd_even(r.T, np.eye(r.shape[0]))
            _, sw1 = jacobi_tracked(r.T)
            x2, sw2, rec = jacobi_tracked_guarded(r.T)
            got = np.sort(np.sqrt(np.sum(x2 * x2, axis=0)))[::-1]
            res.append((sw0, sw1, sw2, rec, float(np.max(np.abs(got[:8] - sv[:8]) / sv[:8]))))
        print(label, "(recomputed, tracked, guarded, guard recomputations, max rel err of the top 8):", res)
