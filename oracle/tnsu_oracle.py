"""CPU oracle for the tnsu Simple-Update hot path.  TEST INFRASTRUCTURE ONLY.

A numpy/scipy restatement of the reference algorithm (RoyElkabetz/Tensor-Networks-Simple-Update,
tnsu 1.0.4).  Nothing in the product package imports this module: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may.

Parity status: PINNED.  `oracle/make_golden.py` runs the unmodified reference (imported from
/root/reference/src with the `oracle/_stubs` ncon/matplotlib stand-ins) and this restatement on
the same seeds and writes both to `tests/golden/`; `tests/test_oracle.py` re-checks the
restatement against those reference outputs and against the reference's own known-answer vector
(tests/test_main.py:17-35: the pickled 10x10 OBC D=4 state, energy -0.5616206254235453 +- 1e-13).

Every function cites the reference lines (relative to /root/reference/src/tnsu/) it follows.
The LAPACK-backed decompositions (scipy.linalg.rq / svd / expm) are the same calls the reference
makes; they live in scipy, not under /root/reference.
"""
from __future__ import annotations

import numpy as np
from scipy import linalg


# --------------------------------------------------------------------------------------
# structure-matrix lookups (simple_update.py:298-338, tensor_network.py:232-240)
# --------------------------------------------------------------------------------------
def edge_endpoints(smat: np.ndarray, ek: int):
    """Two tensors on edge `ek`, ascending tensor index, with their 1-based axis numbers
    (simple_update.py:306-318)."""
    rows = np.nonzero(smat[:, ek])[0]
    return int(rows[0]), int(smat[rows[0], ek]), int(rows[1]), int(smat[rows[1], ek])


def tensor_legs(smat: np.ndarray, t: int, skip_edge: int | None = None):
    """(edges, axes) of tensor `t` in ascending edge order, optionally without one edge
    (simple_update.py:320-330, tensor_network.py:232-240)."""
    edges = [int(e) for e in np.nonzero(smat[t, :])[0] if skip_edge is None or e != skip_edge]
    axes = [int(smat[t, e]) for e in edges]
    return edges, axes


def scale_legs(tensor: np.ndarray, weights, edges, axes, power: float = 1.0) -> np.ndarray:
    """Multiply axis `axes[n]` of the tensor by weights[edges[n]]**power, one leg at a time in
    edge order (tensor_network.py:242-278; power=-1 uses np.power(w, -1), :274)."""
    out = tensor
    for e, a in zip(edges, axes):
        w = weights[e] if power == 1.0 else np.power(weights[e], power)
        shape = [1] * out.ndim
        shape[a] = len(w)
        out = out * np.reshape(w, shape)
    return out


# --------------------------------------------------------------------------------------
# Hamiltonian and gate (simple_update.py:420-473)
# --------------------------------------------------------------------------------------
def two_site_hamiltonian(j_e, h_k, s_i, s_j, s_k, z_i: int, z_j: int, override=None) -> np.ndarray:
    """d_i*d_j x d_i*d_j complex Hamiltonian of one edge (simple_update.py:420-454).
    The field term is divided by the coordination numbers z_i, z_j (:451-453)."""
    if override is not None:
        return override
    di, dj = s_i[0].shape[0], s_j[0].shape[0]
    h_int = np.zeros((di * dj, di * dj), dtype=complex)
    for a, b in zip(s_i, s_j):
        h_int += np.kron(a, b)
    f_i = np.zeros((di * dj, di * dj), dtype=complex)
    f_j = np.zeros((di * dj, di * dj), dtype=complex)
    for s in s_k:
        f_i += np.kron(s, np.eye(dj))
        f_j += np.kron(np.eye(di), s)
    return j_e * h_int + h_k * (f_i / z_i + f_j / z_j)


def ite_gate(hamiltonian: np.ndarray, dt, d_i: int, d_j: int) -> np.ndarray:
    """expm(-dt*H) as a (d_i, d_j, d_i, d_j) tensor (simple_update.py:470-473)."""
    return np.reshape(linalg.expm(-dt * hamiltonian), (d_i, d_j, d_i, d_j))


# --------------------------------------------------------------------------------------
# one edge update (simple_update.py:219-296)
# --------------------------------------------------------------------------------------
def _to_bond_matrix(tensor: np.ndarray, axis: int):
    """swap axis<->1 (a swap, not a roll: simple_update.py:347-349) and flatten to
    (d*D_k, prod(rest)) (:352-369, :243-250).  Returns matrix and the permuted shape."""
    perm = np.arange(tensor.ndim)
    perm[[1, axis]] = perm[[axis, 1]]
    moved = np.transpose(tensor, perm)
    shape = moved.shape
    rest = int(np.prod(shape[2:])) if tensor.ndim > 2 else 1
    return np.reshape(moved, (shape[0] * shape[1], rest)), shape, perm


def edge_update(tensors, weights, smat, ek: int, gate: np.ndarray, d_max):
    """One ITE step on edge `ek`, in place on the `tensors` / `weights` lists
    (simple_update.py:219-296).  `gate` is the (d,d,d,d) tensor from `ite_gate`."""
    ti, ai, tj, aj = edge_endpoints(smat, ek)
    ei, axi = tensor_legs(smat, ti, ek)
    ej, axj = tensor_legs(smat, tj, ek)
    lam = weights[ek]

    # absorb environment weights (:231-232), bring the bond to axis 1 and flatten (:235-240)
    pi, shape_i, perm_i = _to_bond_matrix(scale_legs(tensors[ti], weights, ei, axi), ai)
    pj, shape_j, perm_j = _to_bond_matrix(scale_legs(tensors[tj], weights, ej, axj), aj)
    d_i, d_j = shape_i[0], shape_j[0]

    # RQ reduction to the bond subspace (:243-256)
    r_i, q_i = linalg.rq(pi, mode="economic")
    r_j, q_j = linalg.rq(pj, mode="economic")
    r_i = np.reshape(r_i, (d_i, r_i.shape[0] // d_i, r_i.shape[1]))  # (i, e, qi)
    r_j = np.reshape(r_j, (d_j, r_j.shape[0] // d_j, r_j.shape[1]))  # (j, e, qj)

    # theta = R_i . diag(lam) . R_j . G, gate contracted on its row index (:476-483)
    theta = np.einsum("ieq,e->ieq", r_i, lam)
    theta = np.einsum("ieq,jep->qijp", theta, r_j)
    theta = np.einsum("qijp,ijab->qabp", theta, gate)

    # truncated SVD (:393-409): scipy gesdd, keep the first d_max, no cutoff
    n_i = theta.shape[0] * theta.shape[1]
    n_j = theta.shape[2] * theta.shape[3]
    u, s, vh = linalg.svd(np.reshape(theta, (n_i, n_j)), full_matrices=False)
    if d_max is not None:
        u, s, vh = u[:, :d_max], s[:d_max], vh[:d_max, :]
    d_new = len(s)

    # back to rank-3 (:268-271) and recompose with Q (:274-275)
    ri_t = np.transpose(np.reshape(u, (q_i.shape[0], d_i, d_new)), (1, 2, 0))  # (i', D', qi)
    rj_t = np.transpose(np.reshape(vh, (d_new, d_j, q_j.shape[0])), (1, 0, 2))  # (j', D', qj)
    pi_new = np.einsum("ijk,kl->ijl", ri_t, q_i)
    pj_new = np.einsum("ijk,kl->ijl", rj_t, q_j)

    # restore shapes/axis order (:278-287), un-absorb (:290-291), normalise (:294-296)
    new_i = list(shape_i)
    new_i[1] = d_new
    new_j = list(shape_j)
    new_j[1] = d_new
    t_i = np.transpose(np.reshape(pi_new, new_i), perm_i)
    t_j = np.transpose(np.reshape(pj_new, new_j), perm_j)
    t_i = scale_legs(t_i, weights, ei, axi, power=-1.0)
    t_j = scale_legs(t_j, weights, ej, axj, power=-1.0)
    tensors[ti] = t_i / frobenius(t_i)
    tensors[tj] = t_j / frobenius(t_j)
    weights[ek] = s / np.sum(s)


def frobenius(tensor: np.ndarray):
    """sqrt(sum |T|^2) (simple_update.py:487-496)."""
    return np.sqrt(np.sum(tensor * np.conj(tensor)))


class Model:
    """Hamiltonian parameters of a SimpleUpdate instance (simple_update.py:20-35, 66-77)."""

    def __init__(self, j_ij, h_k, s_i, s_j, s_k, hamiltonian=None):
        self.j_ij, self.h_k, self.s_i, self.s_j, self.s_k = j_ij, h_k, s_i, s_j, s_k
        self.hamiltonian = hamiltonian

    def edge_hamiltonian(self, smat, ek):
        """get_hamiltonian with the neighbour counts the reference passes (:259-260, :627-631):
        all legs of each tensor, the edge itself included."""
        ti, _, tj, _ = edge_endpoints(smat, ek)
        z_i = int(np.sum(smat[ti, :] > 0))
        z_j = int(np.sum(smat[tj, :] > 0))
        j_e = None if self.hamiltonian is not None else self.j_ij[ek]
        return two_site_hamiltonian(j_e, self.h_k, self.s_i, self.s_j, self.s_k, z_i, z_j, self.hamiltonian)

    @property
    def dims(self):
        """spin dims are read from s_i[0]/s_j[0] even with a Hamiltonian override (:467-468)."""
        return self.s_i[0].shape[0], self.s_j[0].shape[0]


def sweep(tensors, weights, smat, model: Model, dt, d_max):
    """One Simple-Update sweep over all edges in column order (simple_update.py:208-296)."""
    d_i, d_j = model.dims
    for ek in range(smat.shape[1]):
        gate = ite_gate(model.edge_hamiltonian(smat, ek), dt, d_i, d_j)
        edge_update(tensors, weights, smat, ek, gate, d_max)


# --------------------------------------------------------------------------------------
# reduced density matrices and expectations (simple_update.py:498-661)
# --------------------------------------------------------------------------------------
def _gram_open_bond(tensor, axis):
    """G[s, e, s', e'] = sum_rest T[s,e,rest] conj(T[s',e',rest]) with the bond on `axis`."""
    m, shape, _ = _to_bond_matrix(tensor, axis)
    g = m @ np.conj(m.T)
    return np.reshape(g, (shape[0], shape[1], shape[0], shape[1]))


def pair_rdm(tensors, weights, smat, ek: int) -> np.ndarray:
    """Two-site RDM rho[i,j,i',j'] on edge `ek`, trace-normalised (simple_update.py:525-593).
    The ncon network there is Gram(T_i) and Gram(T_j) (legs other than ek traced within each
    tensor, weights absorbed) joined by diag(lam_k) on the ket and on the bra side."""
    ti, ai, tj, aj = edge_endpoints(smat, ek)
    ei, axi = tensor_legs(smat, ti, ek)
    ej, axj = tensor_legs(smat, tj, ek)
    g_i = _gram_open_bond(scale_legs(tensors[ti], weights, ei, axi), ai)
    g_j = _gram_open_bond(scale_legs(tensors[tj], weights, ej, axj), aj)
    lam = weights[ek]
    rho = np.einsum("iakb,a,b,jalb->ijkl", g_i, lam, lam, g_j)
    d2 = rho.shape[0] * rho.shape[1]
    return rho / np.trace(np.reshape(rho, (d2, d2)))


def site_rdm(tensors, weights, smat, t: int) -> np.ndarray:
    """One-site RDM with ALL weights absorbed (simple_update.py:498-523)."""
    edges, axes = tensor_legs(smat, t)
    a = scale_legs(tensors[t], weights, edges, axes)
    m = np.reshape(a, (a.shape[0], -1))
    rho = m @ np.conj(m.T)
    return rho / np.trace(rho)


def pair_expectation(tensors, weights, smat, ek, operator) -> complex:
    """sum rho[i,j,i',j'] O[i,j,i',j'] (simple_update.py:605-613)."""
    return np.einsum("ijkl,ijkl->", pair_rdm(tensors, weights, smat, ek), operator)


def energy_per_site(tensors, weights, smat, model: Model) -> float:
    """simple_update.py:615-637."""
    d_i, d_j = model.dims
    energy = 0
    for ek in range(smat.shape[1]):
        h = np.reshape(model.edge_hamiltonian(smat, ek), (d_i, d_j, d_i, d_j))
        energy += pair_expectation(tensors, weights, smat, ek, h)
    return float(np.real(energy / len(tensors)))


def pair_expectation_per_site(tensors, weights, smat, operator) -> float:
    """simple_update.py:639-649."""
    total = 0
    for ek in range(smat.shape[1]):
        total += pair_expectation(tensors, weights, smat, ek, operator)
    return float(np.real(total) / len(tensors))


def expectation_per_site(tensors, weights, smat, operator) -> float:
    """simple_update.py:651-661 (trace(rho @ O) per tensor, :595-603)."""
    total = 0
    for t in range(len(tensors)):
        total += np.trace(site_rdm(tensors, weights, smat, t) @ operator)
    return float(np.real(total) / len(tensors))


# --------------------------------------------------------------------------------------
# run loop (simple_update.py:107-206)
# --------------------------------------------------------------------------------------
def convergence_error(old_weights, weights) -> float:
    """max over edges of the L2 distance (simple_update.py:196-206, utils.py:78-85)."""
    return float(np.max([np.sqrt(np.sum(np.square(o - n))) for o, n in zip(old_weights, weights)]))


def run(tensors, weights, smat, model: Model, dts, d_max, max_iterations=1000, tol=1e-6):
    """SimpleUpdate.run() without the prints (simple_update.py:107-194).  Returns a logger dict
    with the reference's keys plus `converged` and `sweeps`."""
    log = {"error": [], "dt": [], "iteration": [], "energy": [], "converged": False, "sweeps": 0}
    step = 0
    for dt in dts:
        for i in range(max_iterations):
            step += 1
            old = [np.copy(w) for w in weights]
            sweep(tensors, weights, smat, model, dt, d_max)
            if i % 2 == 0 and i > 0:
                err = convergence_error(old, weights)
                log["error"].append(err)
                log["dt"].append(dt)
                log["iteration"].append(step)
                log["energy"].append(energy_per_site(tensors, weights, smat, model))
                if err <= tol and dt == dts[-1]:
                    log["converged"] = True
                    log["sweeps"] = step
                    return log
                if err <= tol:
                    break
    log["sweeps"] = step
    return log


# --------------------------------------------------------------------------------------
# input construction, identical RNG consumption to TensorNetwork.__init__
# --------------------------------------------------------------------------------------
def random_network(smat, spin_dim: int, virtual_dim: int):
    """Random tensors N(1,1) from the legacy global numpy stream, one draw per tensor in row order,
    and uniform weights 1/D (tensor_network.py:77-115).  Seed with np.random.seed first."""
    n, m = smat.shape
    tensors = []
    for t in range(n):
        shape = [spin_dim] + [virtual_dim] * int(np.sum(smat[t, :] > 0))
        tensors.append(np.random.normal(loc=1.0 * np.ones(shape), scale=1.0))
    weights = [np.ones(virtual_dim, dtype=float) / virtual_dim for _ in range(m)]
    return tensors, weights


def spin_operators(spin: float):
    """Spin-S matrices from the ladder-operator formula in the docstring of math_objects.py:4-27."""
    n = int(round(2 * spin + 1))
    m = spin - np.arange(n)  # magnetic quantum numbers spin, spin-1, ..., -spin
    sz = np.diag(m).astype(complex)
    # <m+1| S+ |m> = sqrt(s(s+1) - m(m+1)); row index a has m_a = spin - a
    off = np.sqrt(spin * (spin + 1) - m[1:] * (m[1:] + 1))
    sp = np.zeros((n, n), dtype=complex)
    sp[np.arange(n - 1), np.arange(1, n)] = off
    sx = (sp + sp.conj().T) / 2
    sy = (sp - sp.conj().T) / 2j
    return sx, sy, sz
