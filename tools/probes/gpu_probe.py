"""Verbose GPU parity probe (not a pytest file): prints the deviation of every golden case so one gpurun
call shows everything at once.  python tools/probes/gpu_probe.py [case ...]"""
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 3)[0] + "/tests")
from conftest import build_case, load_golden, unpack_list  # noqa: E402

import cases  # noqa: E402
import device_model as dm  # noqa: E402
import tnsu_b200 as tnsu  # noqa: E402
import tnsu_oracle as orc  # noqa: E402


def make_network(case):
    np.random.seed(case["seed"])
    tn = tnsu.TensorNetwork(structure_matrix=case["smat"], spin_dim=case["d"], virtual_dim=case["virtual_dim"])
    if case.get("complex_init"):
        tn.tensors = [t + 1j * np.random.normal(size=t.shape) for t in tn.tensors]
    return tn


def probe_sweep(name):
    case = build_case(name)
    g = load_golden("sweep_" + name)
    tn = make_network(case)
    sim = tnsu.SimpleUpdate(tn, case["j_ij"], case["h_k"], case["s_i"], case["s_j"], case["s_k"], [case["dt"]],
                            d_max=case["d_max"], print_process=False, hamiltonian=case["hamiltonian"])
    sim.dt = case["dt"]
    t0 = time.time()
    worst_w = 0.0
    for sw in range(case["sweeps"]):
        sim.simple_update()
        ref_w = unpack_list(g, f"weights_s{sw}")
        shapes_ok = all(len(a) == len(b) for a, b in zip(ref_w, tn.weights))
        dw = max(np.abs(a - b).max() for a, b in zip(ref_w, tn.weights)) if shapes_ok else float("nan")
        e = sim.energy_per_site()
        de = abs(e - g["energies"][sw])
        worst_w = max(worst_w, dw)
        print(f"  sweep {sw}: dW {dw:.2e} dE {de:.2e} E {e:.12f} ref {g['energies'][sw]:.12f}", flush=True)
    m = case["smat"].shape[1]
    for e in sorted(set([0, m // 2, m - 1])):
        print(f"  pair_rdm[{e}] dev {np.abs(sim.tensor_pair_rdm(e) - g[f'pair_rdm_{e}']).max():.2e}")
    for t in sorted(set([0, case['smat'].shape[0] - 1])):
        print(f"  site_rdm[{t}] dev {np.abs(sim.tensor_rdm(t) - g[f'site_rdm_{t}']).max():.2e}")
    print(f"  expect1 dev {abs(sim.expectation_per_site(case['s_i'][-1]) - float(g['expect1'])):.2e}")
    print(f"  |T0| dev {np.abs(np.abs(tn.tensors[0]) - g['abs_tensor_0']).max():.2e}   ({time.time() - t0:.1f}s)")
    return worst_w


def probe_run(name):
    case = build_case(name)
    g = load_golden("run_" + name)
    tn = make_network(case)
    sim = tnsu.SimpleUpdate(tn, case["j_ij"], case["h_k"], case["s_i"], case["s_j"], case["s_k"], case["dts"],
                            d_max=case["d_max"], max_iterations=case["max_iterations"], convergence_error=case["tol"],
                            log_energy=True, print_process=False, hamiltonian=case["hamiltonian"])
    t0 = time.time()
    sim.run()
    dt = time.time() - t0
    e = sim.energy_per_site()
    ref_w = unpack_list(g, "weights")
    dw = max(np.abs(a - b).max() for a, b in zip(ref_w, tn.weights))
    nref = len(g["log_error"])
    nmine = len(sim.logger["error"])
    k = min(nref, nmine)
    dlog = np.abs(np.array(sim.logger["energy"][:k]) - g["log_energy"][:k]).max() if k else 0.0
    derr = np.abs(np.array(sim.logger["error"][:k]) - g["log_error"][:k]).max() if k else 0.0
    print(f"  E {e:.13f} ref {float(g['energy']):.13f} dE {abs(e - float(g['energy'])):.2e} dW {dw:.2e} checks {nmine}/{nref} "
          f"conv {sim.converged}/{bool(g['converged'])} max|dlogE| {dlog:.2e} max|dlogErr| {derr:.2e}  ({dt:.1f}s)")


def probe_pkl():
    g = load_golden("pkl_afh_10x10_obc_D4")
    tensors, weights = unpack_list(g, "tensors"), unpack_list(g, "weights")
    smat = tnsu.smc.rectangular_peps_obc(10, 10)
    tn = tnsu.TensorNetwork(structure_matrix=smat, tensors=tensors, weights=weights)
    sx, sy, sz = tnsu.spin_operators(0.5)
    sim = tnsu.SimpleUpdate(tn, [1.0] * smat.shape[1], 0.0, [sx, sy, sz], [sx, sy, sz], [sx], dts=[])
    e = sim.energy_per_site()
    print(f"  pkl energy {e:.16f} known {float(g['energy_known_answer']):.16f} dev {abs(e - float(g['energy_known_answer'])):.2e}")


def debug_first_edge(name):
    """Compare kernel intermediates of edge 0 with the numpy device model."""
    case = build_case(name)
    tn = make_network(case)
    model = orc.Model(case["j_ij"], case["h_k"], case["s_i"], case["s_j"], case["s_k"], case["hamiltonian"])
    sim = tnsu.SimpleUpdate(tn, case["j_ij"], case["h_k"], case["s_i"], case["s_j"], case["s_k"], [case["dt"]],
                            d_max=case["d_max"], print_process=False, hamiltonian=case["hamiltonian"])
    impl = sim._impl()
    eng = impl._ensure_engine()
    gates = impl._gates(impl._all_hamiltonians(), [case["dt"]])
    eng.set_gates(1, gates)
    dbg = eng.debug_edge(1, 0, 0)
    smat = case["smat"]
    ti, ai, tj, aj = orc.edge_endpoints(smat, 0)
    T = [np.array(t) for t in tn.tensors]
    W = [np.array(w) for w in tn.weights]
    outs = []
    for t in (ti, tj):
        p, _ = dm.bond_matrix(T[t], W, smat, t, 0)
        outs.append(dm.householder_lq_inplace(p)[0])
    print(f"  dbg L_i dev {np.abs(dbg['l_i'] - outs[0]).max():.2e}  L_j dev {np.abs(dbg['l_j'] - outs[1]).max():.2e} "
          f"svd sweeps {dbg['svd_sweeps']} d_new {dbg['d_new']}")
    d = case["d"]
    g = gates[0, 0].reshape(d, d, d, d)
    dk = len(W[0])
    theta = np.einsum("ieq,e,jep,ijab->qabp", outs[0].reshape(d, dk, -1), W[0], outs[1].reshape(d, dk, -1), g)
    s = np.linalg.svd(theta.reshape(theta.shape[0] * d, -1), compute_uv=False)
    print(f"  dbg sigma dev {np.abs(dbg['sigma'] - s).max():.2e}")


if __name__ == "__main__":
    names = sys.argv[1:] or list(cases.CASES)
    print("debug edge 0:")
    for nm in [n for n in names if cases.CASES.get(n, {}).get("kind") == "sweep" and "growth" not in n and n not in
               ("pyro_field",)][:6]:
        print(nm)
        try:
            debug_first_edge(nm)
        except Exception:  # noqa: BLE001
            traceback.print_exc()
    for nm in names:
        kind = cases.CASES[nm]["kind"] if nm in cases.CASES else nm
        print(f"== {nm} ({kind})", flush=True)
        try:
            if nm == "pkl":
                probe_pkl()
            elif kind == "sweep":
                probe_sweep(nm)
            else:
                probe_run(nm)
        except Exception:  # noqa: BLE001
            traceback.print_exc()
    if not sys.argv[1:]:
        print("== pkl")
        try:
            probe_pkl()
        except Exception:  # noqa: BLE001
            traceback.print_exc()
