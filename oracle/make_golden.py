"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/src/tnsu).

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference tree does not exist on the
GPU box):   OMP_NUM_THREADS=1 python oracle/make_golden.py

The reference is imported as-is; its two missing third-party imports are satisfied by
`oracle/_stubs` (an einsum `ncon`, an empty `matplotlib`).  For every case the script also runs the
restatement in `oracle/tnsu_oracle.py` on the same seed and stores the largest deviation it saw in
the fixture (`oracle_dev_*`), so `tests/test_oracle.py` can assert the pinning without the
reference being present.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402

import tnsu.simple_update as ref_su  # noqa: E402
import tnsu.structure_matrix_constructor as ref_smc  # noqa: E402
from tnsu.math_objects import spin_operators as ref_spin_operators  # noqa: E402
from tnsu.tensor_network import TensorNetwork as RefTensorNetwork  # noqa: E402

import tnsu_oracle as orc  # noqa: E402
from cases import CASES, build_case  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _pack_list(prefix, arrays, store):
    for n, a in enumerate(arrays):
        store[f"{prefix}_{n}"] = np.asarray(a)
    store[f"{prefix}_count"] = np.array(len(arrays))


def sweep_case(name):
    """n sweeps of the reference from a seeded random network; record invariants."""
    case = build_case(name)
    smat, d, dim0, d_max, dt = case["smat"], case["d"], case["virtual_dim"], case["d_max"], case["dt"]
    np.random.seed(case["seed"])
    tn = RefTensorNetwork(structure_matrix=smat, tensors=None, weights=None, spin_dim=d, virtual_dim=dim0)
    if case.get("complex_init"):
        tn.tensors = [t + 1j * np.random.normal(size=t.shape) for t in tn.tensors]
    tensors0 = [np.array(t) for t in tn.tensors]
    sim = ref_su.SimpleUpdate(
        tn, case["j_ij"], case["h_k"], case["s_i"], case["s_j"], case["s_k"], [dt], d_max=d_max,
        print_process=False, hamiltonian=case["hamiltonian"],
    )
    sim.dt = dt
    model = orc.Model(case["j_ij"], case["h_k"], case["s_i"], case["s_j"], case["s_k"], case["hamiltonian"])
    o_t = [np.array(t) for t in tensors0]
    o_w = [np.array(w) for w in tn.weights]

    store = {}
    dev_w = dev_e = dev_rdm = 0.0
    energies = []
    for sw in range(case["sweeps"]):
        sim.simple_update()
        orc.sweep(o_t, o_w, smat, model, dt, d_max)
        e_ref = sim.energy_per_site()
        e_orc = orc.energy_per_site(o_t, o_w, smat, model)
        energies.append(e_ref)
        dev_e = max(dev_e, abs(e_ref - e_orc))
        dev_w = max(dev_w, max(np.abs(a - b).max() for a, b in zip(tn.weights, o_w)))
        _pack_list(f"weights_s{sw}", tn.weights, store)
    m = smat.shape[1]
    rdm_edges = sorted(set([0, m // 2, m - 1]))
    for e in rdm_edges:
        r_ref = sim.tensor_pair_rdm(e)
        dev_rdm = max(dev_rdm, np.abs(r_ref - orc.pair_rdm(o_t, o_w, smat, e)).max())
        store[f"pair_rdm_{e}"] = r_ref
    for t in sorted(set([0, smat.shape[0] - 1])):
        r_ref = sim.tensor_rdm(t)
        dev_rdm = max(dev_rdm, np.abs(r_ref - orc.site_rdm(o_t, o_w, smat, t)).max())
        store[f"site_rdm_{t}"] = r_ref
    op1 = case["s_i"][-1]
    store["expect1"] = np.array(sim.expectation_per_site(op1))
    # two-site observable family (simple_update.py:605-613, :639-649) and the one-site expectation (:595-603)
    op2 = np.reshape(np.kron(case["s_i"][-1], case["s_j"][0]), (d, d, d, d))
    store["pair_expect_per_site"] = np.array(sim.pair_expectation_per_site(op2))
    store["pair_expect_edge"] = np.array([sim.tensor_pair_expectation(e, op2) for e in rdm_edges])
    store["site_expect_0"] = np.array(sim.tensor_expectation(0, op1))
    dev_e = max(dev_e, abs(float(store["pair_expect_per_site"]) - orc.pair_expectation_per_site(o_t, o_w, smat, op2)))
    store["energies"] = np.array(energies)
    store["abs_tensor_0"] = np.abs(tn.tensors[0])
    store["oracle_dev_weights"] = np.array(dev_w)
    store["oracle_dev_energy"] = np.array(dev_e)
    store["oracle_dev_rdm"] = np.array(dev_rdm)
    np.savez_compressed(os.path.join(OUT, f"sweep_{name}.npz"), **store)
    return {"energy": energies[-1], "dev_w": dev_w, "dev_e": dev_e, "dev_rdm": dev_rdm}


def run_case(name):
    """Whole SimpleUpdate.run() of the reference; record logger, final energy and weights."""
    case = build_case(name)
    smat, d = case["smat"], case["d"]
    np.random.seed(case["seed"])
    tn = RefTensorNetwork(structure_matrix=smat, tensors=None, weights=None, spin_dim=d,
                          virtual_dim=case["virtual_dim"])
    if case.get("complex_init"):
        tn.tensors = [t + 1j * np.random.normal(size=t.shape) for t in tn.tensors]
    sim = ref_su.SimpleUpdate(
        tn, case["j_ij"], case["h_k"], case["s_i"], case["s_j"], case["s_k"], case["dts"],
        d_max=case["d_max"], max_iterations=case["max_iterations"], convergence_error=case["tol"],
        log_energy=True, print_process=False, hamiltonian=case["hamiltonian"],
    )
    with contextlib.redirect_stdout(io.StringIO()):
        sim.run()
    e_ref = sim.energy_per_site()

    np.random.seed(case["seed"])
    o_t, o_w = orc.random_network(smat, d, case["virtual_dim"])
    if case.get("complex_init"):
        o_t = [t + 1j * np.random.normal(size=t.shape) for t in o_t]
    model = orc.Model(case["j_ij"], case["h_k"], case["s_i"], case["s_j"], case["s_k"], case["hamiltonian"])
    log = orc.run(o_t, o_w, smat, model, case["dts"], case["d_max"], case["max_iterations"], case["tol"])
    e_orc = orc.energy_per_site(o_t, o_w, smat, model)

    store = {
        "energy": np.array(e_ref),
        "converged": np.array(sim.converged),
        "log_error": np.array(sim.logger["error"]),
        "log_energy": np.array(sim.logger["energy"]),
        "log_dt": np.array(sim.logger["dt"]),
        "log_iteration": np.array(sim.logger["iteration"]),
        "expect1": np.array(sim.expectation_per_site(case["s_i"][-1])),
        "oracle_dev_energy": np.array(abs(e_ref - e_orc)),
        "oracle_dev_checks": np.array(abs(len(log["error"]) - len(sim.logger["error"]))),
        "oracle_dev_weights": np.array(max(np.abs(a - b).max() for a, b in zip(tn.weights, o_w))),
        "min_weight": np.array(min(float(np.min(w)) for w in tn.weights)),
        # largest reference-vs-restatement spread along the whole logged trajectory (same number of checks)
        "oracle_dev_log_energy": np.array(np.abs(np.array(log["energy"]) - np.array(sim.logger["energy"])).max()
                                          if len(log["energy"]) == len(sim.logger["energy"]) else np.inf),
        "oracle_dev_log_error": np.array(np.abs(np.array(log["error"]) - np.array(sim.logger["error"])).max()
                                         if len(log["error"]) == len(sim.logger["error"]) else np.inf),
    }
    _pack_list("weights", tn.weights, store)
    np.savez_compressed(os.path.join(OUT, f"run_{name}.npz"), **store)
    return {"energy": e_ref, "checks": len(sim.logger["error"]), "converged": bool(sim.converged),
            "dev_e": float(store["oracle_dev_energy"]), "dev_w": float(store["oracle_dev_weights"]),
            "dev_log_e": float(store["oracle_dev_log_energy"]), "dev_log_err": float(store["oracle_dev_log_error"])}


def pickled_state():
    """The reference's own known-answer fixture (tests/test_main.py:17-35): re-export the pickled
    10x10 OBC D=4 AFH state as npz (data, not source) and record the reference's energy for it."""
    from tnsu.examples import load_a_tensor_network_from_memory

    smat = ref_smc.rectangular_peps_obc(height=10, width=10)
    with contextlib.redirect_stdout(io.StringIO()):
        tn = load_a_tensor_network_from_memory(structure_matrix=smat, network_name="AFH_10x10_obc_D_4")
    sx, sy, sz = ref_spin_operators(0.5)
    m = smat.shape[1]
    sim = ref_su.SimpleUpdate(tensor_network=tn, dts=[], j_ij=[1.0] * m, h_k=0.0, s_i=[sx, sy, sz],
                              s_j=[sx, sy, sz], s_k=[sx])
    e_ref = sim.energy_per_site()
    model = orc.Model([1.0] * m, 0.0, [sx, sy, sz], [sx, sy, sz], [sx])
    e_orc = orc.energy_per_site(tn.tensors, tn.weights, smat, model)
    store = {"energy_ref_live": np.array(e_ref), "energy_known_answer": np.array(-0.5616206254235453),
             "oracle_dev_energy": np.array(abs(e_ref - e_orc)), "smat": smat,
             "log_error_tail": np.array(tn.su_logger["error"][-5:]),
             "log_energy_tail": np.array(tn.su_logger["energy"][-5:])}
    _pack_list("tensors", tn.tensors, store)
    _pack_list("weights", tn.weights, store)
    np.savez_compressed(os.path.join(OUT, "pkl_afh_10x10_obc_D4.npz"), **store)
    # the shipped pickle itself (data, 778 kB): the input of the load_network known-answer test on the GPU box
    import shutil
    os.makedirs(os.path.join(OUT, "networks"), exist_ok=True)
    shutil.copyfile("/root/reference/src/tnsu/networks/AFH_10x10_obc_D_4.pkl",
                    os.path.join(OUT, "networks", "AFH_10x10_obc_D_4.pkl"))
    return {"energy": e_ref, "dev_e": abs(e_ref - e_orc), "vs_known": abs(e_ref + 0.5616206254235453)}


def smat_fixtures():
    """Structure matrices of the reference constructors, to pin the re-derived ones in the package."""
    store = {}
    for nm in ["chain", "kagome_3", "peps", "star", "cube", "triangle", "kagome_12", "pyrochlore"]:
        store[f"inf_{nm}"] = ref_smc.infinite_structure_matrix_dict(nm)
    for side in (2, 3, 4):
        store[f"pbc_{side}"] = ref_smc.square_peps_pbc(side)
    for h, w in ((2, 2), (3, 4), (10, 10), (16, 16), (1, 5)):
        store[f"obc_{h}x{w}"] = ref_smc.rectangular_peps_obc(h, w)
    for s in (0.5, 1.0, 1.5, 2.0):
        sx, sy, sz = ref_spin_operators(s)
        store[f"spin_{s}"] = np.stack([sx, sy, sz])
    np.savez_compressed(os.path.join(OUT, "structure_matrices.npz"), **store)


def main():
    os.makedirs(OUT, exist_ok=True)
    manifest = {"numpy": np.__version__, "scipy": __import__("scipy").__version__, "cases": {}}
    mpath = os.path.join(OUT, "manifest.json")
    if len(sys.argv) > 1 and os.path.exists(mpath):  # a subset: keep the other cases' entries
        manifest["cases"] = json.load(open(mpath))["cases"]
    else:
        smat_fixtures()
        manifest["cases"]["pkl_afh_10x10_obc_D4"] = pickled_state()
        print("pkl", manifest["cases"]["pkl_afh_10x10_obc_D4"])
    for name, case in CASES.items():
        kind = case["kind"]
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        res = sweep_case(name) if kind == "sweep" else run_case(name)
        manifest["cases"][name] = res
        print(kind, name, res, flush=True)
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, default=float)


if __name__ == "__main__":
    main()
