"""Per-phase cycle counts of one edge update (debug path of the kernel), not a pytest file."""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 3)[0] + "/tests")
from conftest import build_case  # noqa: E402

import tnsu_b200 as tnsu  # noqa: E402

for name, cplx in (("peps_D8", False), ("peps_D8", True), ("tfi_D4", False), ("blbq_D4", False), ("obc16_D6", False),
                   ("star_D2", False)):
    case = build_case(name)
    np.random.seed(case["seed"])
    tn = tnsu.TensorNetwork(structure_matrix=case["smat"], spin_dim=case["d"], virtual_dim=case["virtual_dim"])
    if cplx:
        tn.tensors = [t + 1e-3j * np.random.normal(size=t.shape) for t in tn.tensors]
    sim = tnsu.SimpleUpdate(tn, case["j_ij"], case["h_k"], case["s_i"], case["s_j"], case["s_k"], [case["dt"]],
                            d_max=case["d_max"], print_process=False, hamiltonian=case["hamiltonian"])
    impl = sim._impl()
    eng = impl._ensure_engine()
    eng.set_gates(1, impl._gates(impl._all_hamiltonians(), [case["dt"]]))
    edge = case["smat"].shape[1] // 2
    for rep in range(2):
        dbg = eng.debug_edge(1, edge, 0)
    pc = dbg["phase_cycles"]
    eng.sweep(1, 2)
    prof = eng.profile_sweep(1)
    print(f"{name:10s} cplx={int(cplx)} edge {edge} svd_sweeps {dbg['svd_sweeps']:2d} K2 cycles " +
          "  ".join(f"{k}={v}" for k, v in pc.items()) + "   sweep ms (B=1): " +
          "  ".join(f"{k}={v:.3f}" for k, v in prof.items()), flush=True)
