"""Debug (not a pytest file): K2 phase cycles at the TFI D=4 shape (theta 16 x 16), lone warp."""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 3)[0] + "/tests")
from conftest import build_case  # noqa: E402
import tnsu_b200 as tnsu  # noqa: E402

for name in ("tfi_D4", "star_D2", "obc16_D6"):
    case = build_case(name)
    np.random.seed(case["seed"])
    tn = tnsu.TensorNetwork(structure_matrix=case["smat"], spin_dim=case["d"], virtual_dim=case["virtual_dim"])
    sim = tnsu.SimpleUpdate(tn, case["j_ij"], case["h_k"], case["s_i"], case["s_j"], case["s_k"], [case["dt"]],
                            d_max=case["d_max"], print_process=False, hamiltonian=case["hamiltonian"])
    impl = sim._impl()
    eng = impl._ensure_engine()
    eng.set_gates(1, impl._gates(impl._all_hamiltonians(), [case["dt"]]))
    eng.sweep(1, 6)
    edge = case["smat"].shape[1] // 2
    for rep in range(2):
        dbg = eng.debug_edge(1, edge, 0)
    pc = dbg["phase_cycles"]
    print(f"{name:10s} jacobi sweeps {dbg['svd_sweeps']:2d}  cycles: stage+QR {pc['theta']}  jacobi {pc['jacobi']}  extract {pc['extract']}")
