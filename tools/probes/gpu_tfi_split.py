"""Debug (not a pytest file): K1/K2/K3 device time per sweep at the TFI D=4 shape, batch 256 and 1."""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 3)[0])
import bench  # noqa: F401,E402
import tnsu_b200 as tnsu  # noqa: E402
from scipy import linalg

smat = tnsu.smc.infinite_structure_matrix_dict("peps")
pz, px = np.array([[1.0, 0], [0, -1.0]]), np.array([[0, 1.0], [1.0, 0]])
H = (-np.kron(pz, pz) - 3.0 * (np.kron(px, np.eye(2)) + np.kron(np.eye(2), px)) / 4).astype(complex)
G = linalg.expm(-0.01 * H)
for B in (256, 4096, 32768):
    eng = tnsu.Engine(smat, 2, [4] * 8, 4, False, batch=B)
    eng.reserve_gate_slots(1)
    eng.set_gates(0, np.broadcast_to(G, (B, 8, 4, 4)))
    eng.set_hamiltonians(np.broadcast_to(H, (B, 8, 4, 4)))
    rng = np.random.default_rng(0)
    for t in range(4):
        eng.set_tensor(t, rng.normal(1, 1, size=(B, 2, 4, 4, 4, 4)))
    for e in range(8):
        eng.set_weights(e, np.full((B, 4), 0.25))
    eng.sweep(0, 20); eng.synchronize()
    prof = eng.profile_sweep(0)
    import time
    t0 = time.perf_counter(); eng.sweep(0, 50); eng.synchronize(); t1 = time.perf_counter()
    t2 = time.perf_counter(); [eng.energy() for _ in range(10)]; t3 = time.perf_counter()
    print(f"B={B}: per sweep (6 levels) lq {1e3*prof['lq']:.0f} us  svd {1e3*prof['svd']:.0f} us  rc {1e3*prof['recompose']:.0f} us | 50 sweeps back to back: {1e6*(t1-t0)/50:.0f} us/sweep | energy {1e6*(t3-t2)/10:.0f} us")
    eng.close()
