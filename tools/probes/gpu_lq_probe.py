"""Debug (not a pytest file): per-phase cycles of the rolled K1 (build with TNSU_NVCC_EXTRA=-DLQR_PROFILE)."""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 3)[0])
import bench  # noqa: E402
import tnsu_b200 as tnsu  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
smat, tensors, weights = bench.seeded_networks(B)
ham, gate = bench.afh_operators()
eng = tnsu.Engine(smat, 2, [8] * 8, 8, False, batch=B)
eng.reserve_gate_slots(1)
eng.set_gates(0, np.broadcast_to(gate, (B, 8, 4, 4)))
eng.set_hamiltonians(np.broadcast_to(ham, (B, 8, 4, 4)))
for t in range(4):
    eng.set_tensor(t, tensors[t])
for e in range(8):
    eng.set_weights(e, weights)
eng.sweep(0, 2)
for edge in (0, 3):
    dbg = eng.debug_edge(0, edge, 0)
    c = dbg["lq_cycles"]
    print(f"B={B} edge {edge}: K1 cycles over 16 steps: dots+store {c[0]}  warp-sum {c[1]}  barrier {c[2]}  xwarp+scalar {c[3]}  update+ystore {c[4]}  total {sum(c)} ({sum(c)/16:.0f}/step)")
