"""Debug (not a pytest file): Jacobi sweep counts and phase cycles of K2 on bench-like states."""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 3)[0])
import bench  # noqa: E402
import tnsu_b200 as tnsu  # noqa: E402

B = 64
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
for nsw in (0, 1, 3, 8, 12):
    if nsw:
        eng.sweep(0, nsw - done)
    done = nsw
    rows = []
    for pt in range(0, B, 16):
        for edge in (0, 3, 6):
            dbg = eng.debug_edge(0, edge, pt)   # NB: debug_edge UPDATES that edge of every point
            rows.append((dbg["svd_sweeps"], dbg["phase_cycles"]["theta"], dbg["phase_cycles"]["jacobi"], dbg["phase_cycles"]["extract"]))
    a = np.array(rows)
    print(f"after {nsw:2d} sweeps: jacobi sweeps mean {a[:,0].mean():.1f} min {a[:,0].min()} max {a[:,0].max()}  pre-phase cycles {a[:,1].mean():.0f}  jacobi cycles {a[:,2].mean():.0f}  extract cycles {a[:,3].mean():.0f}")
