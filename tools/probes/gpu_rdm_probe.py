"""Debug (not a pytest file): time energy_per_site (pair-RDM kernel) for the DMMA and the staged kernel.
usage: TNSU_RDM=simple python tools/probes/gpu_rdm_probe.py"""
import os
import sys
import time
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 3)[0])
import bench  # noqa: E402
import tnsu_b200 as tnsu  # noqa: E402

CASES = (("peps", 2, 8, 4096), ("peps", 2, 4, 16384), ("triangle", 3, 4, 512), ("star", 2, 2, 65536))
if len(sys.argv) > 1:
    CASES = CASES[:int(sys.argv[1])]
for lattice, d, dim, B in CASES:
    smat = tnsu.smc.infinite_structure_matrix_dict(lattice)
    n, m = smat.shape
    rng = np.random.default_rng(0)
    eng = tnsu.Engine(smat, d, [dim] * m, dim, False, batch=B)
    sx, sy, sz = tnsu.spin_operators((d - 1) / 2)
    ham = sum(np.kron(s, s) for s in (sx, sy, sz)).astype(complex)
    eng.set_hamiltonians(np.broadcast_to(ham, (B, m, d * d, d * d)))
    nbytes = 0
    for t in range(n):
        shp = eng.tensor_shape(t)
        a = rng.normal(1.0, 1.0, (B,) + tuple(shp))
        nbytes += a.nbytes
        eng.set_tensor(t, a)
    for e in range(m):
        eng.set_weights(e, np.sort(rng.uniform(0.1, 1.0, (B, dim)), axis=1)[:, ::-1].copy())
    en = eng.energy()
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        en = eng.energy()
    dt = (time.perf_counter() - t0) / reps
    print(f"{os.environ.get('TNSU_RDM', 'dmma'):7s} {lattice:9s} d={d} D={dim} B={B}: {dt*1e3:8.3f} ms per energy batch, "
          f"{B/dt/1e6:7.3f} M eval/s, minimal {nbytes/dt/1e9:7.1f} GB/s, e[0]={en[0]:.12f} e[-1]={en[-1]:.12f}", flush=True)
    eng.close()
