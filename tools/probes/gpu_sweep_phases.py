"""Debug (not a pytest file): host / device phases of the 256-point TFI sweep (bench.py Metric 2)."""
import sys, time
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 3)[0])
import bench  # noqa: F401,E402
import tnsu_b200 as tnsu  # noqa: E402

n = 256
smat = tnsu.smc.infinite_structure_matrix_dict("peps")
pz, px = np.array([[1.0, 0], [0, -1.0]]), np.array([[0, 1.0], [1.0, 0]])
fields = -np.linspace(0.0, 4.0, n)
for rep in range(2):
    nets = []
    for k in range(n):
        np.random.seed(k)
        nets.append(tnsu.TensorNetwork(structure_matrix=smat, virtual_dim=4, spin_dim=2))
    T = [time.perf_counter()]
    batch = tnsu.SimpleUpdateBatch(nets, [-1.0] * 8, tnsu.per_point([float(f) for f in fields]), [pz], [pz], [px],
                                   [0.1, 0.01, 0.001, 0.0001, 0.00001], d_max=4, max_iterations=200, convergence_error=1e-6)
    T.append(time.perf_counter())
    eng = batch._ensure_engine(); T.append(time.perf_counter())
    batch._set_schedule_gates(); T.append(time.perf_counter())
    is_last = [int(batch.dts[0][s] == batch.dts[0][-1]) for s in range(5)]
    res = eng.run(is_last, 200, 1e-6); T.append(time.perf_counter())
    batch.download_state(); T.append(time.perf_counter())
    e = batch.energy_per_site(); T.append(time.perf_counter())
    mx = batch.expectation_per_site(px); T.append(time.perf_counter())
    mz = batch.expectation_per_site(pz); T.append(time.perf_counter())
    names = ["ctor", "ensure_engine(hams, create, upload)", "gates(expm)+upload", "engine.run", "download_state", "energy", "Mx", "Mz"]
    print(f"rep {rep}: total {T[-1]-T[0]:.3f} s | " + " | ".join(f"{nm} {1e3*(b-a):.1f} ms" for nm, a, b in zip(names, T[:-1], T[1:])))
