"""Debug (not a pytest file): where the time of a LARGE TFI sweep (Metric 2, `sweep_large` of bench.py) goes: host-side
construction, gates (scipy expm), upload, run() on the device, download, observables.
    python tools/probes/gpu_sweep_large_phases.py [n_points]"""
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 3)[0])
sys.path.insert(0, __file__.rsplit("/", 3)[0] + "/tensor-networks-simple-update_b200")
import tnsu_b200 as tnsu  # noqa: E402

n_points = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
smat = tnsu.smc.infinite_structure_matrix_dict("peps")
pz, px = np.array([[1.0, 0], [0, -1.0]]), np.array([[0, 1.0], [1.0, 0]])
fields = -np.linspace(0.0, 4.0, n_points)
if len(sys.argv) > 2 and sys.argv[2] == "shuffled":  # points in random order: finished points are scattered over the batch
    fields = np.random.default_rng(1).permutation(fields)
nets = []
for k in range(n_points):
    np.random.seed(k)
    nets.append(tnsu.TensorNetwork(structure_matrix=smat, virtual_dim=4, spin_dim=2))
t0 = time.perf_counter()
batch = tnsu.SimpleUpdateBatch(nets, [-1.0] * 8, tnsu.per_point([float(h) for h in fields]), [pz], [pz], [px],
                               [0.1, 0.01, 0.001, 0.0001, 0.00001], d_max=4, max_iterations=200, convergence_error=1e-6)
t1 = time.perf_counter()
import cProfile, pstats
pr0 = cProfile.Profile()
pr0.enable()
eng = batch._ensure_engine()
t2 = time.perf_counter()
batch._set_schedule_gates()
pr0.disable()
t3 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
logs = batch.run()
pr.disable()
t4 = time.perf_counter()
en = batch.energy_per_site()
mx = batch.expectation_per_site(px)
t5 = time.perf_counter()
sweeps = np.array([lg["sweeps"] for lg in logs])
print(f"{n_points} points: ctor {t1 - t0:.3f} s, engine + upload {t2 - t1:.3f}, gates {t3 - t2:.3f}, run() {t4 - t3:.3f}, "
      f"energy + Mx {t5 - t4:.3f}; total {t5 - t0:.3f} s; sweeps max {sweeps.max()} mean {sweeps.mean():.0f}")
if len(sys.argv) > 3:
    pstats.Stats(pr0).sort_stats("cumulative").print_stats(22)
pstats.Stats(pr).sort_stats("cumulative").print_stats(3)
