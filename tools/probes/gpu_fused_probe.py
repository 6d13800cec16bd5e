"""Debug (not a pytest file): the 256-point TFI sweep (Metric 2) through the on-chip persistent run kernel and through the
general launch path (TNSU_FUSED=0): seconds, points/s, identical logs?"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 3)[0])
sys.path.insert(0, __file__.rsplit("/", 3)[0] + "/tensor-networks-simple-update_b200")
import tnsu_b200 as tnsu  # noqa: E402

n_points = int(sys.argv[1]) if len(sys.argv) > 1 else 256
smat = tnsu.smc.infinite_structure_matrix_dict("peps")
pz, px = np.array([[1.0, 0], [0, -1.0]]), np.array([[0, 1.0], [1.0, 0]])
fields = -np.linspace(0.0, 4.0, n_points)
res = {}
for mode in ("1", "0", "1"):
    os.environ["TNSU_FUSED"] = mode
    nets = []
    for k in range(n_points):
        np.random.seed(k)
        nets.append(tnsu.TensorNetwork(structure_matrix=smat, virtual_dim=4, spin_dim=2))
    t0 = time.perf_counter()
    batch = tnsu.SimpleUpdateBatch(nets, [-1.0] * 8, tnsu.per_point([float(h) for h in fields]), [pz], [pz], [px],
                                   [0.1, 0.01, 0.001, 0.0001, 0.00001], d_max=4, max_iterations=200, convergence_error=1e-6)
    t1 = time.perf_counter()
    eng = batch._ensure_engine()
    batch._set_schedule_gates()
    t2 = time.perf_counter()
    logs = batch.run()
    t3 = time.perf_counter()
    en = batch.energy_per_site()
    t4 = time.perf_counter()
    st = batch.engine.stats()
    sweeps = np.array([lg["sweeps"] for lg in logs])
    print(f"TNSU_FUSED={mode}: total {t4 - t0:.3f} s (ctor {t1 - t0:.3f}, engine+gates {t2 - t1:.3f}, run {t3 - t2:.3f}, energy {t4 - t3:.3f}) "
          f"-> {n_points / (t4 - t0):.0f} points/s; sweeps max {sweeps.max()} mean {sweeps.mean():.0f}; stats {st}", flush=True)
    res[mode] = (logs, en, [np.array(w) for tn in nets for w in tn.weights])
a, b = res["1"], res["0"]
same_checks = all(la["iteration"] == lb["iteration"] and la["converged"] == lb["converged"] for la, lb in zip(a[0], b[0]))
dev_err = max(np.abs(np.array(la["error"]) - np.array(lb["error"])).max() for la, lb in zip(a[0], b[0]) if la["error"])
dev_en = max(np.abs(np.array(la["energy"]) - np.array(lb["energy"])).max() for la, lb in zip(a[0], b[0]) if la["energy"]) if same_checks else None
print("identical control flow:", same_checks, " max |d error log|", dev_err, " max |d energy log|", dev_en,
      " max |d energy|", np.abs(a[1] - b[1]).max(), " max |d weights|", max(np.abs(x - y).max() for x, y in zip(a[2], b[2])))
