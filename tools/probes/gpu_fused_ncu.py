"""Debug (not a pytest file): a short on-chip run (256 TFI points, 10 iterations per dt) for ncu."""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 3)[0] + "/tensor-networks-simple-update_b200")
import tnsu_b200 as tnsu  # noqa: E402

n_points = int(sys.argv[1]) if len(sys.argv) > 1 else 256
smat = tnsu.smc.infinite_structure_matrix_dict("peps")
pz, px = np.array([[1.0, 0], [0, -1.0]]), np.array([[0, 1.0], [1.0, 0]])
fields = -np.linspace(0.0, 4.0, n_points)
nets = []
for k in range(n_points):
    np.random.seed(k)
    nets.append(tnsu.TensorNetwork(structure_matrix=smat, virtual_dim=4, spin_dim=2))
batch = tnsu.SimpleUpdateBatch(nets, [-1.0] * 8, tnsu.per_point([float(h) for h in fields]), [pz], [pz], [px],
                               [0.1, 0.01, 0.001, 0.0001, 0.00001], d_max=4, max_iterations=10, convergence_error=1e-6)
logs = batch.run()
print("sweeps", sum(lg["sweeps"] for lg in logs), batch.engine.stats())
