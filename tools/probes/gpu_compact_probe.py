"""Debug (not a pytest file): tnsu_run with and without the compaction of running points — where do the logs first differ?"""
import os
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 3)[0] + "/tensor-networks-simple-update_b200")
import tnsu_b200 as tnsu  # noqa: E402

pz, px = np.array([[1, 0], [0, -1]]), np.array([[0, 1], [1, 0]])
smat = tnsu.smc.infinite_structure_matrix_dict("peps")
fields = [-0.2 - 0.1 * k for k in range(40)]
dim0, dmax = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2, 3)
os.environ["TNSU_FUSED"] = "0"
out = {}
for mode in ("1", "0"):
    os.environ["TNSU_COMPACT"] = mode
    nets = []
    for k in range(len(fields)):
        np.random.seed(500 + k)
        nets.append(tnsu.TensorNetwork(structure_matrix=smat, virtual_dim=dim0, spin_dim=2))
    b = tnsu.SimpleUpdateBatch(nets, [-1.0] * 8, tnsu.per_point(fields), [pz], [pz], [px], [0.1, 0.01], d_max=dmax,
                               max_iterations=30, convergence_error=1e-6)
    logs = b.run()
    out[mode] = logs
a, c = out["1"], out["0"]
print("sweeps per point:", [lg["sweeps"] for lg in a])
for p, (la, lc) in enumerate(zip(a, c)):
    ea, ec = np.array(la["error"]), np.array(lc["error"])
    d = np.nonzero(ea != ec)[0]
    if len(d):
        print(f"point {p}: first differing check {d[0]} (sweep {la['iteration'][d[0]]}), rel diff {abs(ea[d[0]] - ec[d[0]]) / abs(ec[d[0]]):.2e}, {len(d)} of {len(ea)} checks differ")
