"""Debug (not a pytest file): per-edge pair expectation of the pickled 10x10 state, GPU vs oracle."""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 3)[0] + "/tests")
from conftest import load_golden, unpack_list  # noqa: E402

import tnsu_b200 as tnsu  # noqa: E402
import tnsu_oracle as orc  # noqa: E402

g = load_golden("pkl_afh_10x10_obc_D4")
tensors, weights = unpack_list(g, "tensors"), unpack_list(g, "weights")
smat = tnsu.smc.rectangular_peps_obc(10, 10)
sx, sy, sz = tnsu.spin_operators(0.5)
model = orc.Model([1.0] * smat.shape[1], 0.0, [sx, sy, sz], [sx, sy, sz], [sx])
for d_max in (2, 4):
    tn = tnsu.TensorNetwork(structure_matrix=smat, tensors=[t.copy() for t in tensors], weights=[w.copy() for w in weights])
    sim = tnsu.SimpleUpdate(tn, [1.0] * smat.shape[1], 0.0, [sx, sy, sz], [sx, sy, sz], [sx], dts=[], d_max=d_max)
    e = sim.energy_per_site()
    print(f"d_max={d_max}: E {e:.16f} dev {abs(e + 0.5616206254235453):.2e}")
    bad = 0
    for ek in range(smat.shape[1]):
        r_gpu = sim.tensor_pair_rdm(ek)
        r_cpu = orc.pair_rdm(tensors, weights, smat, ek)
        dev = np.abs(r_gpu - r_cpu).max()
        if dev > 1e-10:
            bad += 1
            if bad < 8:
                print(f"  edge {ek}: rdm dev {dev:.2e} tensors {np.nonzero(smat[:, ek])[0]}")
    print(f"  {bad} bad edges of {smat.shape[1]}")
