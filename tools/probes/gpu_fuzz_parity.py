"""Debug (not a pytest file): randomized parity sweep — random lattice / bond dimension / spin / field / time step, a few
sweeps on the GPU against the oracle on the same seeded input; prints the worst deviations.
    python tools/probes/gpu_fuzz_parity.py [n_cases] [seed]"""
import sys
import time

import numpy as np

ROOT = __file__.rsplit("/", 3)[0]
sys.path.insert(0, ROOT + "/tensor-networks-simple-update_b200")
sys.path.insert(0, ROOT + "/oracle")
import tnsu_b200 as tnsu  # noqa: E402
import tnsu_oracle as orc  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
lattices = {"chain": (2, 12), "star": (2, 8), "peps": (2, 8), "triangle": (2, 3), "cube": (2, 3)}
worst_w = worst_e = 0.0
t0 = time.time()
for case in range(n_cases):
    name = list(lattices)[rng.integers(len(lattices))]
    lo, hi = lattices[name]
    d = int(rng.choice([2, 2, 3]))
    hi = max(lo, min(hi, 32 // d))
    dim0 = int(rng.integers(lo, hi + 1))
    d_max = int(rng.integers(dim0, min(hi, dim0 + 3) + 1))
    smat = tnsu.smc.infinite_structure_matrix_dict(name)
    m = smat.shape[1]
    s = (d - 1) / 2.0
    sx, sy, sz = tnsu.spin_operators(s)
    use_sy = bool(rng.integers(4) == 0)
    cplx_init = bool(rng.integers(5) == 0)
    field = float(rng.normal())
    dt = float(rng.choice([0.1, 0.03, 0.005]))
    s_k = [sy] if use_sy else [sx]
    j = [float(x) for x in rng.choice([-1.0, 1.0, 0.5], size=m)]
    np.random.seed(1000 + case)
    tn = tnsu.TensorNetwork(structure_matrix=smat, virtual_dim=dim0, spin_dim=d)
    if cplx_init:
        tn.tensors = [t + 1j * np.random.normal(size=t.shape) for t in tn.tensors]
    t_ref, w_ref = [np.array(t) for t in tn.tensors], [np.array(w) for w in tn.weights]
    su = tnsu.SimpleUpdate(tn, j, field, [sx, sy, sz], [sx, sy, sz], s_k, [dt], d_max=d_max, print_process=False)
    su.dt = dt
    model = orc.Model(j, field, [sx, sy, sz], [sx, sy, sz], s_k)
    n_sweeps = 3
    try:
        for _ in range(n_sweeps):
            su.simple_update()
            orc.sweep(t_ref, w_ref, smat, model, dt, d_max)
        dw = max(np.abs(a - b).max() for a, b in zip(w_ref, tn.weights))
        e_ref = orc.energy_per_site(t_ref, w_ref, smat, model)
        de = abs(su.energy_per_site() - e_ref) / max(1.0, abs(e_ref))
        worst_w, worst_e = max(worst_w, dw), max(worst_e, de)
        flag = "" if (dw < 1e-10 and de < 1e-10) else "   <-- ABOVE 1e-10"
        print(f"{case:3d} {name:9s} d={d} D {dim0}->{d_max} field {field:+.2f} {'Sy' if use_sy else 'Sx'} dt {dt} "
              f"{'complex init' if cplx_init else 'real init   '}: max|dW| {dw:.1e}  |dE| {de:.1e}{flag}", flush=True)
    except tnsu.EngineError as exc:
        print(f"{case:3d} {name:9s} d={d} D {dim0}->{d_max}: engine refused: {str(exc)[:80]}", flush=True)
print(f"worst max|dW| {worst_w:.2e}  worst |dE| {worst_e:.2e}  ({time.time() - t0:.0f} s)")
