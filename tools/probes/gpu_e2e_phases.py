"""Debug (not a pytest file): wall time of the phases of one end-to-end step of bench.py."""
import sys, time
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 3)[0])
import bench  # noqa: E402
import torch  # noqa: E402
import tnsu_b200 as tnsu  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
smat, tensors, weights = bench.seeded_networks(B)
ham, gate = bench.afh_operators()
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
host_t = [pin(t) for t in tensors]
host_w = pin(weights)
eng = tnsu.Engine(smat, 2, [8] * 8, 8, False, batch=B)
eng.reserve_gate_slots(1)
eng.set_gates(0, np.broadcast_to(gate, (B, 8, 4, 4)))
eng.set_hamiltonians(np.broadcast_to(ham, (B, 8, 4, 4)))
for rep in range(3):
    t0 = time.perf_counter()
    for t in range(4):
        eng.set_tensor(t, host_t[t])
    t1 = time.perf_counter()
    for e in range(8):
        eng.set_weights(e, host_w)
    t2 = time.perf_counter()
    eng.sweep(0, 1); eng.synchronize()
    t3 = time.perf_counter()
    en = eng.energy()
    t4 = time.perf_counter()
    ws = [eng.get_weights(e) for e in range(8)]
    t5 = time.perf_counter()
    nb = sum(t.nbytes for t in host_t)
    print(f"B={B} set_tensor {1e3*(t1-t0):.1f} ms ({nb/(t1-t0)/1e9:.1f} GB/s)  set_weights {1e3*(t2-t1):.1f}  sweep {1e3*(t3-t2):.1f}  energy {1e3*(t4-t3):.1f}  get_weights {1e3*(t5-t4):.1f}")
# raw torch copy for comparison
x = torch.from_numpy(host_t[0]); d = torch.empty_like(x, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(x, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"torch H2D {x.numel()*8/(t1-t0)/1e9:.1f} GB/s")
