"""Debug (not a pytest file): do uploads of one engine overlap sweeps of another?"""
import sys, time, threading
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 3)[0])
import bench  # noqa: E402
import torch  # noqa: E402
import tnsu_b200 as tnsu  # noqa: E402

B = 2048
smat, tensors, weights = bench.seeded_networks(B)
ham, gate = bench.afh_operators()
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
host_t = [pin(t) for t in tensors]
engs = []
for k in range(2):
    eng = tnsu.Engine(smat, 2, [8] * 8, 8, False, batch=B)
    eng.reserve_gate_slots(1)
    eng.set_gates(0, np.broadcast_to(gate, (B, 8, 4, 4)))
    eng.set_hamiltonians(np.broadcast_to(ham, (B, 8, 4, 4)))
    for t in range(4):
        eng.set_tensor(t, host_t[t])
    for e in range(8):
        eng.set_weights(e, weights)
    engs.append(eng)

def upload(eng, n, out):
    t0 = time.perf_counter()
    for _ in range(n):
        for t in range(4):
            eng.set_tensor(t, host_t[t])
    out.append(("upload", time.perf_counter() - t0))

def sweep(eng, n, out):
    t0 = time.perf_counter()
    for _ in range(n):
        eng.sweep(0, 1); eng.synchronize()
    out.append(("sweep", time.perf_counter() - t0))

for name, jobs in (("upload alone", [(upload, engs[0])]), ("sweep alone", [(sweep, engs[1])]),
                   ("both", [(upload, engs[0]), (sweep, engs[1])]), ("two sweeps", [(sweep, engs[0]), (sweep, engs[1])]),
                   ("two uploads", [(upload, engs[0]), (upload, engs[1])])):
    out = []
    ths = [threading.Thread(target=f, args=(e, 10, out)) for f, e in jobs]
    t0 = time.perf_counter()
    [t.start() for t in ths]; [t.join() for t in ths]
    print(name, f"wall {1e3*(time.perf_counter()-t0):.1f} ms", [(k, round(1e3*v,1)) for k, v in out])
