"""Debug (not a pytest file): do two half-batch engines on two streams overlap K1/K3 with K2?"""
import sys, time, threading
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 3)[0])
import bench  # noqa: E402
import tnsu_b200 as tnsu  # noqa: E402

B = 9472
K = 10
smat, tensors, weights = bench.seeded_networks(B)
ham, gate = bench.afh_operators()

def make(lo, hi):
    n = hi - lo
    eng = tnsu.Engine(smat, 2, [8] * 8, 8, False, batch=n)
    eng.reserve_gate_slots(1)
    eng.set_gates(0, np.broadcast_to(gate, (n, 8, 4, 4)))
    for t in range(4):
        eng.set_tensor(t, tensors[t][lo:hi])
    for e in range(8):
        eng.set_weights(e, weights[lo:hi])
    eng.sweep(0, 3); eng.synchronize()
    return eng

def timed(engs):
    def work(e):
        e.sweep(0, K); e.synchronize()
    ths = [threading.Thread(target=work, args=(e,)) for e in engs]
    t0 = time.perf_counter()
    [t.start() for t in ths]; [t.join() for t in ths]
    return time.perf_counter() - t0

one = make(0, B)
t1 = timed([one]); t1 = timed([one])
print(f"1 engine  x {B}: {1e3*t1/K:.2f} ms/sweep  {B*8*K/t1/1e6:.3f} M edge-updates/s")
one.close()
for parts in (2, 4):
    bounds = np.linspace(0, B, parts + 1).astype(int)
    engs = [make(int(bounds[i]), int(bounds[i+1])) for i in range(parts)]
    t = timed(engs); t = timed(engs)
    print(f"{parts} engines x {B//parts}: {1e3*t/K:.2f} ms/sweep  {B*8*K/t/1e6:.3f} M edge-updates/s")
    [e.close() for e in engs]
