"""What the host link gives, next to what the e2e leg of bench.py gets out of it (GPU box only).

  python tools/probes/gpu_pcie_probe.py [batch]

Prints the pinned-memory copy rates (up alone, down alone, both at once, 1-D and strided) with torch streams, then the
wall time of each C-ABI transfer call of one e2e chunk (set_tensor / set_weights / get_tensor / get_weights).
"""
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tensor-networks-simple-update_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))

import tnsu_b200 as tnsu  # noqa: E402
import bench  # noqa: E402


def rate(nbytes, fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
    n = 620 << 20  # one chunk of the default e2e leg, in bytes
    h_up = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_dn = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def up():
        with torch.cuda.stream(s_up):
            d_a.copy_(h_up, non_blocking=True)

    def down():
        with torch.cuda.stream(s_dn):
            h_dn.copy_(d_b, non_blocking=True)

    def both():
        up()
        down()

    print(f"pinned copies of {n >> 20} MiB: up {rate(n, up):.1f} GB/s, down {rate(n, down):.1f} GB/s, "
          f"both at once {rate(2 * n, both):.1f} GB/s (sum of the two directions)")

    smat, tensors, weights = bench.seeded_networks(batch)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    host_t = [pin(t) for t in tensors]
    host_w = pin(weights)
    out_t = [pin(np.empty_like(t)) for t in host_t]
    out_w = [pin(np.empty((batch, bench.DIM))) for _ in range(bench.N_E)]
    ham, gate = bench.afh_operators()
    eng = tnsu.Engine(smat, bench.D, [bench.DIM] * bench.N_E, bench.DIM, False, batch=batch)
    eng.reserve_gate_slots(1)
    eng.set_gates(0, np.broadcast_to(gate, (batch, bench.N_E, 4, 4)))

    def timed(label, fn, nbytes, reps=3):
        fn()
        eng.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        eng.synchronize()
        dt = (time.perf_counter() - t0) / reps
        print(f"  {label:<34s} {dt * 1e3:8.3f} ms  {nbytes / dt / 1e9:7.2f} GB/s")
        return dt

    tb = sum(t.nbytes for t in host_t)
    wb = bench.N_E * host_w.nbytes
    print(f"one chunk of {batch} networks: tensors {tb >> 20} MiB, weights {wb >> 10} KiB")
    t_up_t = timed("set_tensor x 4", lambda: [eng.set_tensor(t, host_t[t]) for t in range(bench.N_T)], tb)
    t_up_w = timed("set_weights x 8", lambda: [eng.set_weights(e, host_w) for e in range(bench.N_E)], wb)
    t_sw = timed("sweep + sweep_error", lambda: (eng.sweep(0, 1), eng.sweep_error()), 0)
    t_dn_w = timed("get_weights x 8", lambda: [eng.get_weights(e, out=out_w[e]) for e in range(bench.N_E)], wb)
    t_dn_t = timed("get_tensor x 4", lambda: [eng.get_tensor(t, out=out_t[t]) for t in range(bench.N_T)], tb)
    print(f"  serial sum {1e3 * (t_up_t + t_up_w + t_sw + t_dn_w + t_dn_t):.2f} ms; up {1e3 * (t_up_t + t_up_w):.2f}, "
          f"gpu {1e3 * t_sw:.2f}, down {1e3 * (t_dn_w + t_dn_t):.2f}")

    # two engines: one uploads while the other downloads (what the e2e pipeline overlaps)
    eng2 = tnsu.Engine(smat, bench.D, [bench.DIM] * bench.N_E, bench.DIM, False, batch=batch)
    for t in range(bench.N_T):
        eng2.set_tensor(t, host_t[t])

    def updown():
        th = threading.Thread(target=lambda: [eng2.get_tensor(t, out=out_t[t]) for t in range(bench.N_T)])
        th.start()
        for t in range(bench.N_T):
            eng.set_tensor(t, host_t[t])
        th.join()
        eng2.synchronize()

    timed("set_tensor x4 || get_tensor x4", updown, 2 * tb)


if __name__ == "__main__":
    main()
