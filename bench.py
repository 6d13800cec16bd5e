#!/usr/bin/env python
"""bench.py — SU edge-updates/s at D=8 on the square-lattice ("peps" cell) spin-1/2 AFH model.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

One "step" = one Simple-Update sweep (8 edge updates: absorb, QR reduction, gate, truncated SVD,
renormalise) over a batch of B independent random-init networks of BASELINE.json configs[1]
(d=2, virtual_dim = d_max = 8, j_ij=1, h_k=0, dt=0.1, seed b for network b).  B (default 9472 per GPU) is chosen
so the state (B x 4 tensors x 8192 float64 = B x 256 KiB = 2.2 GiB) is far larger than the 126 MB L2.

  value  edge-updates/s with the state resident in HBM (CUDA events on the engine's stream, max over ranks)
  e2e    the same metric through the C ABI with host buffers: every step uploads the B networks from pinned host
         memory, runs the sweep and reads EVERYTHING it produced back (tensors, weights, convergence error)
  value_converged / converged   the same sweeps on converged states at the last time step, with the share of bond
         matrices the streaming reduction had to hand to the Householder kernels
  c128   the complex128 engine at the same shape;  e2e_run   whole SimpleUpdateBatch.run() calls, host objects to host objects
  sweep / sweep_large   Metric 2: the 256-point TFI sweep and a saturating size (strong scaling, cost-sorted placement)
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md "Measurement"

--impl reference times the reference's algorithm on the host cores (the numpy/scipy oracle port; the
reference itself is pure Python and cannot travel to the GPU box), one process per core.
"""
import os

os.environ.setdefault("OMP_NUM_THREADS", "1")  # the reference's operating point is one BLAS thread
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")

import argparse  # noqa: E402
import json  # noqa: E402
import subprocess  # noqa: E402
import sys  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tensor-networks-simple-update_b200"))

METRIC = "SU edge-updates/s at D=8 square AFH"
UNIT = "edge-updates/s"
D, DIM, N_T, N_E = 2, 8, 4, 8  # physical dim, bond dim, tensors, edges of the "peps" cell
WORKLOAD = "configs[1]: 2D square-lattice iPEPS ('peps' cell) spin-1/2 AFH, d=2, virtual_dim=d_max=8, dt=0.1"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:  # noqa: BLE001
        return {"hbm_gbs": 6650.0}, "fallback"


def algorithmic_bytes_per_edge(scalar_bytes):
    """read both site tensors once, write both once, read the weight vectors of all legs, write
    one: 2*s*d*(D^z_i + D^z_j) + 8*(sum of leg dims)."""
    return 2 * scalar_bytes * D * (DIM ** 4 + DIM ** 4) + 8 * (8 * DIM)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms (one streaming process) while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.thread = threading.Thread(target=self._loop, daemon=True)

    def _loop(self):
        try:
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.strip().split(",")])
        except Exception:  # noqa: BLE001
            pass

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread.start()
            time.sleep(0.35)  # let the first samples arrive before the timed region starts
            self.first = len(self.rows)
        except Exception:  # noqa: BLE001
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.12)
            self.proc.terminate()
            self.thread.join(timeout=3)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[max(getattr(self, "first", 1) - 1, 0):]:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:  # noqa: BLE001
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(torch, local_rank):
    """Pin this rank's threads (and with them the first-touch placement of its pinned host buffers) to the CPUs NVML lists
    as local to its GPU: with 8 ranks the uploads of one rank otherwise cross the socket interconnect.  Instrumentation
    only; returns the number of CPUs bound to, or None when NVML / the affinity call is not available."""
    try:
        import pynvml

        pynvml.nvmlInit()
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{dom:08x}:{bus:02x}:{dev:02x}.0")
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:  # noqa: BLE001 - best effort
        pass
    return None


def link_rates(torch, barrier, nbytes=512 << 20, reps=3):
    """Pinned-memory copy rates of THIS rank while every rank copies at the same time (GB/s: up alone, down alone, both
    directions together): what the host side can feed, next to what the e2e leg gets out of it."""
    h_up = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_dn = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def run(up, down):
        def once():
            if up:
                with torch.cuda.stream(s_up):
                    d_a.copy_(h_up, non_blocking=True)
            if down:
                with torch.cuda.stream(s_dn):
                    h_dn.copy_(d_b, non_blocking=True)

        once()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        return (int(up) + int(down)) * nbytes * reps / dt / 1e9

    return run(True, False), run(False, True), run(True, True)


def seeded_networks(batch, first_seed=0):
    """Host state exactly as TensorNetwork.__init__ draws it (legacy global numpy stream, seed b for network b)."""
    import tnsu_b200 as tnsu

    smat = tnsu.smc.infinite_structure_matrix_dict("peps")
    tensors = [np.empty((batch, D) + (DIM,) * 4) for _ in range(N_T)]
    for b in range(batch):
        np.random.seed(first_seed + b)
        tn = tnsu.TensorNetwork(structure_matrix=smat, spin_dim=D, virtual_dim=DIM)
        for t in range(N_T):
            tensors[t][b] = tn.tensors[t]
    weights = np.full((batch, DIM), 1.0 / DIM)
    return smat, tensors, weights


def afh_operators():
    import tnsu_b200 as tnsu
    from scipy import linalg

    sx, sy, sz = tnsu.spin_operators(0.5)
    ham = sum(np.kron(s, s) for s in (sx, sy, sz)).astype(complex)  # j_ij = 1, h_k = 0
    return ham, linalg.expm(-0.1 * ham)


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle port of the reference algorithm)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, budget_s, min_sweeps = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tensor-networks-simple-update_b200"))
    import tnsu_oracle as orc
    from tnsu_b200 import smc

    smat = smc.infinite_structure_matrix_dict("peps")
    sx, sy, sz = orc.spin_operators(0.5)
    model = orc.Model([1.0] * N_E, 0.0, [sx, sy, sz], [sx, sy, sz], [])
    np.random.seed(seed)
    tensors, weights = orc.random_network(smat, D, DIM)
    orc.sweep(tensors, weights, smat, model, 0.1, DIM)  # warm-up
    t0 = time.perf_counter()
    sweeps = 0
    while sweeps < min_sweeps or time.perf_counter() - t0 < budget_s:
        orc.sweep(tensors, weights, smat, model, 0.1, DIM)
        sweeps += 1
    return sweeps * N_E, time.perf_counter() - t0


def cpu_baseline(cores, budget_s, min_sweeps=2, pool=None):
    """edge-updates/s of the oracle on `cores` host processes running concurrently: the SUM of the per-process rates
    (edge updates of a process / its own timed loop), so process start-up and imports are never in the number."""
    import multiprocessing as mp

    if cores == 1:
        res = [_cpu_worker((0, budget_s, min_sweeps))]
    elif pool is not None:
        res = pool.map(_cpu_worker, [(s, budget_s, min_sweeps) for s in range(cores)])
    else:
        with mp.get_context("spawn").Pool(cores) as own:
            res = own.map(_cpu_worker, [(s, budget_s, min_sweeps) for s in range(cores)])
    return sum(n / t for n, t in res), sum(n for n, _ in res)


def run_reference_arm(args):
    import multiprocessing as mp

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step_s = 4.0
    rates, edges = [], 0
    with mp.get_context("spawn").Pool(cores) as pool:  # one pool for the whole arm: workers are warm before step 1
        for _ in range(max(args.warmup, 1)):
            cpu_baseline(cores, 0.5, 1, pool)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            rate, n = cpu_baseline(cores, per_step_s, 2, pool)
            rates.append(rate)
            edges += n
        dt = time.perf_counter() - t0
    value = float(np.mean(rates))
    sample = (f"{args.steps} steps x {cores} processes x >= {per_step_s:.0f} s of sweeps on one peps D=8 network each; value = mean over "
              f"steps of the summed per-process rates ({edges} edge updates in {dt:.1f} s of wall clock incl. dispatch)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": {"workload": WORKLOAD, "host": "numpy/scipy oracle port of the reference algorithm, 1 BLAS thread per process"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
# Metric 2: converged sweep points/s — BASELINE.json configs[2], the transverse-field Ising sweep
# ------------------------------------------------------------------------------------------------
def tfi_sweep(n_points, world, rank, local_rank, dist, torch, costs=None):
    """256-point TFI sweep on the "peps" cell (SURVEY.md 8d Metric 2): j_ij=-1, h_k=-linspace(0,4,n), d_max=4,
    dts=[1e-1..1e-5], 200 iterations per dt, tol 1e-6, fresh random network per point (seed = point index);
    full run() + energy + Mx + Mz per point.  Points are sharded round-robin over the ranks; the only collective is
    the final all_gather of [energy, Mx, Mz, converged, sweeps] per point.  Strong scaling (total work fixed)."""
    import tnsu_b200 as tnsu
    from tnsu_b200 import sharding

    smat = tnsu.smc.infinite_structure_matrix_dict("peps")
    pz, px = np.array([[1.0, 0], [0, -1.0]]), np.array([[0, 1.0], [1.0, 0]])
    fields = -np.linspace(0.0, 4.0, n_points)
    mine = sharding.partition(n_points, world, rank, costs)  # costs given: cost-sorted placement, longest points first
    nets = []
    for k in mine:
        np.random.seed(int(k))
        nets.append(tnsu.TensorNetwork(structure_matrix=smat, virtual_dim=4, spin_dim=2))
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    batch = tnsu.SimpleUpdateBatch(nets, [-1.0] * N_E, tnsu.per_point([float(fields[k]) for k in mine]), [pz], [pz], [px],
                                   [0.1, 0.01, 0.001, 0.0001, 0.00001], d_max=4, max_iterations=200,
                                   convergence_error=1e-6, device=local_rank)
    logs = batch.run()
    launches = batch.engine.launch_count
    energy = batch.energy_per_site()
    mx = batch.expectation_per_site(px)
    mz = batch.expectation_per_site(pz)
    local = np.stack([energy, mx, mz, logs.converged.astype(float), logs.sweeps.astype(float)], axis=1)
    full = sharding.gather_points(local, n_points, world, rank, device=torch.device("cuda", local_rank), costs=costs)
    fused = batch.engine.stats()["fused_runs"]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.cpu()[0])
    return {"metric": f"converged sweep points/s ({n_points}-point TFI sweep, peps cell, D=4)", "value": n_points / dt,
            "unit": "points/s", "seconds": dt, "n_points": n_points, "scaling": "strong",
            "placement": "round-robin" if costs is None else "cost-sorted (predicted sweeps interpolated from the 256-point sweep)",
            "on_chip_run_kernel": bool(fused), "sweeps_per_point": full[:, 4].copy(),
            "max_sweeps_of_a_point": int(full[:, 4].max()),
            "converged_points": int(full[:, 3].sum()), "total_sweeps": int(full[:, 4].sum()),
            "edge_updates_per_s": float(full[:, 4].sum()) * N_E / dt, "gpu_launches_rank0": int(launches),
            "energy_h0": float(full[0, 0]), "energy_h4": float(full[-1, 0]), "mx_h4": float(full[-1, 1]),
            "what": "TensorNetwork construction excluded; gates (scipy expm), upload, run(), observables, gather included"}


def tfi_sweep_cpu(n_sample, n_points):
    """Oracle run() on `n_sample` of the sweep's points, one host core: seconds per point."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tnsu_oracle as orc
    from tnsu_b200 import smc

    smat = smc.infinite_structure_matrix_dict("peps")
    pz, px = np.array([[1.0, 0], [0, -1.0]]), np.array([[0, 1.0], [1.0, 0]])
    fields = -np.linspace(0.0, 4.0, n_points)
    picks = np.linspace(0, n_points - 1, n_sample).astype(int)
    t0 = time.perf_counter()
    for k in picks:
        np.random.seed(int(k))
        tensors, weights = orc.random_network(smat, D, 4)
        orc.run(tensors, weights, smat, orc.Model([-1.0] * N_E, float(fields[k]), [pz], [pz], [px]),
                [0.1, 0.01, 0.001, 0.0001, 0.00001], 4, 200, 1e-6)
    return (time.perf_counter() - t0) / n_sample, [int(k) for k in picks]


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
DTS = [0.1, 0.01, 0.001, 0.0001, 0.00001]


def ncu_evidence(name):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:  # noqa: BLE001
        return {}


def resident_sweeps(eng, torch, stream, barrier, slot, steps, warmup, sampler=None):
    """K timed sweeps with the state resident in HBM: CUDA events on the engine's stream, barrier + synchronize on both
    sides.  Returns (ms of the K sweeps, ms measured by the library around its own launches, launches, clock summary)."""
    eng.sweep(slot, warmup)
    eng.synchronize()
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    clocks = None
    if sampler is not None:
        sampler.__enter__()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        eng.sweep(slot, steps)
        ev1.record(stream)
    barrier()
    if sampler is not None:
        sampler.__exit__()
        clocks = sampler.summary()
    kern_ms, n_launch = eng.last_sweep_kernel_ms()
    return ev0.elapsed_time(ev1), kern_ms, n_launch, eng.launch_count - launches0, clocks


def run_gpu_arm(args):
    import torch

    import tnsu_b200 as tnsu
    from scipy import linalg

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    numa_cpus = bind_to_gpu_numa_node(torch, local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    batch = args.batch
    cplx = args.dtype == "c128"
    sbytes = 16 if cplx else 8
    stream = torch.cuda.Stream()
    smat, tensors, weights = seeded_networks(batch, first_seed=rank * batch)
    ham, gate = afh_operators()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731

    def complexify(ts):
        """genuinely complex data for the complex128 engine: the seeded real draw plus i x (the same draw, rolled)"""
        return [(t + 1j * np.roll(t, 1, axis=1)).astype(np.complex128) for t in ts]

    if cplx:
        tensors = complexify(tensors)
    host_t = [pin(t) for t in tensors]
    host_w = pin(weights)
    out_w = [pin(np.empty((batch, DIM))) for _ in range(N_E)]
    out_t = [pin(np.empty_like(t)) for t in host_t]

    eng = tnsu.Engine(smat, D, [DIM] * N_E, DIM, cplx, batch=batch, device=local_rank, stream=stream.cuda_stream)
    eng.reserve_gate_slots(2)
    eng.set_gates(0, np.broadcast_to(gate, (batch, N_E, 4, 4)))
    eng.set_gates(1, np.broadcast_to(linalg.expm(-DTS[-1] * ham), (batch, N_E, 4, 4)))  # the last time step of a run
    eng.set_hamiltonians(np.broadcast_to(ham, (batch, N_E, 4, 4)))

    def upload(e, ts, ws):
        for t in range(N_T):
            e.set_tensor(t, ts[t])
        for ed in range(N_E):
            e.set_weights(ed, ws if isinstance(ws, np.ndarray) else ws[ed])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    upload(eng, host_t, host_w)
    # ---- resident-state timing: K sweeps from the random start (sweeps W+1 .. W+K at dt = 0.1) ----
    ms, kern_ms, n_launch, gpu_launches, clocks = resident_sweeps(eng, torch, stream, barrier, 0, args.steps, args.warmup,
                                                                   ClockSampler(local_rank))
    n_levels = int(np.max(eng.edge_levels)) + 1
    prof = eng.profile_sweep(0)  # one extra sweep with events between the kernel classes (not part of `value`)

    # ---- energy_per_site over the resident batch: the bandwidth-bound pair-RDM contraction (reference :525-637) ----
    eng.energy()
    eng.energy()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ee0.record(stream)
    for _ in range(args.energy_steps):
        en_vals = eng.energy()  # launches pair_rdm + pair_sum, copies B energies to the host, synchronises
    ee1.record(stream)
    barrier()
    energy_ms = ee0.elapsed_time(ee1) / args.energy_steps
    assert np.all(np.isfinite(en_vals))

    # ---- the same sweep on CONVERGED states: run() of n_conv networks to the end of the dt schedule (the general
    # kernels: these D = 8 networks do not fit the on-chip run kernel), tiled over the batch, then K sweeps at the
    # last time step.  Small weights are what make bond matrices ill-conditioned: the share that the streaming
    # CholeskyQR2 flags for the Householder kernels is reported with the number ----
    converged = None
    if args.converged_networks > 0 and not cplx:
        n_conv = min(args.converged_networks, batch)
        ce = tnsu.Engine(smat, D, [DIM] * N_E, DIM, False, batch=n_conv, device=local_rank)
        ce.reserve_gate_slots(len(DTS))
        for sl, dt in enumerate(DTS):
            ce.set_gates(sl, np.broadcast_to(linalg.expm(-dt * ham), (n_conv, N_E, 4, 4)))
        ce.set_hamiltonians(np.broadcast_to(ham, (n_conv, N_E, 4, 4)))
        upload(ce, [t[:n_conv] for t in host_t], host_w[:n_conv])
        t0 = time.perf_counter()
        res = ce.run([int(dt == DTS[-1]) for dt in DTS], 200, 1e-6)
        run_s = time.perf_counter() - t0
        reps = (batch + n_conv - 1) // n_conv
        conv_t = [np.tile(ce.get_tensor(t), (reps,) + (1,) * 5)[:batch] for t in range(N_T)]
        conv_w = [np.tile(ce.get_weights(ed), (reps, 1))[:batch] for ed in range(N_E)]
        min_w = float(min(w.min() for w in conv_w))
        conv_energy = float(np.mean(ce.energy()))
        ce.close()
        upload(eng, conv_t, conv_w)
        st0 = eng.stats()
        cms, ckern_ms, _, _, _ = resident_sweeps(eng, torch, stream, barrier, 1, args.steps, args.warmup)
        st1 = eng.stats()
        cprof = eng.profile_sweep(1)
        streamed = st1["lq_streamed"] - st0["lq_streamed"]
        converged = {"ms": cms, "kernel_ms_per_sweep": cprof, "lq_streamed": streamed,
                     "lq_flagged": st1["lq_flagged"] - st0["lq_flagged"], "min_weight": min_w, "n_networks": n_conv,
                     "run_seconds": run_s, "run_sweeps_mean": float(np.mean(res["sweeps"])),
                     "run_converged": int(np.sum(res["converged"])), "energy_per_site_mean": conv_energy}
        del conv_t, conv_w
    del eng

    # ---- complex128 (the reference's own arithmetic: every tensor is complex128 after its first update, :435) at the
    # same shape, its own batch: a sub-object of the default line ----
    c128 = None
    if args.c128_batch > 0 and not cplx:
        cb = min(args.c128_batch, batch)
        c_t = complexify([t[:cb] for t in tensors])
        ceng = tnsu.Engine(smat, D, [DIM] * N_E, DIM, True, batch=cb, device=local_rank, stream=stream.cuda_stream)
        ceng.reserve_gate_slots(1)
        ceng.set_gates(0, np.broadcast_to(gate, (cb, N_E, 4, 4)))
        ceng.set_hamiltonians(np.broadcast_to(ham, (cb, N_E, 4, 4)))
        upload(ceng, c_t, weights[:cb])
        zms, zkern_ms, zn_launch, _, _ = resident_sweeps(ceng, torch, stream, barrier, 0, args.steps, args.warmup)
        zprof = ceng.profile_sweep(0)
        zst = ceng.stats()
        c128 = {"ms": zms, "kern_ms": zkern_ms, "batch": cb, "prof": zprof, "stats": zst}
        del ceng, c_t

    # ---- end to end through the API with host buffers ----
    # The batch is split over `e2e_chunks` engines (handles are independent: one stream each) driven by one host
    # thread each, so the pinned-memory upload of one chunk overlaps the sweep of another.  A step's result is what
    # the reference's simple_update() leaves behind: the updated tensors and weights (:294-296) and the convergence
    # error — all of it is read back every step.
    n_ch = max(1, min(args.e2e_chunks, batch))
    bounds = np.linspace(0, batch, n_ch + 1).astype(int)
    parts = []
    for c in range(n_ch):
        lo, hi = int(bounds[c]), int(bounds[c + 1])
        pe = tnsu.Engine(smat, D, [DIM] * N_E, DIM, cplx, batch=hi - lo, device=local_rank)
        pe.reserve_gate_slots(1)
        pe.set_gates(0, np.broadcast_to(gate, (hi - lo, N_E, 4, 4)))
        pe.set_hamiltonians(np.broadcast_to(ham, (hi - lo, N_E, 4, 4)))
        parts.append((pe, lo, hi))
    out_e = pin(np.empty(batch))

    # one lock per shared resource (PCIe up, SM time, PCIe down): without them the chunk threads march in lockstep
    up_lock, gpu_lock, down_lock = threading.Lock(), threading.Lock(), threading.Lock()

    def e2e_chunk(pe, lo, hi, n_steps):
        for _ in range(n_steps):
            with up_lock:
                for t in range(N_T):
                    pe.set_tensor(t, host_t[t][lo:hi])
                for e in range(N_E):
                    pe.set_weights(e, host_w[lo:hi])
            with gpu_lock:
                pe.sweep(0, 1)
                out_e[lo:hi] = pe.sweep_error()
            with down_lock:
                for e in range(N_E):
                    pe.get_weights(e, out=out_w[e][lo:hi])  # straight into the pinned result buffers
                for t in range(N_T):
                    pe.get_tensor(t, out=out_t[t][lo:hi])

    def e2e_run(n_steps):
        threads = [threading.Thread(target=e2e_chunk, args=(pe, lo, hi, n_steps)) for pe, lo, hi in parts]
        for th in threads:
            th.start()
        for th in threads:
            th.join()

    e2e_run(min(args.warmup, 2))
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.e2e_steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    link = link_rates(torch, barrier)  # all ranks copy at the same time, like in the e2e leg above
    h2d = sum(t.nbytes for t in host_t) + N_E * host_w.nbytes
    d2h = out_e.nbytes + sum(w.nbytes for w in out_w) + sum(t.nbytes for t in out_t)
    assert all(abs(np.linalg.norm(out_t[0][b]) - 1.0) < 1e-9 for b in (0, batch - 1))  # the sweep's tensors did come back
    for pe, _, _ in parts:
        pe.close()

    # ---- end to end, whole runs: SimpleUpdateBatch.run() from host TensorNetwork objects to host objects (gates by scipy
    # expm, upload, the whole dt schedule on the device, download of every tensor / weight / log) ----
    e2e_whole = None
    if args.e2e_run_networks > 0 and not cplx:
        n_run = args.e2e_run_networks
        sx, sy, sz = tnsu.spin_operators(0.5)
        nets = []
        for b in range(n_run):
            np.random.seed(rank * n_run + b)
            nets.append(tnsu.TensorNetwork(structure_matrix=smat, spin_dim=D, virtual_dim=DIM))
        barrier()
        t0 = time.perf_counter()
        sb = tnsu.SimpleUpdateBatch(nets, [1.0] * N_E, 0.0, [sx, sy, sz], [sx, sy, sz], [], DTS, d_max=DIM, max_iterations=200,
                                    convergence_error=1e-6, device=local_rank)
        logs = sb.run()
        run_energy = sb.energy_per_site()
        barrier()
        whole_s = time.perf_counter() - t0
        e2e_whole = {"seconds": whole_s, "edges": int(logs.sweeps.sum()) * N_E, "n": n_run,
                     "energy": float(np.mean(run_energy)), "converged": int(logs.converged.sum())}
        sb.engine.close()

    # ---- max over ranks ----
    vals = [ms, e2e_s * 1e3, energy_ms, converged["ms"] if converged else 0.0, c128["ms"] if c128 else 0.0,
            e2e_whole["seconds"] * 1e3 if e2e_whole else 0.0]
    times = torch.tensor(vals, dtype=torch.float64, device="cuda")
    link_sum = torch.tensor(link, dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(link_sum, op=dist.ReduceOp.SUM)
    ms, e2e_ms, energy_ms, conv_ms, c128_ms, whole_ms = (float(x) for x in times.cpu())
    link_up, link_down, link_both = (float(x) for x in link_sum.cpu())
    edges_per_step = world * batch * N_E
    value = edges_per_step * args.steps / (ms * 1e-3)
    e2e_value = edges_per_step * args.e2e_steps / (e2e_ms * 1e-3)

    sweep = sweep_large = None
    if args.sweep_points > 0:
        sweep = tfi_sweep(args.sweep_points, world, rank, local_rank, dist, torch)
        if args.sweep_large_points > 0:
            from tnsu_b200 import sharding

            costs = sharding.predict_costs(np.linspace(0.0, 4.0, args.sweep_points), sweep["sweeps_per_point"],
                                           np.linspace(0.0, 4.0, args.sweep_large_points))
            sweep_large = tfi_sweep(args.sweep_large_points, world, rank, local_rank, dist, torch, costs=costs)

    if rank == 0:
        pk, pk_kind = peaks()
        ncu = ncu_evidence("r02_ncu_edge_update.json")
        fp64_peak = 36.2  # TFLOP/s, measured on this pool (profiles/r01_microbench_b200.txt): DFMA 36.2, DMMA m8n8k4 37.0
        # roofline unit: one edge-update launch group = the kernels of one dependency level (K1c, K1 fallback, K2, K3c, K3 fallback)
        n_groups = float(n_levels * args.steps)

        def hbm_roofline(scalar_bytes, n_batch, kernel_ms, groups, traffic_per_edge):
            alg = algorithmic_bytes_per_edge(scalar_bytes) * n_batch * N_E * args.steps / groups
            ach = alg / (kernel_ms * 1e-3 / groups) / 1e9
            return alg, ach, (traffic_per_edge * n_batch * N_E * args.steps / groups if traffic_per_edge else None)

        alg_launch, achieved, traffic = hbm_roofline(sbytes, batch, kern_ms, n_groups,
                                                     None if cplx else ncu.get("dram_bytes_per_edge_update"))
        fma_per_edge = 6.0e6 if cplx else 1.5e6
        cores = 1
        cpu_val, cpu_edges = cpu_baseline(cores, args.cpu_seconds)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "c128" if cplx else "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": batch, "edges_per_step": edges_per_step,
                       "state_bytes_per_gpu": batch * N_T * D * DIM ** 4 * sbytes,
                       "l2_policy": "inputs larger than L2 (state >> 126 MB), no flush",
                       "arithmetic": "real float64 path (inputs with zero imaginary part)" if not cplx else "complex128 path",
                       "parallelism": f"points sharded over {world} GPU(s), no data-path collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": args.e2e_steps, "chunks": n_ch,
                    "link_gbs": (h2d + d2h) * world * args.e2e_steps / (e2e_ms * 1e-3) / 1e9,
                    "host_link": {"up_gbs": link_up, "down_gbs": link_down, "both_directions_gbs": link_both, "unit": "GB/s",
                                  "numa_bound_cpus_rank0": numa_cpus,
                                  "what": "pinned-memory copies of 512 MiB per direction, all ranks at the same time, summed over "
                                          "the ranks: the ceiling of `link_gbs`, the bytes this leg moves per second in both directions"},
                    "what": "per step: upload all networks from pinned host memory + 1 sweep + download EVERYTHING the sweep produced "
                            "(all tensors, all weights, the convergence error), through the C ABI; the batch is split over `chunks` "
                            "engines driven by one host thread each so uploads, sweeps and downloads of different chunks overlap"},
            "gpu_launches": int(gpu_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk_kind,
                         "kernel": "edge update = lq_chol_warp_kernel + svd_warp_kernel + recompose_chol_kernel (one level; the Householder "
                                   "fallback kernels lq_rolled / recompose_dmma run behind them on flagged items only)",
                         "launches": int(n_groups), "kernels_per_launch_group": n_launch / n_groups,
                         "algorithmic_bytes_per_launch": alg_launch, "avg_launch_ms": kern_ms / n_groups,
                         "ncu": {k: ncu.get(k) for k in ("kernel_time_us", "fp64_pipe_pct", "dmma_pipe_pct", "dram_throughput_pct", "source")},
                         "fp64_peak_tflops_measured": fp64_peak,
                         "fp64_bound": {"fma_per_edge_update": fma_per_edge, "cap_edge_updates_per_s": fp64_peak * 1e12 / 2 / fma_per_edge,
                                        "frac": (value / world) / (fp64_peak * 1e12 / 2 / fma_per_edge),
                                        "what": "FP64 FMA count of one edge update (DESIGN.md section 4) against the measured DFMA/DMMA "
                                                "rate: the bound that actually governs this path"},
                         "note": "the edge update is FP64-pipe / latency bound, not HBM bound (DESIGN.md section 4): the SVD kernel moves "
                                 "almost no bytes; the streaming K1 reads each site tensor twice and K3 reads it again and writes it"},
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{cpu_edges} edge updates of one peps D=8 network, numpy/scipy oracle, 1 BLAS thread"},
            "clocks": clocks,
            "kernel_ms_per_sweep": prof,
        }
        if converged is not None:
            cv = world * batch * N_E * args.steps / (conv_ms * 1e-3)
            line["value_converged"] = cv
            line["converged"] = {
                "value": cv, "unit": UNIT, "ms_per_step": conv_ms / args.steps,
                "what": f"the same K sweeps at the LAST time step (dt = {DTS[-1]}) on converged states: run() of {converged['n_networks']} "
                        "seeded networks per GPU through the whole dt schedule, tiled over the batch",
                "kernel_ms_per_sweep": converged["kernel_ms_per_sweep"],
                "streaming_lq_flagged_fraction": (converged["lq_flagged"] / converged["lq_streamed"]) if converged["lq_streamed"] else None,
                "lq_streamed": converged["lq_streamed"], "lq_flagged": converged["lq_flagged"],
                "min_weight": converged["min_weight"], "energy_per_site_mean": converged["energy_per_site_mean"],
                "run": {"seconds": converged["run_seconds"], "sweeps_mean": converged["run_sweeps_mean"],
                        "converged_networks": converged["run_converged"], "networks": converged["n_networks"]}}
        if c128 is not None:
            zb = c128["batch"]
            zalg, zach, _ = hbm_roofline(16, zb, c128["kern_ms"], n_groups, None)
            zv = world * zb * N_E * args.steps / (c128_ms * 1e-3)
            line["c128"] = {
                "metric": METRIC + " (complex128 engine: complex initial tensors)", "value": zv, "unit": UNIT, "dtype": "c128",
                "batch_per_gpu": zb, "ms_per_step": c128_ms / args.steps, "kernel_ms_per_sweep": c128["prof"],
                "lq_streamed": c128["stats"]["lq_streamed"], "lq_flagged": c128["stats"]["lq_flagged"],
                "roofline": {"bound": "hbm", "achieved": zach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": zach / pk["hbm_gbs"],
                             "algorithmic_bytes_per_launch": zalg, "traffic": None,
                             "kernel": "lq_chol_warp_kernel<4, complex> + svd_kernel<cplx> + recompose_chol_kernel<4, complex>",
                             "fp64_bound": {"fma_per_edge_update": 6.0e6, "frac": (zv / world) / (fp64_peak * 1e12 / 2 / 6.0e6)}}}
        if e2e_whole is not None:
            line["e2e_run"] = {
                "value": world * e2e_whole["edges"] / (whole_ms * 1e-3), "unit": UNIT, "seconds": whole_ms * 1e-3,
                "networks_per_gpu": e2e_whole["n"], "edge_updates_rank0": e2e_whole["edges"], "converged_rank0": e2e_whole["converged"],
                "energy_per_site_mean_rank0": e2e_whole["energy"],
                "what": "SimpleUpdateBatch(...).run() + energy_per_site() from host TensorNetwork objects back to host objects: gates "
                        "(scipy expm), upload, the whole dt schedule (5 x up to 200 sweeps, convergence checks, energies) on the device, "
                        "download of all tensors / weights / logs"}
        # energy_per_site: HBM roofline on the MINIMAL bytes (every site tensor read once per evaluation);
        # the reference-shaped figure reads both tensors of every edge (4x at z = 4; the re-reads hit L2 here)
        en_min = batch * (N_T * D * DIM ** 4 * sbytes + N_E * DIM * 8 + 8)
        en_ref = batch * N_E * (2 * D * DIM ** 4 * sbytes + 7 * DIM * 8)
        en_gbs = en_min / (energy_ms * 1e-3) / 1e9
        mb = (D * DIM + 7) // 8
        dmma_flop = batch * N_E * 2 * (mb * (mb + 1) // 2) * (DIM ** 3 // 4) * 512  # executed m8n8k4 DMMAs x 2*8*8*4
        line["energy"] = {
            "metric": "energy_per_site evaluations/s (peps D=8: 8 pair RDMs + <H> each), state resident in HBM",
            "value": world * batch / (energy_ms * 1e-3), "unit": "evaluations/s", "ms_per_batch": energy_ms,
            "pair_rdms_per_s": world * batch * N_E / (energy_ms * 1e-3), "steps": args.energy_steps,
            "roofline": {"bound": "hbm", "achieved": en_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": en_gbs / pk["hbm_gbs"], "algorithmic_bytes_per_launch": en_min,
                         "reference_shaped_bytes_per_launch": en_ref,
                         "reference_shaped_gbs": en_ref / (energy_ms * 1e-3) / 1e9,
                         "kernel": "pair_rdm_dmma_kernel<2> (+ pair_sum_kernel, 71 KB D2H)" if not cplx else "pair_rdm_kernel<cplx>",
                         "traffic": None if (cplx or "pair_rdm_dram_bytes_per_evaluation" not in ncu) else
                         ncu["pair_rdm_dram_bytes_per_evaluation"] * batch,
                         "dmma_tflops": None if cplx else dmma_flop / (energy_ms * 1e-3) / 1e12,
                         "dmma_peak_tflops_measured": 37.0},
        }
        for sw in (sweep, sweep_large):
            if sw is not None:
                sw.pop("sweeps_per_point", None)
        if sweep is not None:
            if args.sweep_cpu_points > 0:
                sec, picks = tfi_sweep_cpu(args.sweep_cpu_points, args.sweep_points)
                sweep["cpu_baseline"] = {"value": 1.0 / sec, "unit": "points/s", "cores": 1, "kind": "port",
                                         "sample": f"oracle run() of points {picks} of the sweep, 1 BLAS thread"}
            sweep["note"] = ("latency bound: the slowest point's chain of max_sweeps_of_a_point sweeps x 6 dependency levels x ~100 "
                             "dependent Jacobi steps bounds the sweep on any number of GPUs (DESIGN.md section 4); the saturating size "
                             "below is the strong-scaling measurement")
            line["sweep"] = sweep
        if sweep_large is not None:
            line["sweep_large"] = sweep_large
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=9472,
                    help="networks per GPU; 9472 = 4 x 148 SMs x 16 resident SVD warps, so the K2 grid is whole waves")
    ap.add_argument("--sweep-points", type=int, default=256, help="points of the TFI sweep (Metric 2); 0 skips it")
    ap.add_argument("--sweep-cpu-points", type=int, default=2, help="points of the sweep the oracle runs for the CPU baseline")
    ap.add_argument("--sweep-large-points", type=int, default=32768,
                    help="points of the saturating-size TFI sweep (strong scaling: the total is fixed, ranks share it); 0 skips it")
    ap.add_argument("--converged-networks", type=int, default=1184,
                    help="networks per GPU that run() to convergence for the converged-state headline (tiled over the batch); 0 skips it")
    ap.add_argument("--c128-batch", type=int, default=4096, help="batch of the complex128 sub-measurement; 0 skips it")
    ap.add_argument("--e2e-run-networks", type=int, default=2368, help="networks per GPU of the whole-run end-to-end measurement; 0 skips it")
    ap.add_argument("--dtype", choices=["f64", "c128"], default="f64")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--energy-steps", type=int, default=10)
    ap.add_argument("--e2e-chunks", type=int, default=4)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # exactly ONE line on stdout: libraries that write to file descriptor 1 on their own (NCCL prints its version
    # there) are sent to stderr for the duration of the run; the JSON line goes to the real stdout
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    try:
        if args.impl == "reference":
            run_reference_arm(args)
        else:
            run_gpu_arm(args)
    finally:
        real_stdout.flush()


if __name__ == "__main__":
    main()
