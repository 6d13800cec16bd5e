"""`SimpleUpdate` / `SimpleUpdateBatch`: the drop-in algorithm surface over the B200 engine.

`SimpleUpdate` keeps the constructor, methods, attributes, logger keys and print-outs of the
reference's `tnsu.simple_update.SimpleUpdate` (src/tnsu/simple_update.py:12-667).  What stays on
the host is exactly what must stay there for bit-parity of the inputs: the two-site Hamiltonian
(numpy kron, :420-454) and the gate expm(-dt H) (scipy, :470-473) — computed once per (edge, dt)
instead of once per edge per sweep.  Everything else — weight absorption, QR reduction, gate
application, truncated SVD, renormalisation, the RDM contractions and the run loop's convergence
logic — executes in libtnsu_b200.so on the GPU.  There is no CPU fallback.

`SimpleUpdateBatch` is the superset the reference lacks (SURVEY.md §8b "Superset allowed"): many
independent points (fields, couplings, angles, seeds) of one lattice advance in the same launches.
"""
from __future__ import annotations

from collections.abc import Sequence

import numpy as np
from scipy import linalg

from .engine import Engine
from .tensor_network import TensorNetwork

_NO_TRUNCATION = 1 << 20  # d_max=None in the reference means "keep everything"
_FINGERPRINT_SAMPLES = 1 << 16


def _fingerprint(a):
    """Cheap content stamp of an array: shape, dtype and the sum over (at most 65 536 evenly strided) elements.  The
    reference reads tensors / weights at call time, so an in-place edit (`tn.tensors[0] *= 2`,
    `tn.weights[3][:] = ...`) between two calls must reach the device: the wrapper re-uploads when a stamp changed.
    Arrays larger than the sample are strided — a single-element edit of a huge array can go unseen; replace the list
    entry or call `upload_state()` then."""
    a = np.asarray(a)
    flat = a.reshape(-1)
    step = max(1, flat.size // _FINGERPRINT_SAMPLES)
    return a.shape, a.dtype.str, complex(np.add.reduce(flat[::step]))


def _per_point(value, n_points, name):
    """Broadcast a shared parameter or validate a per-point list."""
    if isinstance(value, _PerPoint):
        if len(value.items) != n_points:
            raise ValueError(f"{name}: {len(value.items)} per-point values for {n_points} points")
        return list(value.items)
    return [value] * n_points


class _PerPoint:
    """Marks a constructor argument of SimpleUpdateBatch as 'one value per point'."""

    def __init__(self, items):
        self.items = list(items)


def per_point(items) -> _PerPoint:
    return _PerPoint(items)


class _RunLogs(Sequence):
    """The per-point logger dicts of a batch run — keys error / dt / iteration / energy (the reference's logger lists,
    :143-150) plus converged / sweeps / final_dt — built from the downloaded log arrays when a point is looked at: a
    sweep of tens of thousands of points whose caller reads two scalars per point does not pay for 4 Python lists each."""

    def __init__(self, res, dts):
        self._res, self._dts = res, dts
        self._made = [None] * len(dts)

    def __len__(self):
        return len(self._made)

    def __getitem__(self, p):
        if isinstance(p, slice):
            return [self[i] for i in range(*p.indices(len(self)))]
        if p < 0:
            p += len(self)
        if not 0 <= p < len(self):
            raise IndexError("point index out of range")
        lg = self._made[p]
        if lg is None:
            res, dts = self._res, self._dts[p]
            k = int(res["n_checks"][p])
            lg = self._made[p] = {
                "error": res["log_error"][p, :k].tolist(),
                "dt": [dts[s] for s in res["log_slot"][p, :k].tolist()],
                "iteration": res["log_step"][p, :k].tolist(),
                "energy": res["log_energy"][p, :k].tolist(),
                "converged": bool(res["converged"][p]),
                "sweeps": int(res["sweeps"][p]),
                "final_dt": dts[int(res["final_slot"][p])],
            }
        return lg

    @property
    def sweeps(self) -> np.ndarray:
        """sweeps every point ran, as one array"""
        return np.asarray(self._res["sweeps"])

    @property
    def converged(self) -> np.ndarray:
        return np.asarray(self._res["converged"]).astype(bool)


class SimpleUpdateBatch:
    """Simple Update on `len(tensor_networks)` independent networks of one structure matrix and one shape.

    j_ij, h_k, hamiltonian and dts may be shared or wrapped in `per_point([...])`; s_i, s_j, s_k are shared.
    All points use dt schedules of the same LENGTH (values may differ)."""

    def __init__(self, tensor_networks, j_ij, h_k, s_i, s_j, s_k, dts, d_max=2, max_iterations=1000,
                 convergence_error=1e-6, hamiltonian=None, device=0, stream=None, detect_inplace_edits=False):
        self.tensor_networks = list(tensor_networks)
        # content stamps per array on every call: on for the one-network drop-in facade, off by default for large
        # batches (one numpy reduction per array and call; list-entry replacement is always detected)
        self.detect_inplace_edits = detect_inplace_edits
        self.n_points = len(self.tensor_networks)
        if self.n_points == 0:
            raise ValueError("at least one tensor network is required")
        tn0 = self.tensor_networks[0]
        for tn in self.tensor_networks[1:]:
            if not np.array_equal(tn.structure_matrix, tn0.structure_matrix):
                raise ValueError("all networks of a batch must share one structure matrix")
        self.structure_matrix = np.asarray(tn0.structure_matrix)
        self.s_i, self.s_j, self.s_k = s_i, s_j, s_k
        self.j_ij = _per_point(j_ij, self.n_points, "j_ij")
        self.h_k = _per_point(h_k, self.n_points, "h_k")
        self.hamiltonian = _per_point(hamiltonian, self.n_points, "hamiltonian")
        self.dts = _per_point(list(dts) if not isinstance(dts, _PerPoint) else dts, self.n_points, "dts")
        if len({len(x) for x in self.dts}) != 1:
            raise ValueError("all points must use dt schedules of the same length")
        self.d_max = d_max
        self.max_iterations = max_iterations
        self.convergence_error = convergence_error
        self.device = device
        self.stream = stream
        self._engine: Engine | None = None
        self._gates_key = None
        self._hams_uploaded = False
        self._hams_uploaded_obj = None
        self._synced = None
        self._synced_engine = None
        self._state_complex = None
        self.results = None

    # ------------------------------------------------------------------ host-side operators
    @property
    def spin_dims(self):
        return self.s_i[0].shape[0], self.s_j[0].shape[0]

    def coordination(self, ek):
        rows = np.nonzero(self.structure_matrix[:, ek])[0]
        return (int(np.sum(self.structure_matrix[rows[0], :] > 0)), int(np.sum(self.structure_matrix[rows[1], :] > 0)))

    def edge_hamiltonian(self, point, ek, i_neighbors=None, j_neighbors=None):
        """d^2 x d^2 complex Hamiltonian of one edge, same arithmetic as the reference (:420-454)."""
        if self.hamiltonian[point] is not None:
            return self.hamiltonian[point]
        if i_neighbors is None:
            i_neighbors, j_neighbors = self.coordination(ek)
        di, dj = self.spin_dims
        interaction = np.zeros((di * dj, di * dj), dtype=complex)
        for a, b in zip(self.s_i, self.s_j):
            interaction += np.kron(a, b)
        field_i = np.zeros((di * dj, di * dj), dtype=complex)
        field_j = np.zeros((di * dj, di * dj), dtype=complex)
        for s in self.s_k:
            field_i += np.kron(s, np.eye(dj))
            field_j += np.kron(np.eye(di), s)
        return self.j_ij[point][ek] * interaction + self.h_k[point] * (field_i / i_neighbors + field_j / j_neighbors)

    def _params_key(self):
        """Value key of everything the per-edge Hamiltonians depend on (operators by identity)."""
        if all(h is None for h in self.hamiltonian):
            try:  # the common case in one go: couplings and fields as bytes
                values = (np.asarray(self.j_ij, dtype=complex).tobytes(), np.asarray(self.h_k, dtype=complex).tobytes())
            except (TypeError, ValueError):
                values = (tuple(tuple(complex(x) for x in j) for j in self.j_ij), tuple(complex(h) for h in self.h_k))
            return (None, values, tuple(id(a) for a in self.s_i), tuple(id(a) for a in self.s_j), tuple(id(a) for a in self.s_k))
        return (tuple(id(h) for h in self.hamiltonian),
                tuple(tuple(complex(x) for x in j) if h is None else () for j, h in zip(self.j_ij, self.hamiltonian)),
                tuple(complex(h) for h in self.h_k),
                tuple(id(a) for a in self.s_i), tuple(id(a) for a in self.s_j), tuple(id(a) for a in self.s_k))

    def _all_hamiltonians(self):
        key = self._params_key()
        if getattr(self, "_hams_cache_key", None) == key:
            return self._hams_cache
        out = self._compute_hamiltonians()
        self._hams_cache_key, self._hams_cache = key, out
        self._hams_refs = (list(self.hamiltonian), list(self.s_i), list(self.s_j), list(self.s_k))  # keep ids alive
        return out

    def _compute_hamiltonians(self):
        """[point][edge] two-site Hamiltonians.  Same arithmetic per element as edge_hamiltonian / the reference
        (:420-454): j * sum_n kron(s_i[n], s_j[n]) + h * (sum kron(s_k, 1) / z_i + sum kron(1, s_k) / z_j), evaluated
        for all points of an edge at once."""
        m = self.structure_matrix.shape[1]
        di, dj = self.spin_dims
        d2 = di * dj
        out = np.empty((self.n_points, m, d2, d2), dtype=complex)
        interaction = np.zeros((d2, d2), dtype=complex)
        for a, b in zip(self.s_i, self.s_j):
            interaction += np.kron(a, b)
        field_i = np.zeros((d2, d2), dtype=complex)
        field_j = np.zeros((d2, d2), dtype=complex)
        for sk in self.s_k:
            field_i += np.kron(sk, np.eye(dj))
            field_j += np.kron(np.eye(di), sk)
        plain = [p for p in range(self.n_points) if self.hamiltonian[p] is None]
        if plain:
            h = np.array([self.h_k[p] for p in plain])
            for ek in range(m):
                zi, zj = self.coordination(ek)
                field = field_i / zi + field_j / zj
                j = np.array([self.j_ij[p][ek] for p in plain])
                out[plain, ek] = j[:, None, None] * interaction + h[:, None, None] * field
        for p in range(self.n_points):
            if self.hamiltonian[p] is not None:
                out[p, :] = np.asarray(self.hamiltonian[p], dtype=complex)
        return out

    @staticmethod
    def _distinct_gate_inputs(hams, dts_matrix):
        """Distinct (H, dt_0 .. dt_{k-1}) rows over all (point, edge) pairs: (first, which) with rows[first][which] ==
        rows.  dts_matrix: [n_points, k] time steps."""
        n_p, m = hams.shape[:2]
        d2 = hams.shape[-1]
        k = dts_matrix.shape[1]
        rows = np.empty((n_p * m, d2 * d2 + k), dtype=np.complex128)
        rows[:, :d2 * d2] = hams.reshape(n_p * m, d2 * d2)
        rows[:, d2 * d2:] = np.repeat(dts_matrix, m, axis=0)
        rows += 0.0  # -0.0 -> +0.0: equal values, equal bytes
        _, first, which = np.unique(rows.view(np.dtype((np.void, rows.shape[1] * 16))).ravel(), return_index=True,
                                    return_inverse=True)
        return first, which.reshape(n_p, m)

    @staticmethod
    def _expm_distinct(hams, dts_of_point, first, which):
        """expm(-dt H) for the distinct rows only, through ONE stacked scipy.linalg.expm call, which applies the same
        Pade routine per matrix (bit-identical to the reference's one-matrix calls, :470-473; tests/test_host.py)."""
        n_p, m = hams.shape[:2]
        d2 = hams.shape[-1]
        dts = np.asarray(dts_of_point, dtype=np.complex128)
        if np.all(dts.imag == 0.0):
            dts = dts.real
        uniq_h = hams.reshape(n_p * m, d2, d2)[first]
        uniq_dt = np.repeat(dts, m)[first]
        return linalg.expm(-uniq_dt[:, None, None] * uniq_h)[which]

    def _gates(self, hams, dts_of_point):
        """expm(-dt H) per (point, edge) for one time step per point."""
        dts = np.array([complex(x) for x in dts_of_point])
        first, which = self._distinct_gate_inputs(hams, dts[:, None])
        return self._expm_distinct(hams, dts, first, which)

    # ------------------------------------------------------------------ engine lifecycle
    def _current_dims(self):
        tn = self.tensor_networks[0]
        return [len(w) for w in tn.weights]

    def _wants_complex(self, extra=(), in_sync=False):
        """complex128 engine needed?  The tensors are scanned only when they are about to be (re-)uploaded; while the
        device state is the current one the answer of that scan is reused."""
        if not in_sync or self._state_complex is None:
            self._state_complex = any((t.dtype.kind == "c" if isinstance(t, np.ndarray) else np.iscomplexobj(t))
                                      and np.any(np.imag(t) != 0.0) for tn in self.tensor_networks for t in tn.tensors)
        if self._state_complex:
            return True
        for a in extra:
            if np.iscomplexobj(a) and np.any(np.imag(a) != 0.0):
                return True
        return False

    def _ensure_engine(self, extra_ops=()):
        """Create (or re-create when shapes / dtype changed) the engine and upload the host state."""
        dims = self._current_dims()
        hams = self._all_hamiltonians()
        in_sync = self._state_in_sync()
        need_complex = self._wants_complex((hams,) + tuple(extra_ops), in_sync) or any(
            np.iscomplexobj(np.asarray(x)) and np.any(np.imag(np.asarray(x)) != 0.0) for dts in self.dts for x in dts)
        d_max = _NO_TRUNCATION if self.d_max is None else int(self.d_max)
        eng = self._engine
        if (eng is None or list(eng.edge_dims) != dims or eng.d_max != d_max or (need_complex and not eng.complex)):
            if eng is not None:
                eng.close()
            eng = Engine(self.structure_matrix, self.tensor_networks[0].tensors[0].shape[0], dims, d_max,
                         need_complex, batch=self.n_points, device=self.device, stream=self.stream)
            eng.reserve_gate_slots(len(self.dts[0]) + 1)
            self._engine = eng
            self._gates_key = None
            self._hams_uploaded = False
        self._hams = hams
        if not self._hams_uploaded or self._hams_uploaded_obj is not hams:
            eng.set_hamiltonians(hams)
            self._hams_uploaded, self._hams_uploaded_obj = True, hams
        if not (in_sync and self._synced_engine is eng):
            self.upload_state()
        return eng

    def _state_in_sync(self):
        """True when every tensor / weight array is the object last uploaded to or downloaded from the device AND its
        content stamp is unchanged: the reference replaces list entries when it updates (:294-296) and reads the
        arrays at call time, so both a replaced entry and an in-place edit must trigger a re-upload."""
        refs = getattr(self, "_synced", None)
        if refs is None or self._synced_engine is not self._engine:
            return False
        for tn, (ts, ws, stamps) in zip(self.tensor_networks, refs):
            if len(tn.tensors) != len(ts) or len(tn.weights) != len(ws):
                return False
            if any(a is not b for a, b in zip(tn.tensors, ts)) or any(a is not b for a, b in zip(tn.weights, ws)):
                return False
            if stamps is not None and stamps != [_fingerprint(a) for a in list(tn.tensors) + list(tn.weights)]:
                return False
        return True

    def _mark_synced(self):
        stamp = self.detect_inplace_edits
        self._synced = [(list(tn.tensors), list(tn.weights),
                         [_fingerprint(a) for a in list(tn.tensors) + list(tn.weights)] if stamp else None)
                        for tn in self.tensor_networks]
        self._synced_engine = self._engine

    def upload_state(self):
        eng = self._engine
        n, m = self.structure_matrix.shape
        for t in range(n):
            stacked = np.stack([np.asarray(tn.tensors[t]) for tn in self.tensor_networks])
            eng.set_tensor(t, stacked if eng.complex else np.real(stacked))
        for e in range(m):
            eng.set_weights(e, np.stack([np.asarray(tn.weights[e], dtype=float) for tn in self.tensor_networks]))
        self._mark_synced()

    def download_state(self):
        """Write the device state back into the TensorNetwork objects (complex128 tensors, as the
        reference leaves them after an update, :435 + :483)."""
        eng = self._engine
        n, m = self.structure_matrix.shape
        # one conversion per tensor / edge; a point's entry is its row of that array (a view: per-point storage is
        # disjoint, in-place edits of one network touch no other)
        for t in range(n):
            vals = eng.get_tensor(t).astype(np.complex128, copy=False)
            for tn, row in zip(self.tensor_networks, vals):
                tn.tensors[t] = row
        for e in range(m):
            vals = eng.get_weights(e)
            for tn, row in zip(self.tensor_networks, vals):
                tn.weights[e] = row
        self._state_complex = bool(eng.complex)
        self._mark_synced()

    def _set_schedule_gates(self):
        key = ("schedule", tuple(tuple(complex(x) for x in d) for d in self.dts), id(self._hams))
        if self._gates_key == key:
            return
        n_dts = len(self.dts[0])
        # one search for the distinct (Hamiltonian, schedule) rows serves every time step of the schedule
        dts = np.array([[complex(x) for x in d] for d in self.dts]).reshape(self.n_points, n_dts)
        first, which = self._distinct_gate_inputs(self._hams, dts)
        for s in range(n_dts):
            self._engine.set_gates(s, self._expm_distinct(self._hams, dts[:, s], first, which))
        self._gates_key = key

    # ------------------------------------------------------------------ public API
    def simple_update(self, dt=None, n_sweeps=1):
        """n_sweeps sweeps over all edges with time step dt (shared or per_point)."""
        dts = _per_point(0.1 if dt is None else dt, self.n_points, "dt")
        gates = self._gates(self._all_hamiltonians(), dts)
        eng = self._ensure_engine((gates,))
        scratch = len(self.dts[0])
        eng.set_gates(scratch, gates)
        eng.sweep(scratch, n_sweeps)
        self.download_state()

    def run(self):
        """SimpleUpdate.run() for every point.  Returns the per-point logger dicts as a sequence (`_RunLogs`: indexable,
        iterable, sliceable like a list; a point's dict is built when it is first looked at)."""
        eng = self._ensure_engine()
        self._set_schedule_gates()
        n_dts = len(self.dts[0])
        # the reference compares dt VALUES with dts[-1] (:176); with per-point schedules use point 0's pattern
        is_last = [int(self.dts[0][s] == self.dts[0][-1]) for s in range(n_dts)]
        res = eng.run(is_last, self.max_iterations, self.convergence_error)
        self.download_state()
        self.results = _RunLogs(res, [list(d) for d in self.dts])
        return self.results

    def energy_per_site(self) -> np.ndarray:
        return self._ensure_engine().energy()

    def tensor_pair_rdm(self, common_edge) -> np.ndarray:
        return self._ensure_engine().pair_rdm(common_edge)

    def tensor_rdm(self, tensor_index) -> np.ndarray:
        return self._ensure_engine().site_rdm(tensor_index)

    def pair_expectation_per_site(self, operator) -> np.ndarray:
        op = np.asarray(operator, dtype=complex)
        return np.real(self._ensure_engine((op,)).pair_expectation_per_site(op))

    def expectation_per_site(self, operator) -> np.ndarray:
        op = np.asarray(operator, dtype=complex)
        return np.real(self._ensure_engine().expectation_per_site(op))

    @property
    def engine(self) -> Engine:
        return self._engine if self._engine is not None else self._ensure_engine()


class SimpleUpdate:
    """Drop-in for tnsu.simple_update.SimpleUpdate (reference :12-667), one network, GPU engine underneath."""

    def __init__(self, tensor_network: TensorNetwork, j_ij: list, h_k: float, s_i: list, s_j: list, s_k: list,
                 dts: list, d_max: int = 2, max_iterations: int = 1000, convergence_error: float = 1e-6,
                 log_energy: bool = False, print_process: bool = True, hamiltonian=None, device: int = 0):
        self.tensor_network = tensor_network
        self.j_ij, self.s_i, self.s_j, self.h_k, self.s_k = j_ij, s_i, s_j, h_k, s_k
        self.d_max = d_max
        self.dts = dts
        self.max_iterations = max_iterations
        self.convergence_error = convergence_error
        self.dt = 0.1
        self.old_weights = None
        self.hamiltonian = hamiltonian
        self.logger = {"error": [], "dt": [], "iteration": [], "energy": [], "j_ij": j_ij, "h_k": h_k, "d_max": d_max,
                       "max_iterations": max_iterations, "convergence_error": convergence_error}
        self.log_energy = log_energy
        self.print_process = print_process
        self.converged = False
        self.device = device
        self._batch: SimpleUpdateBatch | None = None

    # ---- reference properties (:95-105) ----
    @property
    def tensors(self):
        return self.tensor_network.tensors

    @property
    def weights(self):
        return self.tensor_network.weights

    @property
    def structure_matrix(self):
        return self.tensor_network.structure_matrix

    def _impl(self) -> SimpleUpdateBatch:
        """The one-point batch behind this object; parameters are re-read on every call because the
        reference reads self.j_ij / self.h_k / self.d_max / ... at call time."""
        b = self._batch
        if b is None or b.tensor_networks[0] is not self.tensor_network:
            b = SimpleUpdateBatch([self.tensor_network], self.j_ij, self.h_k, self.s_i, self.s_j, self.s_k,
                                  list(self.dts), self.d_max, self.max_iterations, self.convergence_error,
                                  self.hamiltonian, device=self.device, detect_inplace_edits=True)
            self._batch = b
        b.j_ij, b.h_k, b.hamiltonian = [self.j_ij], [self.h_k], [self.hamiltonian]
        b.s_i, b.s_j, b.s_k = self.s_i, self.s_j, self.s_k
        if [list(self.dts)] != b.dts:
            b.dts = [list(self.dts)]
            b._gates_key = None
            if b._engine is not None and b._engine.n_slots != len(self.dts) + 1:
                b._engine.reserve_gate_slots(len(self.dts) + 1)
        b.d_max, b.max_iterations, b.convergence_error = self.d_max, self.max_iterations, self.convergence_error
        return b  # changed parameters are picked up by value (SimpleUpdateBatch._params_key)

    # ---- run loop (:107-194) ----
    def run(self):
        """The whole dt / iteration / convergence loop.  With print_process=False it runs on the device in one call
        (tnsu_run: per-point state machine, no host round trip per sweep); with print_process=True the same loop is
        driven from here check by check so that every progress line appears live with its wall-clock times, exactly
        as the reference prints them (:142-172)."""
        error = np.inf
        impl = self._impl()
        if len(self.dts) > 0 and self.max_iterations > 0:
            if self.print_process:
                error = self._run_verbose(impl)
                if self.converged:
                    return
            else:
                log = impl.run()[0]
                self.logger["error"] += log["error"]
                self.logger["dt"] += log["dt"]
                self.logger["iteration"] += log["iteration"]
                if self.log_energy:
                    self.logger["energy"] += log["energy"]
                if log["error"]:
                    error = log["error"][-1]
                self.dt = log["final_dt"]
                self.converged = log["converged"]
        elif len(self.dts) > 0:
            self.dt = self.dts[-1]
        self.tensor_network.su_logger = self.logger
        if self.converged:
            print("==> Simple Update converged. final error is {:4.10f} < {:4.10f}.".format(error, self.convergence_error))
        else:
            print("==> Simple Update did not converged. final error is {:4.10f} > {:4.10f}.".format(
                error, self.convergence_error))

    def _run_verbose(self, impl: "SimpleUpdateBatch"):
        """run() driven from the host (reference :117-188 line by line): sweeps, convergence error and energy are the
        same device kernels, called through tnsu_sweep / tnsu_sweep_error / tnsu_energy."""
        import time

        eng = impl._ensure_engine()
        impl._set_schedule_gates()
        error, total_time, step = np.inf, 0.0, 0
        try:
            for slot, dt in enumerate(self.dts):
                start_time = time.time()
                self.dt = dt
                for i in range(self.max_iterations):
                    step += 1
                    eng.sweep(slot, 1)
                    if i % 2 == 0 and i > 0:
                        error = float(eng.sweep_error()[0])
                        elapsed = time.time() - start_time
                        total_time += elapsed
                        energy = float(eng.energy()[0])
                        self.logger["error"].append(error)
                        self.logger["dt"].append(dt)
                        self.logger["iteration"].append(step)
                        if self.log_energy:
                            self.logger["energy"].append(energy)
                        print("| D max: {:2d} | dt: {:8.6f} | iteration: {:5d}/{:5d} | convergence error: {:13.10f} | "
                              "energy per-site: {: 16.11f} | iteration time: {:5.1f} sec | tot time:{:7.1f} min".format(
                                  self.d_max, dt, i, self.max_iterations, error, np.round(energy, 10), elapsed,
                                  total_time // 60 + (total_time % 60) / 60))
                        start_time = time.time()
                        if error <= self.convergence_error and dt == self.dts[-1]:
                            self.converged = True
                            self.tensor_network.su_logger = self.logger
                            print("==> Simple Update converged. final error is {:4.10f} < {:4.10f}.".format(
                                error, self.convergence_error))
                            return error
                        if error <= self.convergence_error:
                            break
        finally:
            impl.download_state()
        return error

    def compute_convergence_error(self):
        """max_e ||lambda_e(old) - lambda_e(new)||_2 (:196-206).  With explicit old_weights the host lists are
        compared as the reference does; otherwise the value the kernels accumulated during the last sweep."""
        if self.old_weights is not None:
            return np.max([np.sqrt(np.sum(np.square(o - n))) for o, n in zip(self.old_weights, self.weights)])
        return float(self._impl().engine.sweep_error()[0])

    def simple_update(self):
        """One sweep over all edges with the current self.dt (:208-296)."""
        self._impl().simple_update(self.dt, 1)

    # ---- structure lookups (:298-338) ----
    def get_tensors(self, edge: int):
        rows = np.nonzero(self.structure_matrix[:, edge])[0]
        axes = self.structure_matrix[rows, edge]
        return ({"tensor": np.copy(self.tensors[rows[0]]), "index": rows[0], "dim": axes[0]},
                {"tensor": np.copy(self.tensors[rows[1]]), "index": rows[1], "dim": axes[1]})

    def get_other_edges(self, tensor_idx, edge_to_delete):
        edges = np.nonzero(self.structure_matrix[tensor_idx, :])[0]
        edges = edges[edges != edge_to_delete]
        return {"edges": edges, "dims": self.structure_matrix[tensor_idx, edges]}

    def get_edges(self, tensor_idx: int):
        return self.tensor_network.get_edges(tensor_idx=tensor_idx)

    def get_hamiltonian(self, i_neighbors, j_neighbors, ek):
        return self._impl().edge_hamiltonian(0, ek, i_neighbors, j_neighbors)

    # ---- observables (:498-661) ----
    def tensor_rdm(self, tensor_index):
        return self._impl().tensor_rdm(tensor_index)[0]

    def tensor_pair_rdm(self, common_edge):
        return self._impl().tensor_pair_rdm(common_edge)[0]

    def tensor_expectation(self, tensor_index, operator):
        return np.trace(np.matmul(self.tensor_rdm(tensor_index), operator))

    def tensor_pair_expectation(self, common_edge, operator):
        return np.einsum("ijkl,ijkl->", self.tensor_pair_rdm(common_edge), operator)

    def energy_per_site(self):
        return float(self._impl().energy_per_site()[0])

    def pair_expectation_per_site(self, operator):
        return float(self._impl().pair_expectation_per_site(operator)[0])

    def expectation_per_site(self, operator):
        return float(self._impl().expectation_per_site(operator)[0])

    def absorb_weights(self, tensor, edges_dims):
        return self.tensor_network.absorb_weights(tensor=tensor, edges_dims=edges_dims)

    def absorb_inverse_weights(self, tensor, edges_dims):
        return self.tensor_network.absorb_inverse_weights(tensor=tensor, edges_dims=edges_dims)
