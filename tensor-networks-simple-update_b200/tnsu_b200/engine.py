"""`Engine`: a thin object wrapper over one `tnsu_engine` handle of libtnsu_b200.so.

All numerics happen in the CUDA library; this class only marshals numpy arrays across the C ABI
(include/tnsu_b200.h) and raises `EngineError` for negative return codes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Engine:
    """`batch` independent networks of one structure matrix on one GPU."""

    def __init__(self, structure_matrix, spin_dim: int, edge_dims, d_max: int, complex_dtype: bool, batch: int = 1,
                 device: int = 0, stream=None):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        smat = np.ascontiguousarray(structure_matrix, dtype=np.int64)
        self.n, self.m = smat.shape
        self.d = int(spin_dim)
        self.batch = int(batch)
        self.complex = bool(complex_dtype)
        self.dtype = np.complex128 if self.complex else np.float64
        self.d_max = int(d_max)
        self.device = int(device)
        self._legs = [[int(e) for e in np.nonzero(smat[t])[0][np.argsort(smat[t, np.nonzero(smat[t])[0]])]]
                      for t in range(self.n)]
        dims = np.ascontiguousarray(edge_dims, dtype=np.int32)
        rc = self._lib.tnsu_create(C.byref(self._h), self.n, self.m, smat.ctypes.data_as(C.POINTER(C.c_int64)), self.d,
                                   dims.ctypes.data_as(C.POINTER(C.c_int32)), self.d_max,
                                   _lib.TNSU_C128 if self.complex else _lib.TNSU_F64, self.batch, self.device,
                                   C.c_void_p(stream) if stream else None)
        if rc != 0:
            raise _lib.EngineError(rc, self._lib.tnsu_last_error(None).decode())
        self.n_slots = 0

    # -- plumbing --
    def _check(self, rc):
        if rc != 0:
            raise _lib.error_from_code(rc, self._lib.tnsu_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.tnsu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    # -- shapes --
    @property
    def edge_dims(self) -> np.ndarray:
        out = np.zeros(self.m, dtype=np.int32)
        self._check(self._lib.tnsu_edge_dims(self._h, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def tensor_shape(self, t: int):
        dims = self.edge_dims
        return (self.d,) + tuple(int(dims[e]) for e in self._legs[t])

    @property
    def edge_levels(self) -> np.ndarray:
        out = np.zeros(self.m, dtype=np.int32)
        self._check(self._lib.tnsu_edge_levels(self._h, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    # -- state --
    def set_tensor(self, t: int, values: np.ndarray):
        """values: [batch, d, D_1..D_z] (or without the batch axis when batch == 1)."""
        a = np.ascontiguousarray(values, dtype=self.dtype).reshape(self.batch, -1)
        self._check(self._lib.tnsu_set_tensor(self._h, t, _ptr(a), a.shape[1]))

    def get_tensor(self, t: int, out: np.ndarray | None = None) -> np.ndarray:
        """Tensor t of every point, [batch, d, D_1..D_z].  `out`: a C-contiguous array of that shape and the engine dtype
        to download into (page-locked memory makes the copy run at PCIe speed; a fresh numpy array is pageable)."""
        shape = self.tensor_shape(t)
        n = int(np.prod(shape))
        if out is None:
            out = np.empty((self.batch, n), dtype=self.dtype)
        elif out.dtype != self.dtype or out.size != self.batch * n or not out.flags.c_contiguous:
            raise ValueError(f"out must be a C-contiguous {np.dtype(self.dtype).name} array of {self.batch * n} elements")
        self._check(self._lib.tnsu_get_tensor(self._h, t, _ptr(out), n))
        return out.reshape((self.batch,) + shape)

    def set_weights(self, e: int, values: np.ndarray):
        a = np.ascontiguousarray(values, dtype=np.float64).reshape(self.batch, -1)
        self._check(self._lib.tnsu_set_weights(self._h, e, a.ctypes.data_as(C.POINTER(C.c_double)), a.shape[1]))

    def get_weights(self, e: int, out: np.ndarray | None = None) -> np.ndarray:
        n = int(self.edge_dims[e])
        if out is None:
            out = np.empty((self.batch, n), dtype=np.float64)
        elif out.dtype != np.float64 or out.size != self.batch * n or not out.flags.c_contiguous:
            raise ValueError(f"out must be a C-contiguous float64 array of {self.batch * n} elements")
        out = out.reshape(self.batch, n)
        self._check(self._lib.tnsu_get_weights(self._h, e, out.ctypes.data_as(C.POINTER(C.c_double)), out.shape[1]))
        return out

    # -- operators --
    def reserve_gate_slots(self, n_slots: int):
        self._check(self._lib.tnsu_reserve_gate_slots(self._h, int(n_slots)))
        self.n_slots = int(n_slots)

    def _op_array(self, values, per_point_edge: bool) -> np.ndarray:
        d2 = self.d * self.d
        a = np.ascontiguousarray(values, dtype=np.complex128)
        want = (self.batch, self.m, d2, d2) if per_point_edge else (d2, d2)
        if a.size != int(np.prod(want)):
            raise ValueError(f"operator array has {a.size} elements, expected shape {want}")
        if not self.complex and np.any(a.imag != 0.0):
            raise ValueError("complex operator passed to a float64 engine")
        return a.reshape(want)

    def set_gates(self, slot: int, gates: np.ndarray):
        a = self._op_array(gates, True)
        self._check(self._lib.tnsu_set_gates(self._h, int(slot), _ptr(a)))

    def set_hamiltonians(self, hams: np.ndarray):
        a = self._op_array(hams, True)
        self._check(self._lib.tnsu_set_hamiltonians(self._h, _ptr(a)))

    # -- compute --
    def sweep(self, slot: int, n_sweeps: int = 1):
        self._check(self._lib.tnsu_sweep(self._h, int(slot), int(n_sweeps)))

    def sweep_error(self) -> np.ndarray:
        out = np.empty(self.batch, dtype=np.float64)
        self._check(self._lib.tnsu_sweep_error(self._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def energy(self) -> np.ndarray:
        out = np.empty(self.batch, dtype=np.float64)
        self._check(self._lib.tnsu_energy(self._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def pair_rdm(self, e: int) -> np.ndarray:
        out = np.empty((self.batch, self.d, self.d, self.d, self.d), dtype=np.complex128)
        self._check(self._lib.tnsu_pair_rdm(self._h, int(e), _ptr(out)))
        return out

    def site_rdm(self, t: int) -> np.ndarray:
        out = np.empty((self.batch, self.d, self.d), dtype=np.complex128)
        self._check(self._lib.tnsu_site_rdm(self._h, int(t), _ptr(out)))
        return out

    def pair_expectation_per_site(self, operator) -> np.ndarray:
        op = self._op_array(operator, False)
        out = np.empty(self.batch, dtype=np.complex128)
        self._check(self._lib.tnsu_pair_expectation_per_site(self._h, _ptr(op), _ptr(out)))
        return out

    def expectation_per_site(self, operator) -> np.ndarray:
        op = np.ascontiguousarray(operator, dtype=np.complex128).reshape(self.d, self.d)
        out = np.empty(self.batch, dtype=np.complex128)
        self._check(self._lib.tnsu_expectation_per_site(self._h, _ptr(op), _ptr(out)))
        return out

    def run(self, is_last_value, max_iterations: int, tol: float):
        """Whole run loop for every point.  Returns a dict of per-point arrays."""
        n_dts = len(is_last_value)
        cap = max(1, n_dts * (max_iterations // 2 + 1))
        last = np.ascontiguousarray(is_last_value, dtype=np.int32)
        i32 = lambda *s: np.zeros(s, dtype=np.int32)  # noqa: E731
        f64 = lambda *s: np.zeros(s, dtype=np.float64)  # noqa: E731
        res = dict(n_checks=i32(self.batch), log_error=f64(self.batch, cap), log_energy=f64(self.batch, cap),
                   log_slot=i32(self.batch, cap), log_step=i32(self.batch, cap), converged=i32(self.batch),
                   sweeps=i32(self.batch), final_slot=i32(self.batch))
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))  # noqa: E731
        fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
        self._check(self._lib.tnsu_run(self._h, n_dts, ip(last), int(max_iterations), float(tol), cap,
                                       ip(res["n_checks"]), fp(res["log_error"]), fp(res["log_energy"]),
                                       ip(res["log_slot"]), ip(res["log_step"]), ip(res["converged"]),
                                       ip(res["sweeps"]), ip(res["final_slot"])))
        return res

    # -- instrumentation --
    def synchronize(self):
        self._check(self._lib.tnsu_synchronize(self._h))

    @property
    def stream(self) -> int:
        return int(self._lib.tnsu_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self._lib.tnsu_launch_count(self._h))

    STAT_NAMES = ("launches", "lq_streamed", "lq_flagged", "svd_unconverged", "nonfinite_events", "fused_runs")

    def stats(self) -> dict:
        """Cumulative counters of this handle (tnsu_stats): what the streaming reduction was given and what came back
        flagged for the Householder fallback, Jacobi iterations that hit their sweep limit, non-finite events."""
        out = np.zeros(len(self.STAT_NAMES), dtype=np.int64)
        self._check(self._lib.tnsu_stats(self._h, out.ctypes.data_as(C.POINTER(C.c_int64)), len(out)))
        return dict(zip(self.STAT_NAMES, (int(x) for x in out)))

    def last_sweep_kernel_ms(self):
        ms = C.c_double()
        n = C.c_int32()
        self._check(self._lib.tnsu_last_sweep_kernel_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def profile_sweep(self, slot: int):
        """One sweep with per-kernel-class CUDA-event timing: ms of (LQ, SVD, recompose) launches."""
        ms = (C.c_double * 3)()
        self._check(self._lib.tnsu_profile_sweep(self._h, int(slot), ms))
        return {"lq": ms[0], "svd": ms[1], "recompose": ms[2]}

    def debug_edge(self, slot: int, edge: int, point: int = 0):
        li = np.zeros(64 * 64, dtype=np.complex128)
        lj = np.zeros(64 * 64, dtype=np.complex128)
        sig = np.zeros(64, dtype=np.float64)
        info = np.zeros(16, dtype=np.int32)
        self._check(self._lib.tnsu_debug_edge(self._h, slot, edge, point, _ptr(li), _ptr(lj),
                                              sig.ctypes.data_as(C.POINTER(C.c_double)),
                                              info.ctypes.data_as(C.POINTER(C.c_int32))))
        mi, ki, mj, kj, ni, nj, dnew, sweeps = (int(x) for x in info[:8])
        return dict(l_i=li[: mi * ki].reshape(mi, ki), l_j=lj[: mj * kj].reshape(mj, kj), sigma=sig[: min(ni, nj)],
                    d_new=dnew, svd_sweeps=sweeps,
                    phase_cycles=dict(zip(["theta", "jacobi", "extract"], (int(x) for x in info[8:11]))),
                    lq_cycles=[int(x) for x in info[11:16]])


def schedule_levels(structure_matrix) -> np.ndarray:
    """Dependency level of every edge (host-only query of the engine's scheduler, no GPU needed)."""
    lib = _lib.load()
    smat = np.ascontiguousarray(structure_matrix, dtype=np.int64)
    n, m = smat.shape
    out = np.zeros(m, dtype=np.int32)
    rc = lib.tnsu_schedule_levels(n, m, smat.ctypes.data_as(C.POINTER(C.c_int64)), out.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc < 0:
        raise _lib.EngineError(rc, lib.tnsu_last_error(None).decode())
    return out
