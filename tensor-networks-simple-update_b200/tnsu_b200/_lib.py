"""ctypes binding of libtnsu_b200.so (include/tnsu_b200.h).  There is no CPU fallback: if the library is
missing, cannot be loaded, or finds no CUDA device, the compute entry points raise."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtnsu_b200.so")
ABI_VERSION = 2

TNSU_F64, TNSU_C128 = 0, 1
ERRORS = {-1: "TNSU_E_ARG", -2: "TNSU_E_UNSUPPORTED", -3: "TNSU_E_CUDA", -4: "TNSU_E_NONFINITE", -5: "TNSU_E_STATE"}

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol declared in include/tnsu_b200.h (the drop-in boundary) and
# include/tnsu_b200_debug.h (profiling / debugging)
PROTOTYPES = {
    "tnsu_abi_version": (C.c_int, []),
    "tnsu_last_error": (C.c_char_p, [_vp]),
    "tnsu_schedule_levels": (C.c_int, [C.c_int, C.c_int, _i64p, _i32p]),
    "tnsu_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, _i64p, C.c_int, _i32p, C.c_int, C.c_int, C.c_int,
                              C.c_int, _vp]),
    "tnsu_destroy": (C.c_int, [_vp]),
    "tnsu_edge_dims": (C.c_int, [_vp, _i32p]),
    "tnsu_tensor_elems": (C.c_int64, [_vp, C.c_int]),
    "tnsu_num_levels": (C.c_int, [_vp]),
    "tnsu_edge_levels": (C.c_int, [_vp, _i32p]),
    "tnsu_set_tensor": (C.c_int, [_vp, C.c_int, _vp, C.c_int64]),
    "tnsu_get_tensor": (C.c_int, [_vp, C.c_int, _vp, C.c_int64]),
    "tnsu_set_weights": (C.c_int, [_vp, C.c_int, _f64p, C.c_int]),
    "tnsu_get_weights": (C.c_int, [_vp, C.c_int, _f64p, C.c_int]),
    "tnsu_reserve_gate_slots": (C.c_int, [_vp, C.c_int]),
    "tnsu_set_gates": (C.c_int, [_vp, C.c_int, _vp]),
    "tnsu_set_hamiltonians": (C.c_int, [_vp, _vp]),
    "tnsu_sweep": (C.c_int, [_vp, C.c_int, C.c_int]),
    "tnsu_sweep_error": (C.c_int, [_vp, _f64p]),
    "tnsu_energy": (C.c_int, [_vp, _f64p]),
    "tnsu_pair_rdm": (C.c_int, [_vp, C.c_int, _vp]),
    "tnsu_site_rdm": (C.c_int, [_vp, C.c_int, _vp]),
    "tnsu_pair_expectation_per_site": (C.c_int, [_vp, _vp, _vp]),
    "tnsu_expectation_per_site": (C.c_int, [_vp, _vp, _vp]),
    "tnsu_run": (C.c_int, [_vp, C.c_int, _i32p, C.c_int, C.c_double, C.c_int, _i32p, _f64p, _f64p, _i32p, _i32p,
                           _i32p, _i32p, _i32p]),
    "tnsu_stream": (_vp, [_vp]),
    "tnsu_synchronize": (C.c_int, [_vp]),
    "tnsu_launch_count": (C.c_int64, [_vp]),
    "tnsu_last_sweep_kernel_ms": (C.c_int, [_vp, _f64p, _i32p]),
    "tnsu_stats": (C.c_int, [_vp, _i64p, C.c_int]),
    "tnsu_profile_sweep": (C.c_int, [_vp, C.c_int, _f64p]),
    "tnsu_debug_edge": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _f64p, _i32p]),
}

_lib = None


class EngineError(RuntimeError):
    """An error code returned across the C ABI."""

    def __init__(self, code: int, message: str):
        super().__init__(f"{ERRORS.get(code, code)}: {message}")
        self.code = code


class NonFiniteError(EngineError, ValueError):
    """TNSU_E_NONFINITE: NaN / Inf met during an update.  Also a ValueError, which is what the reference raises
    there (scipy.linalg.rq / svd run with check_finite=True, simple_update.py:243-250, :404)."""


def error_from_code(code: int, message: str) -> EngineError:
    return NonFiniteError(code, message) if code == -4 else EngineError(code, message)


def load():
    """Load the shared library (once) and bind every prototype.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  The Simple-Update engine has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here means the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.tnsu_abi_version() != ABI_VERSION:
        raise ImportError(f"libtnsu_b200.so has ABI {lib.tnsu_abi_version()}, binding expects {ABI_VERSION}")
    _lib = lib
    return lib
