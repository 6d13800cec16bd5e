"""Spin-S operators, same interface as the reference's `tnsu.math_objects` (src/tnsu/math_objects.py:4-27)."""
from __future__ import annotations

import numpy as np


def spin_operators(spin: float = 0.5):
    """(Sx, Sy, Sz) as complex (2S+1)x(2S+1) matrices, basis ordered m = S, S-1, ..., -S."""
    n = int(2 * spin + 1)
    assert spin > 0, f"the spin should be a positive half integer, instead got {spin}."
    assert (n - (2 * spin + 1)) == 0, f"spin should be a half integer (i.e., 0.5, 1, 1.5, ...), instead got {spin}."
    m = spin - np.arange(n)
    sz = np.diag(m).astype(complex)
    raising = np.zeros((n, n), dtype=complex)  # <m+1| S+ |m> = sqrt(S(S+1) - m(m+1))
    raising[np.arange(n - 1), np.arange(1, n)] = np.sqrt(spin * (spin + 1) - m[1:] * (m[1:] + 1))
    sx = (raising + raising.conj().T) / 2
    sy = (raising - raising.conj().T) / 2j
    return sx, sy, sz
