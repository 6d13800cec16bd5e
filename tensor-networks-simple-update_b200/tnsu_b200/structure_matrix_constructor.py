"""Structure matrices (rows = tensors, columns = edges, entry = 1-based axis of that tensor on that edge,
0 = not connected).  Same interface as the reference's `tnsu.structure_matrix_constructor`
(src/tnsu/structure_matrix_constructor.py:8-178); the matrices are pinned against the reference's in
tests/golden/structure_matrices.npz.  This is the scheduler's only input.
"""
from __future__ import annotations

import numpy as np

# iPEPS unit cells of Jahromi & Orus, "Universal tensor-network algorithm for any infinite lattice" (2019),
# stored as: for each tensor, the edge ids on its axes 1..z.
_CELLS = {
    "chain": [(0, 1), (0, 1)],
    "kagome_3": [(0, 1, 2, 3), (4, 3, 5, 1), (2, 5, 0, 4)],
    "peps": [(0, 1, 2, 3), (0, 1, 4, 5), (2, 3, 6, 7), (4, 5, 6, 7)],
    "star": [(0, 1, 2), (0, 3, 4), (1, 3, 5), (5, 6, 7), (4, 6, 8), (2, 7, 8)],
    "cube": [(0, 1, 2, 3, 4, 5), (0, 1, 6, 7, 8, 9), (6, 7, 10, 11, 12, 13), (2, 3, 10, 11, 14, 15),
             (4, 5, 16, 17, 18, 19), (8, 9, 16, 17, 20, 21), (12, 13, 20, 21, 22, 23), (14, 15, 18, 19, 22, 23)],
    "triangle": [(0, 1, 2, 3, 4, 5), (0, 6, 7, 8, 9, 10), (1, 6, 11, 12, 13, 14), (2, 11, 15, 16, 17, 18),
                 (3, 7, 15, 19, 20, 21), (8, 12, 16, 19, 22, 23), (4, 9, 17, 22, 24, 25), (10, 13, 18, 20, 24, 26),
                 (5, 14, 21, 23, 25, 26)],
    "kagome_12": [(0, 1, 2, 3), (4, 5, 6, 7), (4, 8, 9, 10), (0, 8, 11, 12), (1, 11, 13, 14), (5, 9, 13, 15),
                  (10, 12, 16, 17), (14, 15, 18, 19), (2, 18, 20, 21), (3, 16, 20, 22), (6, 17, 22, 23),
                  (7, 19, 21, 23)],
    "pyrochlore": [(0, 1, 2, 3, 4, 5), (6, 7, 8, 9, 10, 11), (6, 7, 12, 13, 14, 15), (8, 9, 12, 13, 16, 17),
                   (0, 10, 14, 16, 18, 19), (1, 11, 15, 17, 20, 21), (2, 3, 18, 20, 22, 23),
                   (4, 5, 19, 21, 22, 23)],
}


def _from_legs(legs) -> np.ndarray:
    m = 1 + max(e for row in legs for e in row)
    smat = np.zeros((len(legs), m), dtype=int)
    for t, row in enumerate(legs):
        for axis, e in enumerate(row, start=1):
            smat[t, e] = axis
    return smat


def infinite_structure_matrix_dict(name: str) -> np.ndarray:
    """Structure matrix of a named iPEPS unit cell (reference :8-90)."""
    return _from_legs(_CELLS[name])


def square_peps_pbc(side: int) -> np.ndarray:
    """side x side square lattice with periodic boundaries (reference :93-114): tensor i owns edge i on
    axis 4 (right) and edge i+n on axis 1 (down); its left / up neighbours' edges sit on axes 2 / 3."""
    n = side * side
    smat = np.zeros((n, 2 * n), dtype=int)
    for i in range(n):
        row, colm = divmod(i, side)
        left = row * side + (colm - 1) % side
        up = ((row - 1) % side) * side + colm
        smat[i, i] = 4
        smat[i, n + i] = 1
        smat[i, left] = 2
        smat[i, n + up] = 3
    return smat


def rectangular_peps_obc(height: int, width: int) -> np.ndarray:
    """height x width lattice with open boundaries (reference :117-145).  Edges are numbered site by site
    (row-major), "down" before "right"; a tensor's axes follow the fixed order up, left, down, right
    restricted to the bonds it has."""
    edges = []  # (site_a, site_b, kind) ; kind 0 = vertical (a above b), 1 = horizontal (a left of b)
    for i in range(height):
        for j in range(width):
            if i < height - 1:
                edges.append((i * width + j, (i + 1) * width + j, 0))
            if j < width - 1:
                edges.append((i * width + j, i * width + j + 1, 1))
    smat = np.zeros((height * width, len(edges)), dtype=int)
    # reference direction codes: node_a gets 4 (down) / 3 (right), node_b gets 2 (up) / 1 (left); axes are then
    # renumbered 1..z in ascending code order
    code = np.zeros_like(smat)
    for e, (a, b, kind) in enumerate(edges):
        code[a, e] = 4 if kind == 0 else 3
        code[b, e] = 2 if kind == 0 else 1
    for t in range(smat.shape[0]):
        es = np.nonzero(code[t])[0]
        ranks = np.argsort(np.argsort(code[t, es]))
        smat[t, es] = ranks + 1
    return smat


def is_valid(structure_matrix) -> bool:
    """Every row's non-zeros are a permutation of 1..z and every column has exactly two non-zeros
    (reference :148-178).  Prints the reason and returns False otherwise, never raises."""
    try:
        if len(structure_matrix.shape) != 2:
            print(f"structure_matrix must be a matrix, instead got a {len(structure_matrix.shape)} dimension tensor.")
            return False
        n, m = structure_matrix.shape
        for i in range(n):
            legs = np.sort(structure_matrix[i, structure_matrix[i, :] > 0])
            want = np.arange(1, len(legs) + 1)
            if not np.array_equal(legs, want):
                print(f"Error in structure_matrix given. For row [{i}] expected values are "
                      f"{'-'.join(map(str, want))}, instead got {'-'.join(map(str, legs))}.")
                return False
        for j in range(m):
            if np.sum(structure_matrix[:, j] > 0) != 2:
                print(f"Weight vector [{j}] is not connected to two tensors.")
                return False
        return True
    except Exception as exc:  # noqa: BLE001 - mirrors the reference's catch-all
        print(f"Failed with unexpected behavior. {exc}")
        return False
