"""`TensorNetwork`: the state container of the drop-in API — a list of site tensors, a list of bond weight
vectors and the structure matrix.  Mirrors the attribute / method surface of the reference's
`tnsu.tensor_network.TensorNetwork` (src/tnsu/tensor_network.py:18-356).

It stays on the host on purpose: random initial tensors come from the legacy
global numpy stream in the reference's order so that seeds reproduce bit for bit, and pickle
save/load keeps the reference's `state_dict` format.  The engine receives these arrays through
`SimpleUpdate`.
"""
from __future__ import annotations

import copy as cp
import os
import pathlib
import pickle
from typing import Optional

import numpy as np

from .structure_matrix_constructor import is_valid

DEFAULT_NETWORKS_FOLDER = str(pathlib.Path(__file__).parent / "networks")


def _scale_axes(tensor, vectors, axes):
    """tensor with axis axes[n] multiplied by vectors[n], applied leg after leg like the reference's
    absorb_* einsum chain (tensor_network.py:242-316)."""
    for vec, axis in zip(vectors, axes):
        shape = [1] * tensor.ndim
        shape[axis] = len(vec)
        tensor = tensor * np.reshape(vec, shape)
    return tensor


class TensorNetwork:
    """tensors[i]: ndarray (spin_dim, D_1..D_z); weights[j]: 1-D float ndarray; structure_matrix: int n x m."""

    def __init__(self, structure_matrix, tensors: Optional[list] = None, weights: Optional[list] = None,
                 spin_dim: int = 2, virtual_dim: int = 3, dir_path: Optional[str] = None,
                 network_name="tensor_network", random_init_real_loc: Optional[float] = 1.0,
                 random_init_real_scale: Optional[float] = 1.0, random_init_imag_loc: Optional[float] = None,
                 random_init_imag_scale: Optional[float] = None):
        real_init = random_init_real_loc is not None and random_init_real_scale is not None
        imag_init = random_init_imag_loc is not None and random_init_imag_scale is not None
        assert real_init or imag_init, (
            "random initialisation needs (loc, scale) for the real part, the imaginary part, or both; got "
            f"real=({random_init_real_loc}, {random_init_real_scale}) imag=({random_init_imag_loc}, {random_init_imag_scale})")
        assert 0 < spin_dim == int(spin_dim), f"spin_dim must be a positive integer, got {spin_dim}"
        assert is_valid(structure_matrix), "invalid structure matrix"

        n, m = structure_matrix.shape
        if tensors is None:
            # one draw per tensor in row order from the legacy global stream (reference :76-105)
            tensors = []
            for i in range(n):
                shape = [spin_dim] + [virtual_dim] * int(np.sum(structure_matrix[i, :] > 0))
                new = None
                if real_init:
                    new = np.random.normal(loc=random_init_real_loc * np.ones(shape), scale=random_init_real_scale)
                if imag_init:
                    imag = 1j * np.random.normal(loc=random_init_imag_loc * np.ones(shape), scale=random_init_imag_scale)
                    new = imag if new is None else new + imag
                tensors.append(new)
        if weights is None:
            # uniform 1/D on every edge, D read from the first tensor on that edge (reference :107-115)
            weights = []
            for j in range(m):
                i = int(np.nonzero(structure_matrix[:, j])[0][0])
                dim = tensors[i].shape[structure_matrix[i, j]]
                weights.append(np.ones(dim, dtype=float) / dim)

        self.virtual_dim = virtual_dim
        self.spin_dim = spin_dim
        self.tensors = tensors
        self.weights = weights
        self.structure_matrix = structure_matrix
        self.dir_path = dir_path if dir_path is not None else DEFAULT_NETWORKS_FOLDER
        self.network_name = network_name
        self.su_logger = None
        self.state_dict = None
        assert self.is_valid(), "tensors, weights and structure matrix do not fit together"

    # ---- validation (reference :130-175) ----
    def is_valid(self) -> bool:
        n, m = self.structure_matrix.shape
        if m != len(self.weights):
            print(f"structure matrix has {m} columns (edges) but {len(self.weights)} weight vectors were given")
            return False
        if n != len(self.tensors):
            print(f"structure matrix has {n} rows (tensors) but {len(self.tensors)} tensors were given")
            return False
        for i in range(n):
            legs = int(np.sum(self.structure_matrix[i, :] > 0))
            if self.tensors[i].ndim - 1 != legs:
                print(f"tensor {i}: {self.tensors[i].ndim - 1} virtual axes, but row {i} of the structure matrix "
                      f"connects it to {legs} edges")
                return False
        for i in range(n):
            for j in np.nonzero(self.structure_matrix[i, :])[0]:
                axis = self.structure_matrix[i, j]
                if self.tensors[i].shape[axis] != len(self.weights[j]):
                    print(f"tensor {i} axis {axis} has size {self.tensors[i].shape[axis]} but the weight vector of "
                          f"edge {j} has {len(self.weights[j])} entries")
                    return False
        return True

    # ---- persistence, reference pickle format (reference :177-230) ----
    def create_state_dict(self):
        self.state_dict = {
            "tensors": self.tensors, "weights": self.weights, "structure_matrix": self.structure_matrix,
            "path": self.dir_path, "spin_dim": self.spin_dim, "virtual_size": self.virtual_dim,
            "network_name": self.network_name, "su_logger": self.su_logger,
        }

    def unpack_state_dict(self):
        sd = self.state_dict
        self.tensors, self.weights, self.structure_matrix = sd["tensors"], sd["weights"], sd["structure_matrix"]
        self.dir_path, self.spin_dim, self.virtual_dim = sd["path"], sd["spin_dim"], sd["virtual_size"]
        self.network_name, self.su_logger = sd["network_name"], sd["su_logger"]

    def save_network(self, filename="tensor_network"):
        print(f"saving tensor network {self.network_name or filename!r}")
        self.create_state_dict()
        if self.network_name is None:
            self.network_name = filename
        os.makedirs(self.dir_path, exist_ok=True)
        with open(os.path.join(self.dir_path, self.network_name + ".pkl"), "wb") as outfile:
            pickle.dump(self.state_dict, outfile, pickle.DEFAULT_PROTOCOL)

    def load_network(self, network_full_path=None):
        print("loading tensor network")
        path = os.path.join(self.dir_path, self.network_name + ".pkl") if network_full_path is None else network_full_path
        try:
            with open(path, "rb") as infile:
                self.state_dict = pickle.load(infile)
                self.unpack_state_dict()
        except Exception as error:  # noqa: BLE001 - the reference swallows and prints (:226-230)
            print(f"could not load a tensor network from {path}: {error!r}")

    # ---- structure lookups and host-side weight helpers (reference :232-356) ----
    def get_edges(self, tensor_idx: int) -> dict:
        edges = np.nonzero(self.structure_matrix[tensor_idx, :])[0]
        return {"edges": edges, "dims": self.structure_matrix[tensor_idx, edges]}

    def absorb_weights(self, tensor, edges_dims):
        return _scale_axes(tensor, [self.weights[e] for e in edges_dims["edges"]], edges_dims["dims"])

    def absorb_inverse_weights(self, tensor, edges_dims):
        return _scale_axes(tensor, [np.power(self.weights[e], -1) for e in edges_dims["edges"]], edges_dims["dims"])

    def absorb_sqrt_weights(self, tensor, edges_dims):
        return _scale_axes(tensor, [np.sqrt(self.weights[e]) for e in edges_dims["edges"]], edges_dims["dims"])

    def absorb_sqrt_inverse_weights(self, tensor, edges_dims):
        return _scale_axes(tensor, [np.power(np.sqrt(self.weights[e]), -1) for e in edges_dims["edges"]],
                           edges_dims["dims"])

    def absorb_all_weights(self):
        for t in range(self.structure_matrix.shape[0]):
            self.tensors[t] = self.absorb_sqrt_weights(self.tensors[t], self.get_edges(t))

    def absorb_all_inverse_weights(self):
        for t in range(self.structure_matrix.shape[0]):
            self.tensors[t] = self.absorb_sqrt_inverse_weights(self.tensors[t], self.get_edges(t))

    def copy(self) -> "TensorNetwork":
        return cp.deepcopy(self)

    def get_tensor_network_state(self) -> "TensorNetwork":
        """Copy with sqrt(weights) absorbed into the neighbouring tensors: the PEPS of the state."""
        tn = self.copy()
        tn.absorb_all_weights()
        return tn
