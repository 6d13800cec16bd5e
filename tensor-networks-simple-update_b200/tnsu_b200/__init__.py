"""tnsu_b200 — B200-native Simple-Update engine behind the tnsu API.

    import tnsu_b200 as tnsu
    tn = tnsu.TensorNetwork(structure_matrix=tnsu.smc.infinite_structure_matrix_dict("peps"), virtual_dim=8)
    su = tnsu.SimpleUpdate(tn, j_ij, h_k, s_i, s_j, s_k, dts, d_max=8)
    su.run(); su.energy_per_site()

Host-only pieces (TensorNetwork, structure matrices, spin operators) import without a GPU; every
compute entry point goes through libtnsu_b200.so and raises if it or a CUDA device is missing.
"""
from . import structure_matrix_constructor as smc  # noqa: F401
from . import structure_matrix_constructor  # noqa: F401
from . import math_objects  # noqa: F401
from .math_objects import spin_operators  # noqa: F401
from .tensor_network import TensorNetwork  # noqa: F401
from .simple_update import SimpleUpdate, SimpleUpdateBatch, per_point  # noqa: F401
from .engine import Engine  # noqa: F401
from ._lib import EngineError, NonFiniteError  # noqa: F401

__version__ = "0.1.0"
