"""tnsu.structure_matrix_constructor -> tnsu_b200.structure_matrix_constructor (reference :8-178)."""
from tnsu_b200.structure_matrix_constructor import *  # noqa: F401,F403
from tnsu_b200.structure_matrix_constructor import (  # noqa: F401
    infinite_structure_matrix_dict, is_valid, rectangular_peps_obc, square_peps_pbc)
