"""tnsu.tensor_network -> tnsu_b200.tensor_network (reference src/tnsu/tensor_network.py:18-356)."""
from tnsu_b200.tensor_network import DEFAULT_NETWORKS_FOLDER, TensorNetwork  # noqa: F401
