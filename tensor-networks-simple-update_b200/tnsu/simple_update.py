"""tnsu.simple_update -> tnsu_b200.simple_update (reference src/tnsu/simple_update.py:12-667)."""
from tnsu_b200.simple_update import SimpleUpdate, SimpleUpdateBatch, per_point  # noqa: F401
from tnsu_b200.tensor_network import TensorNetwork  # noqa: F401  (the reference module imports it too, :9)
from tnsu.utils import l2  # noqa: F401
