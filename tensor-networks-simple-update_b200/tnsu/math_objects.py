"""tnsu.math_objects -> tnsu_b200.math_objects (reference src/tnsu/math_objects.py:4-27)."""
from tnsu_b200.math_objects import spin_operators  # noqa: F401
