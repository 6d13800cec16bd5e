"""`tnsu` — import-name compatibility for the B200 drop-in.

    import tnsu.simple_update as su
    from tnsu.tensor_network import TensorNetwork
    import tnsu.structure_matrix_constructor as smc
    from tnsu.math_objects import spin_operators

resolve to the classes of `tnsu_b200` (one level up in this source tree), so code written against
RoyElkabetz/Tensor-Networks-Simple-Update — its README snippets, `tests/test_main.py`, `src/tnsu/examples.py`,
the notebooks — runs on the CUDA engine without an edit.  Same module layout as the reference package
(src/tnsu/__init__.py is empty there too); `tnsu.examples` is the reference's own driver file and is not
part of this package (tests load it from the reference tree on top of these modules).
"""
