"""The reference's own test file and example drivers, UNCHANGED, on top of the `tnsu` alias package.

Build-container only: /root/reference does not exist on the GPU box, and this container has no GPU, so what runs here
is everything of the reference suite that needs no device — module layout, imports, the two host-only tests of
tests/test_main.py (spin operators, structure-matrix validation) and the binding of the example drivers to the
drop-in classes.  The device-dependent reference tests are re-stated one to one in tests/test_parity_gpu.py
(test_known_answer_from_shipped_pickle_via_tnsu_names, test_reference_end_to_end_energies, test_reference_experiments).
Nothing is copied: the reference files are loaded from where they lie."""
import importlib.util
import os
import sys

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "tnsu")), reason="reference tree not present")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def reference_suite():
    import tnsu
    import tnsu.tensor_network as tn_mod

    assert "tensor-networks-simple-update_b200" in tnsu.__file__  # ours, not the reference package
    tn_mod.DEFAULT_NETWORKS_FOLDER = os.path.join(REF, "src", "tnsu", "networks")
    examples = _load("tnsu.examples", os.path.join(REF, "src", "tnsu", "examples.py"))
    tests = _load("reference_test_main", os.path.join(REF, "tests", "test_main.py"))
    yield examples, tests
    sys.modules.pop("tnsu.examples", None)
    sys.modules.pop("reference_test_main", None)


def test_reference_modules_bind_to_the_drop_in(reference_suite):
    import tnsu_b200

    examples, tests = reference_suite
    assert examples.su.SimpleUpdate is tnsu_b200.SimpleUpdate
    assert examples.TensorNetwork is tnsu_b200.TensorNetwork
    assert examples.smg.rectangular_peps_obc is tnsu_b200.smc.rectangular_peps_obc
    assert tests.su.SimpleUpdate is tnsu_b200.SimpleUpdate
    for fn in ("load_a_tensor_network_from_memory", "afh_chain_spin_half_ground_state_experiment",
               "afh_star_spin_half_ground_state_experiment", "afh_cubic_spin_half_ground_state_experiment",
               "fhf_pyrochlore_spin_half_ground_state_experiment", "transverse_ising_field_ground_state_experiment"):
        assert callable(getattr(examples, fn))


def test_reference_host_only_tests_pass_unchanged(reference_suite):
    _, tests = reference_suite
    tests.test_spin_operators()
    tests.test_structure_matrix_validation()


def test_reference_pickle_loads_through_the_drop_in(reference_suite, capsys):
    """load_a_tensor_network_from_memory builds the TensorNetwork and loads the shipped pickle with OUR classes; its
    energy evaluation needs the GPU, so stop right before it: the loader's first half, reproduced from its call."""
    import numpy as np
    import tnsu.structure_matrix_constructor as smc
    from tnsu.tensor_network import DEFAULT_NETWORKS_FOLDER, TensorNetwork

    smat = smc.rectangular_peps_obc(height=10, width=10)
    tn = TensorNetwork(structure_matrix=smat, network_name="AFH_10x10_obc_D_4", dir_path=DEFAULT_NETWORKS_FOLDER,
                       tensors=None, weights=None)
    tn.load_network()
    assert len(tn.tensors) == 100 and len(tn.weights) == 180 and np.array_equal(tn.structure_matrix, smat)
    assert tn.is_valid()
    capsys.readouterr()
