"""pytest configuration: import paths, the `gpu` marker, and golden-fixture helpers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tensor-networks-simple-update_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def unpack_list(npz, prefix):
    return [npz[f"{prefix}_{i}"] for i in range(int(npz[f"{prefix}_count"]))]


@pytest.fixture(scope="session")
def tnsu():
    import tnsu_b200

    return tnsu_b200


def build_case(name):
    """oracle/cases.py case expanded with the PACKAGE's structure matrices / spin operators (pinned to the
    reference's by test_host.py)."""
    import cases
    import tnsu_b200

    return cases.build_case(name, smc=tnsu_b200.smc, spin_operators=tnsu_b200.spin_operators)
