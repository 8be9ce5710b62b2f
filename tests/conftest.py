import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from _oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference compiled into oracle/_ref (None when the prebuilt file is absent)."""
    from _oracle import Ref
    return Ref() if Ref.available() else None


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "stralg_golden.npz"))


def golden_cases(golden):
    return sorted({k.split("/")[0] for k in golden.files})


@pytest.fixture(scope="session")
def engine():
    """The product library; fails (does not skip) when it is missing or no GPU is visible."""
    import stralg_b200
    lib = stralg_b200.load()
    assert lib.b200sa_device_count() > 0, "gpu-marked test without a CUDA device"
    return stralg_b200
