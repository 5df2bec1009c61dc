import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def unit_vectors():
    return load_golden("ref_unit_vectors.npz")


@pytest.fixture(scope="session")
def cfg1():
    return load_golden("cfg1_run.npz")


@pytest.fixture(scope="session")
def cfg2():
    return load_golden("cfg2_run.npz")
