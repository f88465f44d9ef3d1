import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fixtures():
    """The reference's own data files, parsed by tests/golden/make_golden.py."""
    return dict(np.load(os.path.join(GOLDEN, "ref_fixtures.npz")))


@pytest.fixture(scope="session")
def refout():
    """Outputs of the unmodified reference C on seeded inputs (tests/golden/make_golden.py)."""
    return dict(np.load(os.path.join(GOLDEN, "ref_outputs.npz")))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O
