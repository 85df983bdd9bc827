import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "schwarz-lib_b200"))
sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def sz():
    """The product library; built on demand so the CPU suite is self-contained."""
    import schwz_b200
    if not os.path.exists(schwz_b200.LIB_PATH):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "schwarz-lib_b200"), "-s"])
    schwz_b200.load()
    return schwz_b200


@pytest.fixture(scope="session")
def ani4():
    """matrices/ani4_crop.mtx of the reference, committed as a fixture
    (tests/golden/ani4_crop.npz, made by tests/golden/make_golden.py)."""
    import numpy as np
    z = np.load(os.path.join(GOLDEN, "ani4_crop.npz"))
    return z["rowptr"], z["col"], z["val"]


@pytest.fixture(scope="session")
def gpu(sz):
    if sz.device_count() < 1:
        pytest.skip("no CUDA device")
    ctx = sz.Context(0)
    yield ctx
    ctx.close()
