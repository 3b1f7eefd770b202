import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    d = {k: z[k] for k in z.files}
    for k in ("kind",):
        if k in d:
            d[k] = str(d[k])
    for k in ("p0", "p1", "R", "K_fro", "R_updated"):
        if k in d:
            d[k] = float(d[k])
    return d


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture(scope="session")
def gpr():
    import gpr_b200
    return gpr_b200


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build()
    return oracle
