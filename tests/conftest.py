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


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def bit_equal(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(np.all(bits(a) == bits(b))) if a.dtype.kind == "f" else a.shape == b.shape and bool(np.all(a == b))


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b ** 2).sum()), 1e-30))


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle_py
    if not oracle_py.available("parity"):
        pytest.skip("oracle/_ref/liboracle_parity.so not built (run oracle/build_oracle.sh where /root/reference exists)")
    return oracle_py
