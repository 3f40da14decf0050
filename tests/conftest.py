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


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def lib_from_fixture(d, prefix="lib_"):
    from irbpp_b200 import shapes
    res = d[prefix + "res"]
    return shapes.ShapeLibrary.from_flat(d[prefix + "dims"], d[prefix + "ext"], d[prefix + "vol"],
                                         d[prefix + "maps"], d[prefix + "offsets"], float(res[0]), float(res[1]))


@pytest.fixture(scope="session")
def golden():
    return load_golden
