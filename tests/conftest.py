import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


LAYER_CASES = sorted(
    os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
    if os.path.basename(p)[:-4] not in ("model_rc_vi", "amortized_re", "amortized_rec", "readout"))


@pytest.fixture(scope="session")
def lib():
    from stag_b200 import _lib
    return _lib.load()
