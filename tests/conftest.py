import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def golden_models():
    return sorted(os.path.basename(f) for f in glob.glob(os.path.join(GOLDEN, "model_*.npz")))


def golden_recursions():
    return sorted(os.path.basename(f) for f in glob.glob(os.path.join(GOLDEN, "recursions_*.npz")))


@pytest.fixture(scope="session")
def engine():
    """One GPU context for the whole -m gpu session (fails loudly without CUDA)."""
    import itrails_b200 as itb
    eng = itb.Engine(0)
    yield eng
    eng.close()
