import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    out = {}
    for name in ("calib", "result_2d", "result_3d", "template"):
        with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
            out[name] = {k: z[k] for k in z.files}
    return out


@pytest.fixture(scope="session")
def lib_built():
    """Make sure the in-tree shared library exists (cross-compiles without a GPU)."""
    from deepfly3d_b200 import build

    return build.build()
