import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run with -m gpu on the B200 box")


@pytest.fixture(scope="session")
def built_lib():
    """The C-ABI shared library; (re)built with nvcc when missing or stale (cross-compiles without a GPU)."""
    from polyphemus_b200 import build

    if os.path.exists(build.LIB_PATH) and os.environ.get("PB_SKIP_BUILD") == "1":
        return build.LIB_PATH
    try:
        return build.build()
    except RuntimeError:
        if os.path.exists(build.LIB_PATH):
            return build.LIB_PATH
        raise


@pytest.fixture(scope="session")
def cuda(built_lib):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


def golden(name):
    import numpy as np

    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
