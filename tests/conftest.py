import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (its name starts with a digit, hence import_module).  The CUDA library is a
    git-ignored build artefact: on a fresh checkout with nvcc present it is compiled here once (this
    builds the product, it is not a fallback -- without the library every op still fails loudly)."""
    import shutil
    p = importlib.import_module("3dvlp_b200")
    if not os.path.exists(p._lib.LIB_PATH) and shutil.which("nvcc"):
        importlib.import_module("3dvlp_b200.build").build()
    return p


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_layers.npz"))


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
