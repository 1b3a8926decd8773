import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def nkbk_lib():
    """The built C-ABI library; building it needs nvcc but no GPU."""
    from nkb_classification_b200 import build as _build
    _build.build()
    from nkb_classification_b200 import _lib
    return _lib.lib()


@pytest.fixture(scope="session")
def cuda_device(nkbk_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu but no CUDA device is visible")
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _seed_everything():
    """Every test starts from the same RNG state (model initialisation, shuffles, augmentation draws), so a result
    never depends on which tests ran before it."""
    import random

    import numpy as np
    random.seed(1234)
    np.random.seed(1234)
    try:
        import torch
        torch.manual_seed(1234)
    except Exception:  # pragma: no cover
        pass
    yield
