"""pytest wiring: the package directory plays the role of the reference's ``src/`` (reference tests/conftest.py:9-12)."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "flashattention-pytorch_b200"
for p in (str(PKG), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)

torch = pytest.importorskip("torch")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def device():
    return "cuda" if torch.cuda.is_available() else "cpu"


@pytest.fixture(scope="session")
def cuda_extension_available():
    """Unlike the reference (tests/conftest.py:31-41) a missing extension on a GPU box is an ERROR, not a skip."""
    if not torch.cuda.is_available():
        return False
    from fa1.cuda.impl import _load_ext

    _load_ext().load_library()
    return True
