import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """Backend whose processors run on the CPU oracle (test infrastructure)."""
    from tests.backends import oracle_backend

    return oracle_backend()


@pytest.fixture(scope="session")
def product():
    """Backend whose processors run through libomb200.so (CUDA). GPU tests only."""
    from tests.backends import product_backend

    return product_backend()


@pytest.fixture(scope="session")
def emu():
    """Kernel sources under the CPU emulator (tests/emu) — development aid, not a product path."""
    from tests.backends import emu_backend

    return emu_backend()
