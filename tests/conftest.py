import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def _have_gpu() -> bool:
    try:
        import intool_rag_b200  # noqa: F401
        from intool_rag_b200 import _lib
        return _lib.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    """GPU tests must FAIL, not skip, when selected on a box without a usable device or library:
    a silent skip would hide a missing CUDA path."""
    import intool_rag_b200  # noqa: F401
    from intool_rag_b200 import _lib
    _lib.require_gpu()
    return 0
