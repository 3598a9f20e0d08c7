import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_devices():
    try:
        from latticeurbanwind_b200 import _cabi
        return _cabi.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a CUDA device skips the `gpu` tests instead of failing in them (the product has no CPU fallback)."""
    if not any("gpu" in item.keywords for item in items) or _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device: `gpu` tests run on the B200 box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle as O
    O.build()
    return O
