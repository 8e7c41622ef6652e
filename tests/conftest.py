import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


def _gpu_available() -> bool:
    try:
        from airdos_b200 import capi
        return capi.lib().adb_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """Without a B200 (or without the built library) the gpu-marked tests are skipped instead of erroring in their fixtures:
    a plain `pytest tests` is green on a CPU box.  `-m gpu` on a box without a device still reports them as skipped, not passed."""
    if _gpu_available():
        return
    skip = pytest.mark.skip(reason="needs a B200 and airdos_b200/lib/libairdos_b200.so")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
