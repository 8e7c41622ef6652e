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


def _nvidia_device_present() -> bool:
    """Is there an NVIDIA GPU on this box, whatever our library says?  (device nodes / the driver's proc tree; no torch import)"""
    if os.path.exists("/dev/nvidia0"):
        return True
    try:
        return len(os.listdir("/proc/driver/nvidia/gpus")) > 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    """On a box WITHOUT a GPU the gpu-marked tests are skipped instead of erroring in their fixtures: a plain `pytest tests` is green on a
    CPU box (`-m gpu` there reports them as skipped, not passed).  On a box WITH a GPU nothing is ever skipped: if libairdos_b200.so is
    missing, does not load or sees no device, the gpu tests run and fail loudly in their first library call -- a GPU box must never turn
    green without the CUDA path."""
    if _gpu_available() or _nvidia_device_present():
        return
    skip = pytest.mark.skip(reason="needs a B200 and airdos_b200/lib/libairdos_b200.so")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
