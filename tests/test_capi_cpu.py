"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol the header
declares, and refuses to run without a GPU instead of falling back to anything."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    hdr = open(os.path.join(ROOT, "include", "airdos_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(adb_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from airdos_b200 import capi
    lib = capi.lib()
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"libairdos_b200.so lacks {s}"
    # and the ctypes table binds exactly the header's entry points
    assert sorted(capi.SYMBOLS) == syms
    assert lib.adb_version() == 100


def test_hamming_distance_host_helper():
    from airdos_b200 import ORBmatcher
    rng = np.random.default_rng(0)
    a, b = rng.integers(0, 256, (2, 32), dtype=np.uint8)
    assert ORBmatcher.DescriptorDistance(a, b) == int(np.unpackbits(a ^ b).sum())


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import airdos_b200 as adb
    with pytest.raises(adb.AdbError) as e:
        adb.ORBextractor(1000, 1.2, 8, 12, 7)
    assert e.value.status == 2   # ADB_ERR_NO_DEVICE


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "airdos_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                for bad in ("import oracle", "from oracle", "liborb_oracle", "libba_oracle", "oracle/_build", "oracle/_ref"):
                    assert bad not in src, (f, bad)
                assert not re.search(r'#include\s*"[^"]*oracle', src), f
