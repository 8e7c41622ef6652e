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
    from airdos_b200 import ba
    # every way into the library fails loudly without a device: the lazily provisioned and the sized extractor, the matcher, the
    # optimizer and the stand-alone dense solve -- none of them computes anything on the host
    for make in (lambda: adb.ORBextractor(1000, 1.2, 8, 12, 7), lambda: adb.ORBextractor(1000, 1.2, 8, 12, 7, 640, 480),
                 lambda: adb.ORBmatcher(), lambda: ba.Optimizer(), lambda: ba.dense_solve(np.eye(4), np.ones(4))):
        with pytest.raises(adb.AdbError) as e:
            make()
        assert e.value.status == 2   # ADB_ERR_NO_DEVICE
        assert "no CUDA device" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "airdos_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                for bad in ("import oracle", "from oracle", "liborb_oracle", "libba_oracle", "oracle/_build", "oracle/_ref"):
                    assert bad not in src, (f, bad)
                assert not re.search(r'#include\s*"[^"]*oracle', src), f


def test_ctypes_structs_match_the_header(tmp_path):
    """Every ctypes mirror has the size and field offsets the C compiler gives the header's struct (a drifted mirror would
    corrupt memory silently: the library reads the caller's struct)."""
    import ctypes as C
    import subprocess
    from airdos_b200 import ba_types as T, capi
    pairs = {"adb_orb_config": capi.OrbConfig, "adb_gather_targets": capi.GatherTargets, "adb_proj_search": capi.ProjSearch,
             "adb_bow_search": capi.BowSearch, "adb_ba_options": T.BAOptions, "adb_ba_problem": T.BAProblem,
             "adb_ba_result": T.BAResult, "adb_pose_problem": T.PoseProblem}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "airdos_b200.h"', 'int main(void) {']
    for cname, ct in pairs.items():
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, *_ in ct._fields_:
            lines.append(f'  printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"; exe = tmp_path / "layout"
    src.write_text("\n".join(lines))
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    assert len(out) == len(pairs)
    for line in out:
        f = line.split()
        ct = pairs[f[0]]
        assert int(f[1]) == C.sizeof(ct), (f[0], f[1], C.sizeof(ct))
        for (fname, *_), off in zip(ct._fields_, f[2:]):
            assert getattr(ct, fname).offset == int(off), (f[0], fname)
