"""The C++ host mirror (airdos_b200/host/airdos_host.hpp) compiles against the C-ABI, links to the
shared library and, on a GPU, reproduces the Python path bit for bit."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "host_cpp", "_build", "test_host")


def _build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    lib_dir = os.path.join(ROOT, "airdos_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "host_cpp", "test_host.cpp"), "-o", EXE,
                           "-L" + lib_dir, "-lairdos_b200", "-Wl,-rpath," + lib_dir])


def test_host_mirror_compiles_links_and_refuses_without_gpu():
    import torch
    _build()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; covered by the gpu test")
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "NO_DEVICE" in out.stdout, out.stdout + out.stderr


def test_host_mirror_header_is_self_contained(tmp_path):
    """A translation unit that includes nothing but airdos_host.hpp compiles cleanly (C++14, the reference's standard is C++11 + lambdas the
    mirror does not need; -Wall -Wextra -Werror): every forwarder -- extractor, matcher (BestTwo, stereo, the guided searches, Fuse,
    distinctive descriptors), optimizer (local / global / human BA, PoseOptimization) -- type-checks against the C-ABI header."""
    src = tmp_path / "only_header.cpp"
    src.write_text('#include "airdos_b200/host/airdos_host.hpp"\nint main() { return 0; }\n')
    for std in ("c++14", "c++17"):
        out = subprocess.run(["g++", "-std=" + std, "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I" + ROOT, str(src)], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr


@pytest.mark.gpu
def test_host_mirror_matches_python_path(tmp_path):
    import airdos_b200 as adb
    from airdos_b200 import synth
    _build()
    img = synth.make_stereo_pair(0)[0]
    raw, outp = tmp_path / "img.raw", tmp_path / "out.bin"
    img.tofile(raw)
    out = subprocess.run([EXE, str(raw), str(outp)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
    buf = outp.read_bytes()
    n = int(np.frombuffer(buf[:4], np.int32)[0])
    kps = np.frombuffer(buf[4:4 + 24 * n], adb.KP_DTYPE)
    desc = np.frombuffer(buf[4 + 24 * n:4 + 56 * n], np.uint8).reshape(n, 32)
    ex = adb.ORBextractor(1000, 1.2, 8, 12, 7)
    k, d = ex(img)
    assert n == len(k) and kps.tobytes() == k.tobytes() and (desc == d).all()
    pose1 = np.frombuffer(buf[4 + 56 * n:], np.float64)
    assert np.allclose(pose1, [-0.5, 0, 0], atol=1e-3)
