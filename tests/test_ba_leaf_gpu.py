"""GPU: the CUDA implementations of the BA leaf arithmetic (device functions of airdos_b200/csrc/ba.cu, through the adb_ba_leaf_eval
hook of the C-ABI) against tests/golden/ba_leaf_ref.npz -- values computed by the REFERENCE's own g2o / AirDOS type sources
(oracle/ref_leaf.cpp, oracle/gen_ref_leaf_golden.py).  Tolerance 1e-11 relative: the device code multiplies by 1/z and 1/z^2 where
the reference divides (two roundings instead of one per term), everything else is the same formula."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-11


def _eval(opt, g, X, huber=None):
    from airdos_b200 import ba_types as T
    from airdos_b200.capi import check, lib, ptr
    n = len(X)
    keep = [np.ascontiguousarray(a, np.float64) for a in (g["pose_q"], g["pose_t"], X, g["obs"], g["pose_update"], g["joint_a"], g["joint_b"], g["bone"],
                                                          g["motion_q"], g["motion_t"], g["motion_dt"], g["motion_update"])]
    out = np.zeros((n, T.LEAF_RECORD))
    io = T.LeafIO()
    io.n = n
    io.fx, io.fy, io.cx, io.cy, io.bf = [float(v) for v in g["cam"]]
    for name, a in zip(("pose_q", "pose_t", "x", "obs", "pose_update", "joint_a", "joint_b", "bone", "motion_q", "motion_t", "motion_dt", "motion_update"), keep):
        setattr(io, name, a.ctypes.data)
    io.out = out.ctypes.data
    if huber is not None:
        hd, he = (np.ascontiguousarray(a, np.float64) for a in huber)
        assert len(hd) == n and len(he) == n
        io.huber_delta, io.huber_e2 = hd.ctypes.data, he.ctypes.data
    check(lib().adb_ba_leaf_eval(opt._s, C.byref(io)))
    return out


def _close(a, b, scale):
    err = np.abs(a - b) / scale
    assert (err <= RTOL).all(), float(err.max())


def test_cuda_leaf_arithmetic_matches_the_reference():
    from airdos_b200 import ba
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "ba_leaf_ref.npz")))
    opt = ba.Optimizer()
    o = _eval(opt, g, g["X"])
    obs_scale = np.abs(g["obs"]).max(1, keepdims=True) + 1
    # Edge(Stereo)SE3ProjectXYZ
    _close(o[:, 0:3], g["stereo_err"], obs_scale + np.abs(g["stereo_err"]))
    _close(o[:, 3:12], g["stereo_Ji"], np.abs(g["stereo_Ji"]).max(1, keepdims=True) + 1)
    _close(o[:, 12:30], g["stereo_Jj"], np.abs(g["stereo_Jj"]).max(1, keepdims=True) + 1)
    _close(o[:, 30:32], g["mono_err"], obs_scale + np.abs(g["mono_err"]))
    _close(o[:, 33:39], g["mono_Ji"], np.abs(g["mono_Ji"]).max(1, keepdims=True) + 1)
    _close(o[:, 42:54], g["mono_Jj"], np.abs(g["mono_Jj"]).max(1, keepdims=True) + 1)
    assert (o[:, 39:42] == 0).all() and (o[:, 54:60] == 0).all()          # third row of a monocular edge is empty
    # vertices
    _close(o[:, 102:106], g["pose_oplus_q"], 1.0)
    _close(o[:, 106:109], g["pose_oplus_t"], np.abs(g["pose_oplus_t"]).max(1, keepdims=True) + 1)
    # articulated edges
    _close(o[:, 109], g["rigid_err"], np.abs(g["bone"]) + 1)
    _close(o[:, 110:113], g["motion_err"], np.abs(g["joint_a"]).max(1, keepdims=True) + 1)
    q = o[:, 113:117]
    x, y, z, w = q.T
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                  2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1)
    assert np.abs(R - g["motion_oplus_R"]).max() < 1e-10
    _close(o[:, 117:120], g["motion_oplus_t"], np.abs(g["motion_oplus_t"]).max(1, keepdims=True) + 1)
    # Edge(Stereo)SE3ProjectXYZOnlyPose take float map-point positions
    o = _eval(opt, g, g["Xf"])
    _close(o[:, 60:63], g["onlypose_stereo_err"], obs_scale + np.abs(g["onlypose_stereo_err"]))
    _close(o[:, 63:81], g["onlypose_stereo_J"], np.abs(g["onlypose_stereo_J"]).max(1, keepdims=True) + 1)
    _close(o[:, 81:83], g["onlypose_mono_err"], obs_scale + np.abs(g["onlypose_mono_err"]))
    _close(o[:, 84:96], g["onlypose_mono_J"], np.abs(g["onlypose_mono_J"]).max(1, keepdims=True) + 1)
    opt.close()


def test_cuda_huber_kernel_matches_the_reference_function():
    """The device Huber weight against RobustKernelHuber::robustify of the reference (Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:65-92,
    compiled from /root/reference by oracle/ref_lm.cpp; tests/golden/lm_ref.npz): rho(e2) and rho'(e2) for the deltas the Optimizer sets and
    squared errors on both sides of delta^2 -- including the values between the double delta^2 and the FLOAT `dsqr` member the reference
    compares with (core/robust_kernel_impl.h:84), where a double dsqr would classify differently."""
    from airdos_b200 import ba
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "ba_leaf_ref.npz")))
    lm = np.load(os.path.join(ROOT, "tests", "golden", "lm_ref.npz"))
    hd, he, rho = lm["huber_delta"], lm["huber_e2"], lm["huber_rho"]
    n = len(g["pose_q"])
    opt = ba.Optimizer()
    for a in range(0, len(hd), n):
        m = min(n, len(hd) - a)
        d = np.ones(n); e = np.ones(n); d[:m] = hd[a:a + m]; e[:m] = he[a:a + m]
        o = _eval(opt, g, g["X"], (d, e))
        assert (o[:m, 120] == rho[a:a + m, 0]).all() and (o[:m, 121] == rho[a:a + m, 1]).all()
    opt.close()
