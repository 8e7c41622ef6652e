"""Host-side mirror of ORB_SLAM2::Optimizer's bundle-adjustment entry points (include/Optimizer.h:42-67)
over the C-ABI.  The Map / KeyFrame / MapPoint objects of the reference become one flat problem
dict (layout of adb_ba_problem; airdos_b200.synth.make_ba_problem builds synthetic ones)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import ba_types as T
from .capi import check, lib, ptr, ERR_STOPPED, AdbError


def default_options() -> T.BAOptions:
    o = T.BAOptions()
    lib().adb_ba_default_options(C.byref(o))
    return o


def global_options(n_iterations: int = 5, robust: bool = True) -> T.BAOptions:
    """Constants of Optimizer::BundleAdjustment / GlobalBundleAdjustemnt (src/Optimizer.cc:52-230)."""
    o = T.BAOptions()
    lib().adb_ba_global_options(C.byref(o), n_iterations, int(robust))
    return o


def pose_from_tcw(tcw: np.ndarray):
    """Converter::toSE3Quat on a 4x4 float32 Tcw -> (q[x,y,z,w], t) float64."""
    tcw = np.ascontiguousarray(tcw, np.float32).reshape(16)
    q, t = np.zeros(4), np.zeros(3)
    lib().adb_ba_pose_from_tcw(ptr(tcw), ptr(q), ptr(t))
    return q, t


def pose_to_tcw(q: np.ndarray, t: np.ndarray) -> np.ndarray:
    """Converter::toCvMat(SE3Quat): float32 4x4."""
    out = np.zeros(16, np.float32)
    lib().adb_ba_pose_to_tcw(ptr(np.ascontiguousarray(q, np.float64)), ptr(np.ascontiguousarray(t, np.float64)), ptr(out))
    return out.reshape(4, 4)


def dense_solve(a: np.ndarray, b: np.ndarray, device: int = 0, cluster: int = 0, reps: int = 1):
    """g2o::LinearSolverDense::solve (linear_solver_dense.h:64-113) on the device: x with A x = b for SPD A (lower triangle read).
    Returns (x, info, ms_per_solve): info != 0 = not positive definite (g2o returns false and LM rejects the step)."""
    a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
    n = a.shape[0]
    assert a.shape == (n, n) and b.shape == (n,)
    x = np.zeros(n); info = C.c_int32(0); ms = C.c_float(0)
    check(lib().adb_dense_solve(device, n, ptr(a), ptr(b), ptr(x), C.byref(info), cluster, reps, C.byref(ms)))
    return x, int(info.value), float(ms.value)


class Optimizer:
    """Optimizer::LocalBundleAdjustment / LocalBundleAdjustmentHumanTrajactory on one GPU."""

    def __init__(self, device: int = 0):
        self._s = C.c_void_p()
        check(lib().adb_ba_create(device, C.byref(self._s)))

    def close(self):
        if getattr(self, "_s", None) is not None and self._s:
            lib().adb_ba_destroy(self._s)
            self._s = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def LocalBundleAdjustment(self, problem: dict, pbStopFlag: np.ndarray | None = None, options: T.BAOptions | None = None,
                              trace_cap: int = 256):
        """Runs the two-round LM schedule on a copy of `problem`.  Returns (Problem with the optimised
        state, Result, status): status 6 (ADB_ERR_STOPPED) = stop flag was already set, nothing written."""
        p = T.Problem(problem)
        r = T.Result(p, trace_cap)
        o = options or default_options()
        st = lib().adb_ba_solve(self._s, C.byref(p.c), C.byref(o), ptr(pbStopFlag) if pbStopFlag is not None else None, C.byref(r.c))
        if st not in (0, ERR_STOPPED):
            raise AdbError(st, lib().adb_last_error().decode(errors="replace"))
        return p, r, st

    def GlobalBundleAdjustemnt(self, problem: dict, nIterations: int = 5, pbStopFlag: np.ndarray | None = None, bRobust: bool = True):
        """Optimizer::GlobalBundleAdjustemnt / BundleAdjustment (src/Optimizer.cc:52-230; the reference's spelling): every
        key-frame a pose (id 0 fixed), every map point marginalised, one round of nIterations."""
        return self.LocalBundleAdjustment(problem, pbStopFlag, global_options(nIterations, bRobust))

    BundleAdjustment = GlobalBundleAdjustemnt

    # the dynamic window is the same call: the human arrays of the problem dict switch it on
    LocalBundleAdjustmentHumanTrajactory = LocalBundleAdjustment

    def PoseOptimization(self, cam: dict, frames: list):
        """Optimizer::PoseOptimization for a batch of frames (dicts with pose_q, pose_t, xw, obs, inv_sigma2).
        Returns a PoseBatch whose pose_q / pose_t / outlier / n_inliers hold the results."""
        pb = T.PoseBatch(cam, frames)
        check(lib().adb_pose_optimize(self._s, C.byref(pb.c)))
        return pb

    def stage_ms(self):
        ms = (C.c_float * 6)()
        check(lib().adb_ba_stage_ms(self._s, ms))
        return dict(zip(("linearize", "schur", "reduced_solve", "backsub_eval", "other", "lm_loop"), ms))

    def launch_count(self) -> int:
        return int(lib().adb_ba_launch_count(self._s))

