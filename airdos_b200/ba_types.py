"""ctypes mirror of adb_ba_problem / adb_ba_options / adb_ba_result (include/airdos_b200.h)."""
from __future__ import annotations

import ctypes as C

import numpy as np

_dp, _ip, _bp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)


class BAProblem(C.Structure):
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("bf", C.c_double),
                ("n_poses", C.c_int32), ("pose_q", _dp), ("pose_t", _dp), ("pose_fixed", _bp),
                ("n_points", C.c_int32), ("points", _dp),
                ("n_edges", C.c_int32), ("edge_pose", _ip), ("edge_point", _ip), ("edge_obs", _dp), ("edge_info", _dp),
                ("n_joints", C.c_int32), ("joints", _dp),
                ("n_joint_edges", C.c_int32), ("jedge_pose", _ip), ("jedge_joint", _ip), ("jedge_obs", _dp), ("jedge_info", _dp),
                ("n_dists", C.c_int32), ("dists", _dp),
                ("n_rigid_edges", C.c_int32), ("redge_i", _ip), ("redge_j", _ip), ("redge_dist", _ip), ("redge_info", _dp),
                ("n_motions", C.c_int32), ("motion_q", _dp), ("motion_t", _dp),
                ("n_motion_edges", C.c_int32), ("medge_p1", _ip), ("medge_p2", _ip), ("medge_motion", _ip), ("medge_dt", _dp),
                ("medge_info", _dp)]


class BAOptions(C.Structure):
    _fields_ = [("iterations", C.c_int32 * 2), ("max_trials", C.c_int32), ("tau", C.c_double),
                ("chi2_mono", C.c_double), ("chi2_stereo", C.c_double), ("chi2_rigid", C.c_double), ("chi2_motion", C.c_double),
                ("huber_mono", C.c_double), ("huber_stereo", C.c_double), ("huber_rigid", C.c_double), ("huber_motion", C.c_double),
                ("robust", C.c_int32 * 2)]


class BAResult(C.Structure):
    _fields_ = [("iterations_run", C.c_int32 * 2), ("trials_run", C.c_int32), ("stopped", C.c_int32),
                ("chi2_initial", C.c_double), ("chi2_round", C.c_double * 2), ("lambda_final", C.c_double),
                ("edge_outlier", _bp), ("jedge_outlier", _bp), ("redge_outlier", _bp), ("medge_outlier", _bp),
                ("edge_chi2", _dp), ("trace", _dp), ("trace_cap", C.c_int32), ("trace_len", C.c_int32)]


_F64 = ("pose_q", "pose_t", "points", "edge_obs", "edge_info", "joints", "jedge_obs", "jedge_info", "dists", "redge_info",
        "motion_q", "motion_t", "medge_dt", "medge_info")
_I32 = ("edge_pose", "edge_point", "jedge_pose", "jedge_joint", "redge_i", "redge_j", "redge_dist", "medge_p1", "medge_p2",
        "medge_motion")


class Problem:
    """Owns contiguous numpy copies of a problem dict (airdos_b200.synth.make_ba_problem layout) and
    the ctypes struct pointing at them.  In/out arrays are updated in place by a solve."""

    def __init__(self, d: dict):
        self.a = {}
        for k in _F64:
            self.a[k] = np.ascontiguousarray(d.get(k, np.zeros(0)), np.float64).copy()
        for k in _I32:
            self.a[k] = np.ascontiguousarray(d.get(k, np.zeros(0, np.int32)), np.int32).copy()
        self.a["pose_fixed"] = np.ascontiguousarray(d["pose_fixed"], np.uint8).copy()
        s = BAProblem()
        s.fx, s.fy, s.cx, s.cy, s.bf = d["fx"], d["fy"], d["cx"], d["cy"], d["bf"]
        s.n_poses = len(self.a["pose_t"].reshape(-1, 3)); s.n_points = len(self.a["points"].reshape(-1, 3))
        s.n_edges = len(self.a["edge_pose"]); s.n_joints = len(self.a["joints"].reshape(-1, 3))
        s.n_joint_edges = len(self.a["jedge_pose"]); s.n_dists = len(self.a["dists"])
        s.n_rigid_edges = len(self.a["redge_i"]); s.n_motions = len(self.a["motion_t"].reshape(-1, 3))
        s.n_motion_edges = len(self.a["medge_p1"])
        for k in _F64:
            setattr(s, k, self.a[k].ctypes.data_as(_dp))
        for k in _I32:
            setattr(s, k, self.a[k].ctypes.data_as(_ip))
        s.pose_fixed = self.a["pose_fixed"].ctypes.data_as(_bp)
        self.c = s

    def __getitem__(self, k):
        return self.a[k]


class Result:
    def __init__(self, p: Problem, trace_cap: int = 256):
        self.edge_outlier = np.zeros(p.c.n_edges, np.uint8); self.jedge_outlier = np.zeros(p.c.n_joint_edges, np.uint8)
        self.redge_outlier = np.zeros(p.c.n_rigid_edges, np.uint8); self.medge_outlier = np.zeros(p.c.n_motion_edges, np.uint8)
        self.edge_chi2 = np.zeros(p.c.n_edges, np.float64)
        self.trace = np.zeros((trace_cap, 5), np.float64)
        r = BAResult()
        r.edge_outlier = self.edge_outlier.ctypes.data_as(_bp); r.jedge_outlier = self.jedge_outlier.ctypes.data_as(_bp)
        r.redge_outlier = self.redge_outlier.ctypes.data_as(_bp); r.medge_outlier = self.medge_outlier.ctypes.data_as(_bp)
        r.edge_chi2 = self.edge_chi2.ctypes.data_as(_dp); r.trace = self.trace.ctypes.data_as(_dp)
        r.trace_cap = trace_cap
        self.c = r

    @property
    def trace_rows(self):
        return self.trace[:self.c.trace_len]


class PoseProblem(C.Structure):
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("bf", C.c_double),
                ("n_frames", C.c_int32), ("frame_ptr", _ip), ("pose_q", _dp), ("pose_t", _dp),
                ("xw", C.POINTER(C.c_float)), ("obs", C.POINTER(C.c_float)), ("inv_sigma2", C.POINTER(C.c_float)),
                ("outlier", _bp), ("n_inliers", _ip)]


class PoseBatch:
    """Owns the arrays of an adb_pose_problem: frames = list of dicts(pose_q, pose_t, xw[n,3], obs[n,3], inv_sigma2[n])."""

    def __init__(self, cam: dict, frames: list):
        n = [len(f["xw"]) for f in frames]
        self.frame_ptr = np.zeros(len(frames) + 1, np.int32); self.frame_ptr[1:] = np.cumsum(n)
        self.pose_q = np.ascontiguousarray([f["pose_q"] for f in frames], np.float64).reshape(-1, 4)
        self.pose_t = np.ascontiguousarray([f["pose_t"] for f in frames], np.float64).reshape(-1, 3)
        cat = lambda k, w: np.ascontiguousarray(np.concatenate([np.asarray(f[k], np.float32).reshape(-1, w) for f in frames]) if frames else np.zeros((0, w)), np.float32)
        self.xw, self.obs, self.inv_sigma2 = cat("xw", 3), cat("obs", 3), cat("inv_sigma2", 1).reshape(-1)
        self.outlier = np.zeros(max(int(self.frame_ptr[-1]), 1), np.uint8)
        self.n_inliers = np.zeros(len(frames), np.int32)
        s = PoseProblem()
        s.fx, s.fy, s.cx, s.cy, s.bf = cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["bf"]
        s.n_frames = len(frames)
        s.frame_ptr = self.frame_ptr.ctypes.data_as(_ip); s.pose_q = self.pose_q.ctypes.data_as(_dp); s.pose_t = self.pose_t.ctypes.data_as(_dp)
        fp = C.POINTER(C.c_float)
        s.xw = self.xw.ctypes.data_as(fp); s.obs = self.obs.ctypes.data_as(fp); s.inv_sigma2 = self.inv_sigma2.ctypes.data_as(fp)
        s.outlier = self.outlier.ctypes.data_as(_bp); s.n_inliers = self.n_inliers.ctypes.data_as(_ip)
        self.c = s


LEAF_RECORD = 122   # ADB_BA_LEAF_RECORD


class LeafIO(C.Structure):
    """adb_ba_leaf_io (include/airdos_b200.h): inputs / outputs of the leaf-arithmetic pinning hook."""
    _fields_ = [("n", C.c_int32), ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("bf", C.c_double),
                ("pose_q", C.c_void_p), ("pose_t", C.c_void_p), ("x", C.c_void_p), ("obs", C.c_void_p), ("pose_update", C.c_void_p),
                ("joint_a", C.c_void_p), ("joint_b", C.c_void_p), ("bone", C.c_void_p),
                ("motion_q", C.c_void_p), ("motion_t", C.c_void_p), ("motion_dt", C.c_void_p), ("motion_update", C.c_void_p), ("out", C.c_void_p),
                ("huber_delta", C.c_void_p), ("huber_e2", C.c_void_p)]
