"""ctypes binding of libairdos_b200.so (the C-ABI in include/airdos_b200.h).

This is the only way Python reaches the CUDA path; there is no fallback.  If the shared
library is missing the import raises, and on a machine without a B200 every ``*_create``
returns ADB_ERR_NO_DEVICE, which is surfaced as :class:`AdbError`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libairdos_b200.so")

OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_CAPACITY, ERR_NOT_POSDEF, ERR_STOPPED = range(7)

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4")])
assert KP_DTYPE.itemsize == 24


class AdbError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"libairdos_b200 status {status}: {msg}")
        self.status = status


class GatherTargets(C.Structure):
    _fields_ = [("n", C.c_int32), ("multicast", C.c_int32), ("kps", C.c_void_p * 8), ("desc", C.c_void_p * 8), ("counts", C.c_void_p * 8)]


class OrbConfig(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32), ("width", C.c_int32),
                ("height", C.c_int32), ("max_batch", C.c_int32), ("device", C.c_int32)]


class ProjSearch(C.Structure):
    """adb_proj_search (include/airdos_b200.h)."""
    _fields_ = [("n_kp", C.c_int32), ("kps", C.c_void_p), ("u_right", C.c_void_p), ("desc", C.c_void_p), ("taken", C.c_void_p),
                ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float),
                ("grid_inv_w", C.c_float), ("grid_inv_h", C.c_float),
                ("n_q", C.c_int32), ("q_u", C.c_void_p), ("q_v", C.c_void_p), ("q_ur", C.c_void_p), ("q_radius", C.c_void_p),
                ("q_min_level", C.c_void_p), ("q_max_level", C.c_void_p), ("q_flags", C.c_void_p), ("q_desc", C.c_void_p),
                ("q_angle", C.c_void_p), ("use_ratio", C.c_int32), ("nn_ratio", C.c_float), ("check_orientation", C.c_int32),
                ("last_xw", C.c_void_p), ("last_octave", C.c_void_p), ("tcw_cur", C.c_void_p), ("tcw_last", C.c_void_p),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("mbf", C.c_float), ("mb", C.c_float),
                ("scale_factors", C.c_void_p), ("n_levels", C.c_int32), ("th", C.c_float), ("mono", C.c_int32),
                ("mp_xw", C.c_void_p), ("mp_normal", C.c_void_p), ("mp_min_distance", C.c_void_p), ("mp_max_distance", C.c_void_p),
                ("ow", C.c_void_p), ("view_cos_limit", C.c_float), ("log_scale_factor", C.c_float),
                ("q_track", C.c_void_p), ("q_level", C.c_void_p), ("fuse", C.c_int32), ("inv_level_sigma2", C.c_void_p),
                ("kp_match", C.c_void_p), ("q_best_idx", C.c_void_p), ("q_best_dist", C.c_void_p), ("n_matches", C.c_int32)]


class BowSearch(C.Structure):
    """adb_bow_search (include/airdos_b200.h)."""
    _fields_ = [("mode", C.c_int32), ("n1", C.c_int32), ("kps1", C.c_void_p), ("u_right1", C.c_void_p), ("desc1", C.c_void_p), ("flags1", C.c_void_p),
                ("n2", C.c_int32), ("kps2", C.c_void_p), ("u_right2", C.c_void_p), ("desc2", C.c_void_p), ("flags2", C.c_void_p),
                ("n_buckets", C.c_int32), ("b_ptr1", C.c_void_p), ("b_idx1", C.c_void_p), ("b_ptr2", C.c_void_p), ("b_idx2", C.c_void_p),
                ("nn_ratio", C.c_float), ("check_orientation", C.c_int32), ("f12", C.c_void_p), ("ex", C.c_float), ("ey", C.c_float),
                ("scale_factors2", C.c_void_p), ("level_sigma2_2", C.c_void_p), ("n_levels", C.c_int32),
                ("match21", C.c_void_p), ("match12", C.c_void_p), ("n_matches", C.c_int32)]


# every symbol include/airdos_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _f32, _sz = C.c_void_p, C.c_int32, C.c_float, C.c_size_t
_ip = C.POINTER(C.c_int32)
_fp = C.POINTER(C.c_float)
SYMBOLS = {
    "adb_last_error": (C.c_char_p, []),
    "adb_version": (C.c_int, []),
    "adb_device_count": (C.c_int, []),
    "adb_orb_create": (C.c_int, [C.POINTER(OrbConfig), C.POINTER(_vp)]),
    "adb_orb_destroy": (C.c_int, [_vp]),
    "adb_orb_levels": (_i32, [_vp]),
    "adb_orb_capacity": (_i32, [_vp]),
    "adb_orb_level_info": (C.c_int, [_vp, _i32, _ip, _ip, _ip, _fp, _fp, _fp, _fp, _ip]),
    "adb_orb_extract": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _i32, _ip]),
    "adb_orb_extract_batch": (C.c_int, [_vp, _i32, _vp, _sz, _i32, _i32, _i32, _vp, _sz, _i32, _vp, _vp, _i32, _vp]),
    "adb_orb_extract_batch_device": (C.c_int, [_vp, _i32, _vp, _sz, _i32, _i32, _i32, _vp, _sz, _i32]),
    "adb_orb_sync": (C.c_int, [_vp]),
    "adb_orb_stream": (_vp, [_vp]),
    "adb_orb_set_gather": (C.c_int, [_vp, _vp]),
    "adb_orb_profile": (C.c_int, [_vp, _i32]),
    "adb_orb_stage_ms": (C.c_int, [_vp, _fp]),
    "adb_orb_launch_count": (C.c_int64, [_vp]),
    "adb_orb_results_device": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _ip]),
    "adb_orb_download": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _i32, _vp]),
    "adb_orb_get_pyramid": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _i32]),
    "adb_orb_debug_candidates": (C.c_int, [_vp, _i32, _i32, _vp, _i32, _ip]),
    "adb_hamming_distance": (_i32, [_vp, _vp]),
    "adb_matcher_create": (C.c_int, [_i32, C.POINTER(_vp)]),
    "adb_matcher_destroy": (C.c_int, [_vp]),
    "adb_match_best2": (C.c_int, [_vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "adb_match_best2_device": (C.c_int, [_vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "adb_search_by_projection": (C.c_int, [_vp, _vp, _i32]),
    "adb_search_last_ms": (C.c_int, [_vp, _fp]),
    "adb_search_by_bow": (C.c_int, [_vp, _vp, _i32]),
    "adb_distinctive_descriptors": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _vp]),
    "adb_stereo_match": (C.c_int, [_vp, _vp, _i32, _f32, _f32, _vp, _vp, _vp, _vp, _i32]),
    "adb_stereo_frames_batch": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _sz, _i32, _i32, _i32, _vp, _vp, _sz, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32,
                                          _f32, _f32, _vp, _vp, _vp, _vp]),
    "adb_stereo_match_device": (C.c_int, [_vp, _vp, _i32, _f32, _f32]),
    "adb_stereo_results_device": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "adb_ba_default_options": (None, [_vp]),
    "adb_ba_global_options": (None, [_vp, _i32, _i32]),
    "adb_ba_pose_from_tcw": (None, [_vp, _vp, _vp]),
    "adb_ba_pose_to_tcw": (None, [_vp, _vp, _vp]),
    "adb_ba_create": (C.c_int, [_i32, C.POINTER(_vp)]),
    "adb_ba_destroy": (C.c_int, [_vp]),
    "adb_ba_solve": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "adb_ba_stage_ms": (C.c_int, [_vp, _fp]),
    "adb_ba_launch_count": (C.c_int64, [_vp]),
    "adb_pose_optimize": (C.c_int, [_vp, _vp]),
    "adb_ba_leaf_eval": (C.c_int, [_vp, _vp]),
    "adb_dense_solve": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _ip, _i32, _i32, _fp]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the shared library (raises if it has not been built: run ``make -C airdos_b200/csrc``
    or ``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: the CUDA extension has not been built, and there is no CPU fallback")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)   # AttributeError if the library lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(status: int) -> None:
    if status != OK:
        raise AdbError(status, lib().adb_last_error().decode(errors="replace"))


def ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))
