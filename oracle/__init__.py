"""ctypes bindings for the CPU oracle (TEST INFRASTRUCTURE -- see oracle/orb_oracle.cpp header).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  The product package ``airdos_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4")])
assert KP_DTYPE.itemsize == 24


def build(force: bool = False) -> None:
    """Compile the oracle shared libraries with the committed Makefile."""
    if force:
        subprocess.check_call(["make", "-C", _HERE, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", _HERE, "-j4"], stdout=subprocess.DEVNULL)


_libs = {}


def _load(name: str) -> C.CDLL:
    if name not in _libs:
        path = os.path.join(_BUILD, name)
        if not os.path.exists(path):
            build()
        _libs[name] = C.CDLL(path)
    return _libs[name]


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------------------------------
# ORB extractor oracle
def orb_lib() -> C.CDLL:
    lib = _load("liborb_oracle.so")
    if not getattr(lib, "_typed", False):
        lib.orb_oracle_fast_atan2.restype = C.c_float
        lib.orb_oracle_fast_atan2.argtypes = [C.c_float, C.c_float]
        lib.orb_oracle_extract_batch.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                                 C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                                 C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        lib.orb_oracle_extract.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                           C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.orb_oracle_params.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5
        lib.orb_oracle_tables.argtypes = [C.c_int, C.c_float, C.c_int] + [C.c_void_p] * 7
        lib.orb_oracle_orient_describe.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib._typed = True
    return lib


def border101(src: np.ndarray, b: int = 19) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.zeros((h + 2 * b, w + 2 * b), np.uint8)
    orb_lib().orb_oracle_border101(_p(src), w, h, w, _p(dst), w + 2 * b, b)
    return dst


def erode10(src: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.zeros_like(src)
    orb_lib().orb_oracle_erode10(_p(src), w, h, w, _p(dst), w)
    return dst


def resize(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.zeros((dh, dw), np.uint8)
    orb_lib().orb_oracle_resize(_p(src), w, h, w, _p(dst), dw, dh, dw)
    return dst


def blur7(src: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.zeros_like(src)
    orb_lib().orb_oracle_blur7(_p(src), w, h, w, _p(dst), w)
    return dst


def fast_atan2(y: float, x: float) -> float:
    return float(orb_lib().orb_oracle_fast_atan2(float(y), float(x)))


def fast(img: np.ndarray, threshold: int, mask: np.ndarray | None = None) -> np.ndarray:
    """cv::FastFeatureDetector(threshold, True).detect(img, mask) -> int32 [n, 3] (x, y, score)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = max(16, w * h // 4)
    out = np.zeros((cap, 3), np.int32)
    if mask is not None:
        mask = np.ascontiguousarray(mask, np.uint8)
    n = orb_lib().orb_oracle_fast(_p(img), w, h, w, int(threshold), _p(mask) if mask is not None else None,
                                  w, _p(out), cap)
    return out[:n].copy()


def distribute(cand: np.ndarray, min_x: int, max_x: int, min_y: int, max_y: int, n: int) -> np.ndarray:
    """DistributeOctTree on float32 [m, 3] (x, y, response) -> float32 [k, 3] in the reference's list order."""
    cand = np.ascontiguousarray(cand, np.float32)
    cap = max(len(cand), 1) + 8
    out = np.zeros((cap, 3), np.float32)
    k = orb_lib().orb_oracle_distribute(_p(cand), len(cand), min_x, max_x, min_y, max_y, int(n), _p(out), cap)
    return out[:k].copy()


def orb_params(nfeatures: int, scale: float, nlevels: int, w: int, h: int):
    lw = np.zeros(nlevels, np.int32); lh = np.zeros(nlevels, np.int32); q = np.zeros(nlevels, np.int32)
    sc = np.zeros(nlevels, np.float32); um = np.zeros(16, np.int32)
    orb_lib().orb_oracle_params(nfeatures, scale, nlevels, w, h, _p(lw), _p(lh), _p(q), _p(sc), _p(um))
    return {"w": lw, "h": lh, "quota": q, "scale": sc, "umax": um}


def orb_tables(nfeatures: int, scale: float, nlevels: int, lib=None, prefix: str = "orb_oracle"):
    """Constructor tables (src/ORBextractor.cc:411-472 + the pattern :151-409).  `lib` / `prefix` let the reference-built library
    (oracle/_ref/libref_orb.so, prefix "ref") be called through the same signature."""
    f = np.float32
    sc, isc, s2, is2 = (np.zeros(nlevels, f) for _ in range(4))
    q = np.zeros(nlevels, np.int32); um = np.zeros(16, np.int32); pat = np.zeros(1024, np.int32)
    fn = orb_lib().orb_oracle_tables if lib is None else getattr(lib, "ref_extractor_tables")
    fn.argtypes = [C.c_int, C.c_float, C.c_int] + [C.c_void_p] * 7
    fn(nfeatures, scale, nlevels, _p(sc), _p(isc), _p(s2), _p(is2), _p(q), _p(um), _p(pat))
    return {"scale": sc, "inv_scale": isc, "sigma2": s2, "inv_sigma2": is2, "quota": q, "umax": um, "pattern": pat.reshape(512, 2)}


def orient_describe(img: np.ndarray, blurred: np.ndarray, xy: np.ndarray, lib=None):
    """IC_Angle + computeOrbDescriptor (src/ORBextractor.cc:78-148) at level coordinates xy [n, 2] -> (angle [n] f32, desc [n, 32] u8)."""
    img = np.ascontiguousarray(img, np.uint8); blurred = np.ascontiguousarray(blurred, np.uint8); xy = np.ascontiguousarray(xy, np.float32)
    h, w = img.shape
    ang = np.zeros(len(xy), np.float32); desc = np.zeros((len(xy), 32), np.uint8)
    fn = orb_lib().orb_oracle_orient_describe if lib is None else getattr(lib, "ref_orient_describe")
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    fn(_p(img), _p(blurred), w, h, len(xy), _p(xy), _p(ang), _p(desc))
    return ang, desc


def orb_extract(img: np.ndarray, mask: np.ndarray | None = None, nfeatures: int = 1000, scale: float = 1.2,
                nlevels: int = 8, ini_th: int = 20, min_th: int = 7, want_pyramid: bool = False):
    """Full ORBextractor::operator() restatement.  Returns dict(kps, desc[, pyramid, cand_counts])."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = nfeatures + 8 * nlevels + 64
    kps = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    cand = np.zeros(nlevels, np.int32)
    pyr = None
    if want_pyramid:
        p = orb_params(nfeatures, scale, nlevels, w, h)
        pyr = np.zeros(int((p["w"].astype(np.int64) * p["h"]).sum()), np.uint8)
    if mask is not None:
        mask = np.ascontiguousarray(mask, np.uint8)
        assert mask.shape == img.shape
    n = orb_lib().orb_oracle_extract(_p(img), w, h, w, _p(mask) if mask is not None else None, w,
                                     nfeatures, scale, nlevels, ini_th, min_th, _p(kps), _p(desc), cap,
                                     _p(pyr) if pyr is not None else None, _p(cand))
    if n == -2:
        raise ValueError("image shape unsupported: a pyramid level is taller than twice its width (nIni = round(w / h) = 0, undefined in the reference)")
    if n < 0:
        raise RuntimeError("oracle key-point capacity exceeded")
    out = {"kps": kps[:n].copy(), "desc": desc[:n].copy(), "cand_counts": cand}
    if want_pyramid:
        levels, o = [], 0
        for l in range(nlevels):
            sz = int(p["w"][l]) * int(p["h"][l])
            levels.append(pyr[o:o + sz].reshape(int(p["h"][l]), int(p["w"][l])))
            o += sz
        out["pyramid"] = levels
    return out


def orb_extract_batch(imgs: np.ndarray, nfeatures: int, scale: float, nlevels: int, ini_th: int, min_th: int,
                      threads: int = 1):
    """imgs: u8 [F, H, W] contiguous.  Returns (kps [F, cap], desc [F, cap, 32], counts [F])."""
    imgs = np.ascontiguousarray(imgs, np.uint8)
    f, h, w = imgs.shape
    cap = nfeatures + 8 * nlevels + 64
    kps = np.zeros((f, cap), KP_DTYPE)
    desc = np.zeros((f, cap, 32), np.uint8)
    counts = np.zeros(f, np.int32)
    orb_lib().orb_oracle_extract_batch(_p(imgs), f, h * w, w, h, w, nfeatures, scale, nlevels, ini_th, min_th,
                                       _p(kps), _p(desc), cap, _p(counts), threads)
    return kps, desc, counts


# ------------------------------------------------------------------------------------------
# Matcher oracle (oracle/match_oracle.cpp)
def _match_lib() -> C.CDLL:
    lib = orb_lib()
    if not getattr(lib, "_mtyped", False):
        lib.match_oracle_distance.restype = C.c_int
        lib.match_oracle_distance.argtypes = [C.c_void_p, C.c_void_p]
        lib.match_oracle_best2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 5
        lib.match_oracle_stereo.argtypes = ([C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int] +
                                            [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int] +
                                            [C.c_void_p] * 4)
        lib._mtyped = True
    return lib


def hamming(a: np.ndarray, b: np.ndarray) -> int:
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return int(_match_lib().match_oracle_distance(_p(a), _p(b)))


def best2(q: np.ndarray, t: np.ndarray, cand_off: np.ndarray | None = None, cand_idx: np.ndarray | None = None):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    nq = len(q)
    bi = np.zeros(nq, np.int32); bd = np.zeros(nq, np.int32); sd = np.zeros(nq, np.int32)
    if cand_off is not None:
        cand_off = np.ascontiguousarray(cand_off, np.int32); cand_idx = np.ascontiguousarray(cand_idx, np.int32)
    _match_lib().match_oracle_best2(_p(q), nq, _p(t), len(t), _p(cand_off) if cand_off is not None else None,
                                    _p(cand_idx) if cand_idx is not None else None, _p(bi), _p(bd), _p(sd))
    return bi, bd, sd


def distinctive(desc: np.ndarray, point_ptr: np.ndarray) -> np.ndarray:
    """MapPoint::ComputeDistinctiveDescriptors restatement -> best index per point."""
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    point_ptr = np.ascontiguousarray(point_ptr, np.int32)
    out = np.zeros(len(point_ptr) - 1, np.int32)
    lib = _match_lib()
    lib.match_oracle_distinctive.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.match_oracle_distinctive(_p(desc), _p(point_ptr), len(out), _p(out))
    return out


def search_by_projection(pr: dict, nn_ratio: float = 0.6, check_orientation: bool = True):
    """SearchByProjection restatement on the problem dicts of airdos_b200.ORBmatcher.search_by_projection.
    Returns (nmatches, kp_match, q_best_idx, q_best_dist)."""
    lib = _match_lib()
    kps = np.ascontiguousarray(pr["kps"], KP_DTYPE)
    ur = np.ascontiguousarray(pr["u_right"], np.float32); desc = np.ascontiguousarray(pr["desc"], np.uint8)
    taken = np.ascontiguousarray(pr["taken"], np.uint8) if pr.get("taken") is not None else None
    mnx, mny, mxx, mxy = [np.float32(v) for v in pr["bounds"]]
    inv_w = np.float32(64) / np.float32(mxx - mnx); inv_h = np.float32(48) / np.float32(mxy - mny)
    qf = np.ascontiguousarray(pr["q_flags"], np.uint8); qd = np.ascontiguousarray(pr["q_desc"], np.uint8)
    nq, nk = len(qf), len(kps)
    extra = {}
    if pr.get("fuse"):
        fx, fy, cx, cy, mbf, mb = [float(v) for v in pr["cam"]]
        sf = np.ascontiguousarray(pr["scale_factors"], np.float32); isg = np.ascontiguousarray(pr["inv_level_sigma2"], np.float32)
        arrs = [np.ascontiguousarray(pr[k], np.float32) for k in ("tcw_cur", "ow", "mp_xw", "mp_normal", "mp_min_distance", "mp_max_distance")]
        km = np.zeros(nk, np.int32); bi = np.zeros(nq, np.int32); bd = np.zeros(nq, np.int32)
        lib.match_oracle_fuse_search.restype = C.c_int
        lib.match_oracle_fuse_search.argtypes = ([C.c_void_p] * 3 + [C.c_int] + [C.c_float] * 6 + [C.c_void_p, C.c_void_p, C.c_int] +
                                                 [C.c_void_p] * 6 + [C.c_float] * 6 + [C.c_void_p, C.c_void_p, C.c_int, C.c_float] + [C.c_void_p] * 3)
        n = lib.match_oracle_fuse_search(_p(kps), _p(ur), _p(desc), nk, mnx, mny, mxx, mxy, inv_w, inv_h, _p(arrs[0]), _p(arrs[1]), nq,
                                         _p(arrs[2]), _p(arrs[3]), _p(arrs[4]), _p(arrs[5]), _p(qf), _p(qd), fx, fy, cx, cy, mbf,
                                         float(pr["log_scale_factor"]), _p(sf), _p(isg), len(sf), float(pr["th"]), _p(km), _p(bi), _p(bd))
        return int(n), km, bi, bd, {}
    if "mp_xw" in pr:
        qu = np.zeros(nq, np.float32); qv = np.zeros(nq, np.float32); qur = np.zeros(nq, np.float32); qr = np.zeros(nq, np.float32)
        qmin = np.zeros(nq, np.int32); qmax = np.zeros(nq, np.int32); qfl = np.zeros(nq, np.uint8)
        track = np.zeros((nq, 4), np.float32); level = np.zeros(nq, np.int32)
        fx, fy, cx, cy, mbf, mb = [float(v) for v in pr["cam"]]
        sf = np.ascontiguousarray(pr["scale_factors"], np.float32)
        arrs = [np.ascontiguousarray(pr[k], np.float32) for k in ("tcw_cur", "ow", "mp_xw", "mp_normal", "mp_min_distance", "mp_max_distance")]
        lib.match_oracle_frustum.argtypes = ([C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_float] * 11 +
                                             [C.c_void_p, C.c_int, C.c_float] + [C.c_void_p] * 9)
        lib.match_oracle_frustum(_p(arrs[0]), _p(arrs[1]), nq, _p(arrs[2]), _p(arrs[3]), _p(arrs[4]), _p(arrs[5]), _p(qf), fx, fy, cx, cy,
                                 mbf, mnx, mxx, mny, mxy, float(pr.get("view_cos_limit", 0.5)), float(pr["log_scale_factor"]), _p(sf),
                                 len(sf), float(pr.get("th", 1.0)), _p(qu), _p(qv), _p(qur), _p(qr), _p(qmin), _p(qmax), _p(qfl),
                                 _p(track), _p(level))
        use_ratio, chk = int(pr.get("use_ratio", 1)), 0
        nn_ratio = float(pr.get("nn_ratio", nn_ratio))
        qa = np.zeros(nq, np.float32)
        extra = dict(q_track=track, q_level=level)
    elif "last_xw" in pr:
        qu = np.zeros(nq, np.float32); qv = np.zeros(nq, np.float32); qur = np.zeros(nq, np.float32); qr = np.zeros(nq, np.float32)
        qmin = np.zeros(nq, np.int32); qmax = np.zeros(nq, np.int32); qfl = np.zeros(nq, np.uint8)
        fx, fy, cx, cy, mbf, mb = [float(v) for v in pr["cam"]]
        sf = np.ascontiguousarray(pr["scale_factors"], np.float32)
        xw = np.ascontiguousarray(pr["last_xw"], np.float32); lo = np.ascontiguousarray(pr["last_octave"], np.int32)
        tc = np.ascontiguousarray(pr["tcw_cur"], np.float32); tl = np.ascontiguousarray(pr["tcw_last"], np.float32)
        lib.match_oracle_project_last.argtypes = ([C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] +
                                                  [C.c_float] * 10 + [C.c_void_p, C.c_float, C.c_int] + [C.c_void_p] * 7)
        lib.match_oracle_project_last(_p(tc), _p(tl), nq, _p(xw), _p(lo), _p(qf), fx, fy, cx, cy, mbf, mb, mnx, mxx, mny, mxy,
                                      _p(sf), float(pr["th"]), int(pr.get("mono", 0)), _p(qu), _p(qv), _p(qur), _p(qr),
                                      _p(qmin), _p(qmax), _p(qfl))
        use_ratio, chk = 0, int(pr.get("check_orientation", check_orientation))
        qa = np.ascontiguousarray(pr["q_angle"], np.float32)
    else:
        qu, qv, qur, qr = [np.ascontiguousarray(pr[k], np.float32) for k in ("q_u", "q_v", "q_ur", "q_radius")]
        qmin = np.ascontiguousarray(pr["q_min_level"], np.int32); qmax = np.ascontiguousarray(pr["q_max_level"], np.int32)
        qfl = qf
        use_ratio, chk = int(pr.get("use_ratio", 1)), int(pr.get("check_orientation", 0))
        nn_ratio = float(pr.get("nn_ratio", nn_ratio))
        qa = np.ascontiguousarray(pr["q_angle"], np.float32) if chk else np.zeros(nq, np.float32)
    km = np.zeros(nk, np.int32); bi = np.zeros(nq, np.int32); bd = np.zeros(nq, np.int32)
    lib.match_oracle_search_projection.restype = C.c_int
    lib.match_oracle_search_projection.argtypes = ([C.c_void_p] * 4 + [C.c_int] + [C.c_float] * 4 + [C.c_int] + [C.c_void_p] * 9 +
                                                   [C.c_int, C.c_float, C.c_int] + [C.c_void_p] * 3)
    n = lib.match_oracle_search_projection(_p(kps), _p(ur), _p(desc), _p(taken) if taken is not None else None, nk, mnx, mny,
                                           inv_w, inv_h, nq, _p(qu), _p(qv), _p(qur), _p(qr), _p(qmin), _p(qmax), _p(qd), _p(qfl),
                                           _p(qa), use_ratio, nn_ratio, chk, _p(km), _p(bi), _p(bd))
    return int(n), km, bi, bd, dict(q_u=qu, q_v=qv, q_ur=qur, q_radius=qr, q_min_level=qmin, q_max_level=qmax, q_flags=qfl, **extra)


def search_by_bow(pr: dict, nn_ratio: float = 0.6, check_orientation: bool = True):
    """SearchByBoW / SearchForTriangulation restatement on the problem dicts of ORBmatcher.search_by_bow -> (nmatches, match)."""
    lib = _match_lib()
    mode = int(pr["mode"])
    k1 = np.ascontiguousarray(pr["kps1"], KP_DTYPE); k2 = np.ascontiguousarray(pr["kps2"], KP_DTYPE)
    d1 = np.ascontiguousarray(pr["desc1"], np.uint8); d2 = np.ascontiguousarray(pr["desc2"], np.uint8)
    f1 = np.ascontiguousarray(pr["flags1"], np.uint8)
    n1, n2 = len(k1), len(k2)
    z = np.zeros(max(n1, n2, 1), np.float32); zf = np.ones(max(n2, 1), np.uint8)
    ur1 = np.ascontiguousarray(pr.get("u_right1", z[:n1]), np.float32); ur2 = np.ascontiguousarray(pr.get("u_right2", z[:n2]), np.float32)
    f2 = np.ascontiguousarray(pr.get("flags2", zf[:n2]), np.uint8)
    p1, i1, p2, i2 = [np.ascontiguousarray(pr[k], np.int32) for k in ("b_ptr1", "b_idx1", "b_ptr2", "b_idx2")]
    F = np.ascontiguousarray(pr.get("f12", np.zeros(9)), np.float32)
    ex, ey = [float(v) for v in pr.get("epipole", (0.0, 0.0))]
    sf = np.ascontiguousarray(pr.get("scale_factors2", np.ones(8)), np.float32)
    sg = np.ascontiguousarray(pr.get("level_sigma2_2", np.ones(8)), np.float32)
    m21 = np.zeros(max(n2, 1), np.int32); m12 = np.zeros(max(n1, 1), np.int32)
    lib.match_oracle_bow_search.restype = C.c_int
    lib.match_oracle_bow_search.argtypes = ([C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_int] + [C.c_void_p] * 4 +
                                            [C.c_float, C.c_int, C.c_void_p, C.c_float, C.c_float] + [C.c_void_p] * 4)
    n = lib.match_oracle_bow_search(mode, _p(k1), _p(ur1), _p(d1), _p(f1), n1, _p(k2), _p(ur2), _p(d2), _p(f2), n2, len(p1) - 1,
                                    _p(p1), _p(i1), _p(p2), _p(i2), float(pr.get("nn_ratio", nn_ratio)),
                                    int(pr.get("check_orientation", check_orientation)), _p(F), ex, ey, _p(sf), _p(sg), _p(m21), _p(m12))
    return int(n), (m21[:n2] if mode == 0 else m12[:n1])


def stereo_match(kl, dl, kr, dr, pyr_l, pyr_r, scale, mb: float, mbf: float, stage: int = 0):
    """Frame::ComputeStereoMatches restatement.  pyr_l / pyr_r: lists of level ROIs (u8 2-D).
    Returns (uRight, depth, ham_idx, ham_dist)."""
    kl = np.ascontiguousarray(kl, KP_DTYPE); kr = np.ascontiguousarray(kr, KP_DTYPE)
    dl = np.ascontiguousarray(dl, np.uint8); dr = np.ascontiguousarray(dr, np.uint8)
    nl = len(pyr_l)
    lw = np.array([p.shape[1] for p in pyr_l], np.int32); lh = np.array([p.shape[0] for p in pyr_l], np.int32)
    off = np.zeros(nl, np.int64)
    off[1:] = np.cumsum(lw.astype(np.int64) * lh)[:-1]
    pl = np.concatenate([np.ascontiguousarray(p, np.uint8).ravel() for p in pyr_l])
    pr = np.concatenate([np.ascontiguousarray(p, np.uint8).ravel() for p in pyr_r])
    sc = np.ascontiguousarray(scale, np.float32)
    inv = (np.float32(1.0) / sc).astype(np.float32)
    n = len(kl)
    ur = np.zeros(n, np.float32); dp = np.zeros(n, np.float32); hi = np.zeros(n, np.int32); hd = np.zeros(n, np.int32)
    _match_lib().match_oracle_stereo(_p(kl), _p(dl), n, _p(kr), _p(dr), len(kr), _p(pl), _p(pr), _p(off), _p(lw), _p(lh), nl,
                                     _p(sc), _p(inv), mb, mbf, stage, _p(ur), _p(dp), _p(hi), _p(hd))
    return ur, dp, hi, hd


def stereo_pipeline_batch(pairs: np.ndarray, nfeatures: int, scale: float, nlevels: int, ini_th: int, min_th: int,
                          mb: float, mbf: float, threads: int = 1):
    """ExtractORB(left) + ExtractORB(right) + ComputeStereoMatches for u8 [P, 2, H, W] pairs on
    `threads` host threads -> (counts [P, 2], matched [P], total key-points)."""
    pairs = np.ascontiguousarray(pairs, np.uint8)
    p, two, h, w = pairs.shape
    assert two == 2
    lib = orb_lib()
    lib.pipeline_oracle_stereo_batch.restype = C.c_long
    lib.pipeline_oracle_stereo_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                                 C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    counts = np.zeros((p, 2), np.int32); matched = np.zeros(p, np.int32)
    total = lib.pipeline_oracle_stereo_batch(_p(pairs), p, w, h, nfeatures, scale, nlevels, ini_th, min_th, mb, mbf, threads,
                                             _p(counts), _p(matched))
    return counts, matched, int(total)


# ------------------------------------------------------------------------------------------
# Bundle-adjustment oracle (oracle/ba_oracle.cpp)
def ba_lib() -> C.CDLL:
    from airdos_b200 import ba_types as T
    lib = _load("libba_oracle.so")
    if not getattr(lib, "_typed", False):
        lib.ba_oracle_solve.restype = C.c_int
        lib.ba_oracle_solve.argtypes = [C.POINTER(T.BAProblem), C.POINTER(T.BAOptions), C.c_void_p, C.POINTER(T.BAResult)]
        lib.ba_oracle_default_options.argtypes = [C.POINTER(T.BAOptions)]
        lib.ba_oracle_reproj.restype = C.c_int
        lib.ba_oracle_reproj.argtypes = [C.POINTER(T.BAProblem)] + [C.c_void_p] * 7
        lib.ba_oracle_first_step.restype = C.c_int
        lib.ba_oracle_first_step.argtypes = [C.POINTER(T.BAProblem), C.POINTER(T.BAOptions), C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        lib._typed = True
    return lib


def ba_default_options():
    from airdos_b200 import ba_types as T
    o = T.BAOptions()
    ba_lib().ba_oracle_default_options(C.byref(o))
    return o


def ba_global_options(n_iterations: int = 5, robust: bool = True):
    from airdos_b200 import ba_types as T
    o = T.BAOptions()
    lib = ba_lib()
    lib.ba_oracle_global_options.argtypes = [C.POINTER(T.BAOptions), C.c_int32, C.c_int32]
    lib.ba_oracle_global_options(C.byref(o), n_iterations, int(robust))
    return o


def ba_solve(problem_dict: dict, options=None, stop: np.ndarray | None = None, trace_cap: int = 256):
    """Two-round LM on a copy of `problem_dict` -> (Problem with updated state, Result, status)."""
    from airdos_b200 import ba_types as T
    p = T.Problem(problem_dict)
    r = T.Result(p, trace_cap)
    o = options or ba_default_options()
    st = ba_lib().ba_oracle_solve(C.byref(p.c), C.byref(o), _p(stop) if stop is not None else None, C.byref(r.c))
    return p, r, st


def pose_optimize(cam: dict, frames: list):
    """Optimizer::PoseOptimization restatement on a batch of frames -> PoseBatch with results."""
    from airdos_b200 import ba_types as T
    pb = T.PoseBatch(cam, frames)
    lib = ba_lib()
    lib.ba_oracle_pose_optimize.restype = C.c_int
    lib.ba_oracle_pose_optimize.argtypes = [C.POINTER(T.PoseProblem)]
    lib.ba_oracle_pose_optimize(C.byref(pb.c))
    return pb


# ------------------------------------------------------------------------------------------
# The oracle's LM steps one by one (ba_oracle_lm_*), and the hook table oracle/ref_lm.cpp drives them through
_LM_HOOKS = ("compute_errors", "chi2", "build", "layout", "vectors", "set_lambda", "solve", "update", "push", "pop", "discard_top")


class LmHooks(C.Structure):
    _fields_ = [("ctx", C.c_void_p)] + [(n, C.c_void_p) for n in _LM_HOOKS]


class LmSession:
    """One round of the oracle's solver (Solver of oracle/ba_oracle.cpp) on a copy of `problem_dict`, opened at the initial state."""

    def __init__(self, problem_dict: dict, options=None, robust: bool = True):
        from airdos_b200 import ba_types as T
        lib = ba_lib()
        lib.ba_oracle_lm_open.restype = C.c_void_p
        lib.ba_oracle_lm_open.argtypes = [C.POINTER(T.BAProblem), C.POINTER(T.BAOptions), C.c_int]
        lib.ba_oracle_lm_close.argtypes = [C.c_void_p]
        lib.ba_oracle_lm_optimize.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.ba_oracle_lm_state.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        self.lib = lib
        self.p = T.Problem(problem_dict)
        self.o = options or ba_default_options()
        self.h = lib.ba_oracle_lm_open(C.byref(self.p.c), C.byref(self.o), int(robust))

    def hooks(self) -> LmHooks:
        return LmHooks(self.h, *[C.cast(getattr(self.lib, "ba_oracle_lm_" + n), C.c_void_p) for n in _LM_HOOKS])

    def optimize(self, iterations: int, cap: int = 512):
        """The oracle's own loop (Solver::optimize) -> (iterations run, rows [(lambda, chi2 before, chi2 after, accepted)], final lambda)."""
        tr = np.zeros((cap, 5)); n = C.c_int(); lam = C.c_double()
        it = self.lib.ba_oracle_lm_optimize(self.h, iterations, _p(tr), cap, C.byref(n), C.byref(lam))
        return it, tr[:n.value][:, [0, 1, 2, 4]].copy(), lam.value

    def state(self) -> np.ndarray:
        n = self.lib.ba_oracle_lm_state(self.h, None, 0)
        a = np.zeros(n)
        self.lib.ba_oracle_lm_state(self.h, _p(a), n)
        return a

    def close(self):
        if self.h:
            self.lib.ba_oracle_lm_close(self.h); self.h = None

    # -- one trial step by step, and the system it solves (for the Schur-complement pin, oracle/ref_schur.cpp)
    def linearize(self):
        """computeActiveErrors + buildSystem at the current state."""
        self.lib.ba_oracle_lm_compute_errors.argtypes = [C.c_void_p]; self.lib.ba_oracle_lm_build.argtypes = [C.c_void_p]
        self.lib.ba_oracle_lm_compute_errors(self.h); self.lib.ba_oracle_lm_build(self.h)

    def system(self) -> dict:
        """The assembled normal equations of a static window in BlockSolver_6_3's block form (ba_oracle_lm_system)."""
        f = self.lib.ba_oracle_lm_system
        f.argtypes = [C.c_void_p] * 8
        sz = np.zeros(3, np.int32)
        if not f(self.h, _p(sz), None, None, None, None, None, None):
            raise ValueError("the window has articulated vertices")
        np_, nl, ne = (int(v) for v in sz)
        d = dict(n_poses=np_, n_points=nl, edge_pose=np.zeros(ne, np.int32), edge_point=np.zeros(ne, np.int32), W=np.zeros((ne, 6, 3)),
                 Hpp=np.zeros((np_, 6, 6)), Hll=np.zeros((nl, 3, 3)), b=np.zeros(6 * np_ + 3 * nl))
        f(self.h, _p(sz), _p(d["edge_pose"]), _p(d["edge_point"]), _p(d["W"]), _p(d["Hpp"]), _p(d["Hll"]), _p(d["b"]))
        return d

    def system_x(self) -> dict:
        """The same for any window (articulated ones included) in BlockSolverX's form: the non-marginalised vertices with their widths and
        the dense block H over them, the map points as landmarks (ba_oracle_lm_system_x)."""
        f = self.lib.ba_oracle_lm_system_x
        f.argtypes = [C.c_void_p] * 9
        sz = np.zeros(4, np.int32)
        f(self.h, _p(sz), None, None, None, None, None, None, None)
        nb, nl, ne, nd = (int(v) for v in sz)
        d = dict(n_points=nl, n_dense=nd, dims=np.zeros(nb, np.int32), edge_block=np.zeros(ne, np.int32), edge_point=np.zeros(ne, np.int32),
                 W=np.zeros((ne, 6, 3)), H=np.zeros((nd, nd)), Hll=np.zeros((nl, 3, 3)), b=np.zeros(nd + 3 * nl))
        f(self.h, _p(sz), _p(d["dims"]), _p(d["edge_block"]), _p(d["edge_point"]), _p(d["W"]), _p(d["H"]), _p(d["Hll"]), _p(d["b"]))
        return d

    def solve(self, lam: float):
        """Solver::setLambda + Solver::solve (the oracle's Schur complement, Cholesky and back-substitution) -> (ok, x of the whole system)."""
        self.lib.ba_oracle_lm_set_lambda.argtypes = [C.c_void_p, C.c_double]; self.lib.ba_oracle_lm_solve.argtypes = [C.c_void_p]
        self.lib.ba_oracle_lm_vectors.argtypes = [C.c_void_p] * 4 + [C.c_int]
        self.lib.ba_oracle_lm_set_lambda(self.h, float(lam))
        ok = bool(self.lib.ba_oracle_lm_solve(self.h))
        n = self.lib.ba_oracle_lm_vectors(self.h, None, None, None, 0)
        x = np.zeros(n)
        self.lib.ba_oracle_lm_vectors(self.h, _p(x), None, None, n)
        return ok, x


def ref_schur_solve(ref_lib, system: dict, lam: float):
    """The REFERENCE's BlockSolver<BlockSolverTraits<6, 3>>::solve() (oracle/_ref/libref_schur.so: core/block_solver.hpp:353-483 compiled
    from /root/reference) on `system` (LmSession.system()) with `lam` on both diagonals -> (ok, x, Hschur dense, bschur).  The linear
    solver behind the reduced system is a plain Cholesky, not the reference's LDLT."""
    npz, nl, ne = system["n_poses"], system["n_points"], len(system["edge_pose"])
    ref_lib.ref_schur_solve.argtypes = [C.c_int] * 3 + [C.c_void_p] * 6 + [C.c_double] + [C.c_void_p] * 3
    x = np.zeros(6 * npz + 3 * nl); hs = np.zeros((6 * npz, 6 * npz)); bs = np.zeros(6 * npz)
    keep = [np.ascontiguousarray(system[k]) for k in ("edge_pose", "edge_point", "W", "Hpp", "Hll", "b")]
    ok = ref_lib.ref_schur_solve(npz, nl, ne, *[_p(a) for a in keep], float(lam), _p(x), _p(hs), _p(bs))
    return bool(ok), x, hs, bs


def ref_schur_solve_x(ref_lib, system: dict, lam: float):
    """The same function instantiated as BlockSolverX (run-time block sizes: what LocalBundleAdjustmentHumanTrajactory uses) on
    LmSession.system_x() -> (ok, x, Hschur dense, bschur)."""
    nb, nl, ne, nd = len(system["dims"]), system["n_points"], len(system["edge_block"]), system["n_dense"]
    ref_lib.ref_schur_solve_x.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6 + [C.c_double] + [C.c_void_p] * 3
    x = np.zeros(nd + 3 * nl); hs = np.zeros((nd, nd)); bs = np.zeros(nd)
    keep = [np.ascontiguousarray(system[k]) for k in ("dims", "edge_block", "edge_point", "W", "H", "Hll", "b")]
    ok = ref_lib.ref_schur_solve_x(nb, _p(keep[0]), nl, ne, *[_p(a) for a in keep[1:]], float(lam), _p(x), _p(hs), _p(bs))
    return bool(ok), x, hs, bs


def ref_lm_optimize(ref_lib, session: LmSession, iterations: int, cap: int = 512):
    """The REFERENCE's SparseOptimizer::optimize + OptimizationAlgorithmLevenberg::solve (oracle/_ref/libref_lm.so) over the session's
    arithmetic -> (iterations run, rows, final lambda, number of computeActiveErrors calls, the constructor's tau)."""
    ref_lib.ref_lm_optimize.argtypes = [C.POINTER(LmHooks), C.c_int, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 4
    hk = session.hooks()
    rows = np.zeros((cap, 4)); n = C.c_int(); lam = C.c_double(); ne = C.c_int(); tau = C.c_double()
    it = ref_lib.ref_lm_optimize(C.byref(hk), iterations, session.o.max_trials, _p(rows), cap, C.byref(n), C.byref(lam), C.byref(ne), C.byref(tau))
    return it, rows[:n.value].copy(), lam.value, ne.value, tau.value


def edge_quadratic_form(dim, Ji, Jj, er, w0, delta, robust, pose_fixed=False, lib=None):
    """One reprojection edge's contribution to the normal equations -> (hl 3x3, gl 3, hp 6x6, gp 6, w 6x3); `lib` = libref_lm.so runs the
    reference's BaseBinaryEdge::constructQuadraticForm instead of the oracle's."""
    Ji = np.ascontiguousarray(Ji, np.float64); Jj = np.ascontiguousarray(Jj, np.float64); er = np.ascontiguousarray(er, np.float64)
    out = [np.zeros(9), np.zeros(3), np.zeros(36), np.zeros(6), np.zeros(18)]
    vp = C.c_void_p
    if lib is None:
        f = ba_lib().ba_oracle_edge_quadratic_form
        f.argtypes = [C.c_int, vp, vp, vp, C.c_double, C.c_double, C.c_int] + [vp] * 5
        f(dim, _p(Ji), _p(Jj), _p(er), w0, delta, int(robust), _p(out[0]), _p(out[1]), None if pose_fixed else _p(out[2]), _p(out[3]), _p(out[4]))
    else:
        f = lib.ref_binary_quadratic_form
        f.argtypes = [C.c_int, vp, vp, vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int] + [vp] * 5
        f(dim, _p(Ji), _p(Jj), _p(er), w0, delta, int(robust), 0, int(pose_fixed), *[_p(x) for x in out])
    return out


def pose_quadratic_form(dim, J, er, w0, delta, robust, lib=None):
    """One OnlyPose edge's contribution (h 6x6, g 6); `lib` = libref_lm.so runs BaseUnaryEdge::constructQuadraticForm."""
    J = np.ascontiguousarray(J, np.float64); er = np.ascontiguousarray(er, np.float64)
    h, g = np.zeros(36), np.zeros(6)
    f = ba_lib().ba_oracle_pose_quadratic_form if lib is None else lib.ref_unary_quadratic_form
    f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    f(dim, _p(J), _p(er), w0, delta, int(robust), _p(h), _p(g))
    return h, g


def multi_quadratic_form(dim, dims, Js, er, w0, delta, robust, lib=None):
    """A three-vertex edge's (rigidity: dim 1, vertices 3 + 3 + 1; motion: dim 3, vertices 3 + 3 + 6) contribution to the normal equations as a
    dense (sum dims)^2 matrix and a right-hand side; `lib` = libref_lm.so runs the reference's BaseMultiEdge::constructQuadraticForm."""
    Js = [np.ascontiguousarray(J, np.float64) for J in Js]
    er = np.ascontiguousarray(er, np.float64)
    n = int(sum(dims)); H = np.zeros((n, n)); b = np.zeros(n)
    d = np.array(dims, np.int32)
    ptrs = (C.c_void_p * len(Js))(*[J.ctypes.data for J in Js])
    vp = C.c_void_p
    if lib is None:
        f = ba_lib().ba_oracle_multi_quadratic_form
        f.argtypes = [C.c_int, C.c_int, vp, vp, vp, C.c_double, C.c_double, C.c_int, vp, vp]
        f(dim, len(dims), _p(d), ptrs, _p(er), w0, delta, int(robust), _p(H), _p(b))
    else:
        fx = np.zeros(len(dims), np.uint8)
        f = lib.ref_multi_quadratic_form
        f.argtypes = [C.c_int, C.c_int, vp, vp, vp, C.c_double, C.c_double, C.c_int, vp, vp, vp]
        f(dim, len(dims), _p(d), ptrs, _p(er), w0, delta, int(robust), _p(fx), _p(H), _p(b))
    return H, b


def huber(delta: float, e2: float, lib=None):
    """RobustKernelHuber::robustify -> (rho, rho'); `lib` = libref_lm.so runs the reference's function."""
    if lib is not None:
        r = np.zeros(3)
        lib.ref_huber.argtypes = [C.c_double, C.c_double, C.c_void_p]
        lib.ref_huber(delta, e2, _p(r))
        return r[0], r[1]
    f = ba_lib().ba_oracle_huber
    f.argtypes = [C.c_double, C.c_double, C.c_void_p]
    r = np.zeros(2)
    f(delta, e2, _p(r))
    return r[0], r[1]


# ------------------------------------------------------------------------------------------
# oracle/ref_lba.cpp: the reference's Optimizer::LocalBundleAdjustment over the oracle's solver session
class _LbaBackend(C.Structure):
    _fields_ = ([("steps", LmHooks)] + [(n, C.c_void_p) for n in ("open", "close", "set_levels", "state", "lm_optimize", "default_options")] +
                [("pose_steps", LmHooks)] + [(n, C.c_void_p) for n in ("pose_open", "pose_close", "pose_state", "set_levels4")])


class _LbaIO(C.Structure):
    _fields_ = [("n_kf", C.c_int32), ("kf_id", C.c_void_p), ("kf_tcw", C.c_void_p), ("n_covisible", C.c_int32), ("covisible", C.c_void_p),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float),
                ("n_levels", C.c_int32), ("inv_level_sigma2", C.c_void_p),
                ("n_mp", C.c_int32), ("mp_id", C.c_void_p), ("mp_pos", C.c_void_p),
                ("n_obs", C.c_int32), ("obs_kf", C.c_void_p), ("obs_mp", C.c_void_p), ("obs_uvr", C.c_void_p), ("obs_octave", C.c_void_p),
                ("kf_tcw_out", C.c_void_p), ("mp_pos_out", C.c_void_p), ("erased", C.c_void_p), ("erased_cap", C.c_int32), ("n_erased", C.c_int32),
                ("mp_updates", C.c_void_p)]


class _HumanIO(C.Structure):
    _fields_ = [("n_traj", C.c_int32), ("traj_id", C.c_void_p), ("traj_track_id", C.c_void_p), ("traj_n_poses", C.c_void_p),
                ("rigid_id", C.c_void_p), ("rigid_dist", C.c_void_p),
                ("n_hp", C.c_int32), ("hp_id", C.c_void_p), ("hp_traj", C.c_void_p), ("hp_ref_kf", C.c_void_p), ("hp_time", C.c_void_p),
                ("key_id", C.c_void_p), ("key_pos", C.c_void_p), ("key_uvr", C.c_void_p),
                ("n_current", C.c_int32), ("current_hp", C.c_void_p),
                ("sigma_static", C.c_float), ("sigma_human", C.c_float), ("sigma_rigidity", C.c_float), ("sigma_motion", C.c_float),
                ("th_motion", C.c_float), ("th_rigidity", C.c_float),
                ("key_pos_out", C.c_void_p), ("key_flags", C.c_void_p), ("pair_flags", C.c_void_p), ("traj_out", C.c_void_p), ("traj_motion", C.c_void_p),
                ("n_optimized_tracks", C.c_int32)]


def ref_local_bundle_adjustment(lba_lib, lm_lib, w: dict, global_ba=None, humans: dict | None = None):
    """Runs the reference's Optimizer::LocalBundleAdjustment (oracle/_ref/libref_lba.so) on window `w` (kf_id, kf_tcw [n,4,4] f32, covisible,
    fx fy cx cy bf, inv_level_sigma2, mp_id, mp_pos [m,3] f32, obs_kf, obs_mp, obs_uvr [o,3] f32, obs_octave; key-frame 0 = pKF), its
    solver steps being the oracle's and its LM control the reference's (libref_lm.so).  Returns the function's outputs and the problem /
    trial rows the stand-in optimizer recorded (a dict in make_ba_problem layout, usable with ba_solve).
    global_ba = (nIterations, nLoopKF, bRobust) runs Optimizer::BundleAdjustment (what GlobalBundleAdjustemnt calls) on ALL key-frames and
    points of `w` instead; with nLoopKF != 0 the outputs are mTcwGBA / mPosGBA and mp_updates holds mnBAGlobalForKF."""
    be = _lba_backend(lm_lib)
    a = lambda k, t: np.ascontiguousarray(w[k], t)
    kf_id, tcw, cov = a("kf_id", np.int32), a("kf_tcw", np.float32), a("covisible", np.int32)
    sig, mp_id, mp_pos = a("inv_level_sigma2", np.float32), a("mp_id", np.int32), a("mp_pos", np.float32)
    okf, omp, uvr, octv = a("obs_kf", np.int32), a("obs_mp", np.int32), a("obs_uvr", np.float32), a("obs_octave", np.int32)
    tcw_out = np.zeros_like(tcw); pos_out = np.zeros_like(mp_pos); erased = np.zeros((len(okf), 2), np.int32); upd = np.zeros(len(mp_id), np.int32)
    io = _LbaIO(len(kf_id), _p(kf_id), _p(tcw), len(cov), _p(cov), w["fx"], w["fy"], w["cx"], w["cy"], w["bf"], len(sig), _p(sig),
                len(mp_id), _p(mp_id), _p(mp_pos), len(okf), _p(okf), _p(omp), _p(uvr), _p(octv), _p(tcw_out), _p(pos_out), _p(erased), len(erased), 0, _p(upd))
    rec = C.c_void_p()
    lba_lib.ref_lba_run.argtypes = [C.POINTER(_LbaBackend), C.POINTER(_LbaIO), C.POINTER(C.c_void_p)]
    lba_lib.ref_lba_record_f64.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lba_lib.ref_lba_record_i32.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lba_lib.ref_lba_record_free.argtypes = [C.c_void_p]
    hout = None
    if humans is not None:
        hh = {k: np.ascontiguousarray(humans[k], t) for k, t in (("traj_id", np.int32), ("traj_track_id", np.int32), ("traj_n_poses", np.int32),
              ("rigid_id", np.int32), ("rigid_dist", np.float32), ("hp_id", np.int32), ("hp_traj", np.int32), ("hp_ref_kf", np.int32), ("hp_time", np.float64),
              ("key_id", np.int32), ("key_pos", np.float32), ("key_uvr", np.float32), ("current_hp", np.int32))}
        nt, nh = len(hh["traj_id"]), len(hh["hp_id"])
        hout = dict(key_pos=np.zeros((nh, 14, 3), np.float32), key_flags=np.zeros((nh, 14, 5), np.uint8), pair_flags=np.zeros((nh, 14, 2), np.uint8),
                    traj_out=np.zeros((nt, 2), np.int32), traj_motion=np.zeros((nt, 4, 4), np.float32))
        hio = _HumanIO(nt, _p(hh["traj_id"]), _p(hh["traj_track_id"]), _p(hh["traj_n_poses"]), _p(hh["rigid_id"]), _p(hh["rigid_dist"]),
                       nh, _p(hh["hp_id"]), _p(hh["hp_traj"]), _p(hh["hp_ref_kf"]), _p(hh["hp_time"]), _p(hh["key_id"]), _p(hh["key_pos"]), _p(hh["key_uvr"]),
                       len(hh["current_hp"]), _p(hh["current_hp"]), 1.0, humans["sigma_human"], humans["sigma_rigidity"], humans["sigma_motion"],
                       humans["th_motion"], humans["th_rigidity"], _p(hout["key_pos"]), _p(hout["key_flags"]), _p(hout["pair_flags"]), _p(hout["traj_out"]),
                       _p(hout["traj_motion"]), 0)
        lba_lib.ref_hba_run.argtypes = [C.POINTER(_LbaBackend), C.POINTER(_LbaIO), C.POINTER(_HumanIO), C.POINTER(C.c_void_p)]
        rc = lba_lib.ref_hba_run(C.byref(be), C.byref(io), C.byref(hio), C.byref(rec))
        hout["n_optimized_tracks"] = hio.n_optimized_tracks
    elif global_ba is None:
        rc = lba_lib.ref_lba_run(C.byref(be), C.byref(io), C.byref(rec))
    else:
        lba_lib.ref_gba_run.argtypes = [C.POINTER(_LbaBackend), C.POINTER(_LbaIO), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        rc = lba_lib.ref_gba_run(C.byref(be), C.byref(io), int(global_ba[0]), int(global_ba[1]), int(global_ba[2]), C.byref(rec))
    assert rc == 0

    def f64(which):
        n = lba_lib.ref_lba_record_f64(rec, which, None, 0); v = np.zeros(n); lba_lib.ref_lba_record_f64(rec, which, _p(v), n); return v

    def i32(which):
        n = lba_lib.ref_lba_record_i32(rec, which, None, 0); v = np.zeros(n, np.int32); lba_lib.ref_lba_record_i32(rec, which, _p(v), n); return v

    cam = f64(7)
    prob = dict(fx=cam[0], fy=cam[1], cx=cam[2], cy=cam[3], bf=cam[4], pose_q=f64(0).reshape(-1, 4), pose_t=f64(1).reshape(-1, 3),
                points=f64(2).reshape(-1, 3), edge_obs=f64(3).reshape(-1, 3), edge_info=f64(4), edge_pose=i32(0), edge_point=i32(1),
                pose_fixed=i32(6).astype(np.uint8))
    out = dict(kf_tcw=tcw_out, mp_pos=pos_out, erased=erased[:io.n_erased].copy(), mp_updates=upd, problem=prob, rows=f64(5).reshape(-1, 4),
               final_state=f64(6), pose_id=i32(2), point_id=i32(3), round_iterations=i32(4), round_robust=i32(5), huber=(cam[5], cam[6]))
    if humans is not None:
        lba_lib.ref_lba_record_named.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]

        def named(name, dt=np.float64):
            m = lba_lib.ref_lba_record_named(rec, name.encode(), None, 0)
            v = np.zeros(max(m, 0)); lba_lib.ref_lba_record_named(rec, name.encode(), _p(v), len(v))
            return v.astype(dt)

        for k in ("joints", "dists", "motion_q", "motion_t", "jedge_obs", "jedge_info", "redge_info", "medge_dt", "medge_info"):
            prob[k] = named(k)
        for k in ("jedge_pose", "jedge_joint", "redge_i", "redge_j", "redge_dist", "medge_p1", "medge_p2", "medge_motion"):
            prob[k] = named(k, np.int32)
        prob["joints"] = prob["joints"].reshape(-1, 3); prob["motion_q"] = prob["motion_q"].reshape(-1, 4); prob["motion_t"] = prob["motion_t"].reshape(-1, 3)
        prob["jedge_obs"] = prob["jedge_obs"].reshape(-1, 3)
        out.update(humans=hout, joint_id=named("joint_id", np.int32), dist_id=named("dist_id", np.int32), motion_id=named("motion_id", np.int32),
                   edge_kind=named("edge_kind", np.int32), edge_final_chi2=named("edge_final_chi2"), edge_final_depth_positive=named("edge_final_depth_positive", np.int32),
                   huber_rigid=float(named("huber_rigid")[0]), huber_motion=float(named("huber_motion")[0]))
    lba_lib.ref_lba_record_free(rec)
    return out


class _PoseIO(C.Structure):
    _fields_ = [("tcw", C.c_void_p), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float),
                ("n_levels", C.c_int32), ("inv_level_sigma2", C.c_void_p), ("n", C.c_int32), ("uvr", C.c_void_p), ("octave", C.c_void_p),
                ("xw", C.c_void_p), ("has_point", C.c_void_p), ("outlier", C.c_void_p), ("tcw_out", C.c_void_p), ("n_inliers", C.c_int32)]


def _lba_backend(lm_lib):
    lib = ba_lib()
    lib.ba_oracle_lm_open.restype = C.c_void_p
    lib.ba_oracle_pose_lm_open.restype = C.c_void_p
    be = _LbaBackend()
    be.steps = LmHooks(None, *[C.cast(getattr(lib, "ba_oracle_lm_" + n), C.c_void_p) for n in _LM_HOOKS])
    be.pose_steps = LmHooks(None, *[C.cast(getattr(lib, "ba_oracle_pose_lm_" + n), C.c_void_p) for n in _LM_HOOKS])
    for n, f in (("open", lib.ba_oracle_lm_open), ("close", lib.ba_oracle_lm_close), ("set_levels", lib.ba_oracle_lm_set_levels),
                 ("state", lib.ba_oracle_lm_state), ("lm_optimize", lm_lib.ref_lm_optimize), ("default_options", lib.ba_oracle_default_options),
                 ("pose_open", lib.ba_oracle_pose_lm_open), ("pose_close", lib.ba_oracle_pose_lm_close), ("pose_state", lib.ba_oracle_pose_lm_state),
                 ("set_levels4", lib.ba_oracle_lm_set_levels4)):
        setattr(be, n, C.cast(f, C.c_void_p))
    return be


def ref_pose_optimization(lba_lib, lm_lib, fr: dict):
    """Runs the reference's Optimizer::PoseOptimization (oracle/_ref/libref_lba.so) on frame `fr` (tcw [4,4] f32, fx fy cx cy bf,
    inv_level_sigma2, uvr [n,3] f32, octave [n], xw [n,3] f32, has_point [n]) over the oracle's pose session and the reference's LM control.
    Returns mvbOutlier, the return value, the pose written back and what the stand-in optimizer recorded."""
    be = _lba_backend(lm_lib)
    a = lambda k, t: np.ascontiguousarray(fr[k], t)
    tcw, sig, uvr, octv, xw, has = a("tcw", np.float32), a("inv_level_sigma2", np.float32), a("uvr", np.float32), a("octave", np.int32), a("xw", np.float32), a("has_point", np.uint8)
    n = len(octv)
    outl = np.zeros(max(n, 1), np.uint8); tcw_out = np.zeros((4, 4), np.float32)
    io = _PoseIO(_p(tcw), fr["fx"], fr["fy"], fr["cx"], fr["cy"], fr["bf"], len(sig), _p(sig), n, _p(uvr), _p(octv), _p(xw), _p(has), _p(outl), _p(tcw_out), 0)
    rec = C.c_void_p()
    lba_lib.ref_pose_run.argtypes = [C.POINTER(_LbaBackend), C.POINTER(_PoseIO), C.POINTER(C.c_void_p)]
    for f in (lba_lib.ref_lba_record_f64, lba_lib.ref_lba_record_i32, lba_lib.ref_lba_record_f32):
        f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lba_lib.ref_lba_record_free.argtypes = [C.c_void_p]
    assert lba_lib.ref_pose_run(C.byref(be), C.byref(io), C.byref(rec)) == 0

    def get(fn, which, dt):
        m = fn(rec, which, None, 0); v = np.zeros(m, dt); fn(rec, which, _p(v), m); return v

    cam = get(lba_lib.ref_lba_record_f64, 7, np.float64)
    out = dict(outlier=outl[:n].copy(), n_inliers=io.n_inliers, tcw=tcw_out, rows=get(lba_lib.ref_lba_record_f64, 5, np.float64).reshape(-1, 4),
               final_state=get(lba_lib.ref_lba_record_f64, 6, np.float64), round_iterations=get(lba_lib.ref_lba_record_i32, 4, np.int32),
               round_robust=get(lba_lib.ref_lba_record_i32, 5, np.int32), round_active=get(lba_lib.ref_lba_record_i32, 7, np.int32),
               cam=dict(fx=cam[0], fy=cam[1], cx=cam[2], cy=cam[3], bf=cam[4]),
               frame=dict(pose_q=get(lba_lib.ref_lba_record_f64, 0, np.float64), pose_t=get(lba_lib.ref_lba_record_f64, 1, np.float64),
                          xw=get(lba_lib.ref_lba_record_f32, 0, np.float32).reshape(-1, 3), obs=get(lba_lib.ref_lba_record_f32, 1, np.float32).reshape(-1, 3),
                          inv_sigma2=get(lba_lib.ref_lba_record_f32, 2, np.float32)))
    lba_lib.ref_lba_record_free(rec)
    return out


def pose_optimize_traced(cam: dict, frame: dict):
    """The oracle's four-round PoseOptimization on one frame with its LM trials -> (PoseBatch, rows [(lambda, chi2 before, after, accepted)])."""
    from airdos_b200 import ba_types as T
    pb = T.PoseBatch(cam, [frame])
    lib = ba_lib()
    lib.ba_oracle_pose_optimize_traced.argtypes = [C.POINTER(T.PoseProblem), C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    rows = np.zeros((512, 4)); n = C.c_int()
    lib.ba_oracle_pose_optimize_traced(C.byref(pb.c), 0, _p(rows), 512, C.byref(n))
    return pb, rows[:n.value].copy()


# ------------------------------------------------------------------------------------------
# oracle/ref_orb.cpp: the reference's ORBextractor::operator() over the oracle's pixel primitives
class _OrbPrims(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("border101", "erode10", "resize", "blur7", "fast")]


def ref_orb_extract(ref_lib, img: np.ndarray, mask: np.ndarray | None = None, nfeatures: int = 1000, scale: float = 1.2, nlevels: int = 8,
                    ini_th: int = 20, min_th: int = 7, monotonic: bool = True):
    """ORBextractor::operator() / ComputePyramid / ComputeKeyPointsOctTree of the reference (oracle/_ref/libref_orb.so) with cv::resize,
    copyMakeBorder, erode, GaussianBlur and FastFeatureDetector landing in the oracle's primitives -> (key-points KP_DTYPE, descriptors, pyramid)."""
    lib = orb_lib()
    pr = _OrbPrims(*[C.cast(getattr(lib, "orb_oracle_" + n), C.c_void_p) for n in ("border101", "erode10", "resize", "blur7", "fast")])
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
    cap = nfeatures * 2 + 64
    kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
    p = orb_params(nfeatures, scale, nlevels, w, h)
    pyr = np.zeros(int((p["w"].astype(np.int64) * p["h"]).sum()), np.uint8)
    ref_lib.ref_orb_extract.argtypes = [C.POINTER(_OrbPrims), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    n = ref_lib.ref_orb_extract(C.byref(pr), _p(img), w, h, None if m is None else _p(m), nfeatures, scale, nlevels, ini_th, min_th, _p(kps), _p(desc), cap,
                                _p(pyr), int(monotonic))
    assert n >= 0
    return kps[:n].copy(), desc[:n].copy(), pyr
