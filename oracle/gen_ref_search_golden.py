"""Writes tests/golden/search_ref.npz: outputs of the REFERENCE's own tracking searches -- ORBmatcher::SearchByProjection(Frame&,
const Frame&, th, bMono) (src/ORBmatcher.cc:1328-1470) and SearchByProjection(Frame&, vector<MapPoint*>&, th) (:45-129), with
Frame::AssignFeaturesToGrid / GetFeaturesInArea / PosInGrid (src/Frame.cc:534-549, 645-712), ComputeThreeMaxima and
DescriptorDistance -- compiled from /root/reference by `make -C oracle ref` (oracle/ref_match.cpp) and run on the seeded tracking
problems of airdos_b200/synth.py.  The fixture stores the problems' seeds, a CRC of their arrays and the reference's results
(final mvpMapPoints as query indices, nmatches).  Run in the build container (needs /root/reference):

    python oracle/gen_ref_search_golden.py
"""
import ctypes as C
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_match.so")

LAST_CASES = [  # (seed, n_kp, n_q, th, last_dz, mono, check_orientation): forward / backward / lateral motion, mono flag, no rotation check
    (5, 2000, 1500, 7.0, 0.02, 0, 1), (6, 2000, 1500, 15.0, 0.6, 0, 1), (7, 3000, 2500, 7.0, -0.6, 0, 1), (8, 1500, 1200, 7.0, 0.6, 1, 1),
    (9, 2500, 2000, 10.0, 0.02, 0, 0)]
MAP_CASES = [  # (seed, n_kp, n_q, th, nn_ratio)
    (15, 2000, 1500, 1.0, 0.8), (16, 3000, 3000, 3.0, 0.8), (17, 2000, 1800, 5.0, 0.6)]
FUSE_CASES = [  # ORBmatcher::Fuse(pKF, vpMapPoints, th): (seed, n_kp, n_q, th)
    (35, 2000, 3000, 3.0), (36, 3000, 2500, 4.0), (37, 1500, 2000, 2.5)]
LOCAL_CASES = [  # Tracking::SearchLocalPoints: isInFrustum + PredictScale + SearchByProjection(F, vpMapPoints, th): (seed, n_kp, n_q, th, nn_ratio)
    (25, 2000, 2500, 1.0, 0.8), (26, 3000, 3000, 3.0, 0.8), (27, 1500, 2000, 5.0, 0.8)]


def P(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def crc(*arrays):
    c = 0
    for a in arrays:
        c = zlib.crc32(np.ascontiguousarray(a).tobytes(), c)
    return c


def last_problem(case):
    from airdos_b200 import synth
    seed, n_kp, n_q, th, dz, mono, chk = case
    pr = synth.make_tracking_problem(seed, n_kp=n_kp, n_q=n_q, th=th, last_dz=dz)
    pr["mono"] = mono; pr["check_orientation"] = chk
    return pr


def map_problem(case):
    """Projected map points in the reference's own terms: mTrackProjX / Y / XR, mnTrackScaleLevel, mTrackViewCos; the search radius
    the oracle takes as an array is formed exactly as src/ORBmatcher.cc:63-69 forms it (float operations in the same order)."""
    import oracle
    from airdos_b200 import synth
    seed, n_kp, n_q, th, nn = case
    pr = synth.make_tracking_problem(seed, n_kp=n_kp, n_q=n_q)
    proj = oracle.search_by_projection(pr)[4]
    rng = np.random.default_rng(seed + 1000)
    view_cos = np.where(rng.random(n_q) < 0.4, 0.9995, rng.uniform(0.5, 0.99, n_q)).astype(np.float32)
    r = np.where(view_cos > np.float32(0.998), np.float32(2.5), np.float32(4.0)).astype(np.float32)
    if th != 1.0:
        r = (r * np.float32(th)).astype(np.float32)
    level = np.asarray(pr["last_octave"], np.int32)
    sf = np.asarray(pr["scale_factors"], np.float32)
    out = {k: pr[k] for k in ("kps", "u_right", "desc", "taken", "bounds", "q_desc", "q_angle", "scale_factors")}
    out.update(q_u=proj["q_u"], q_v=proj["q_v"], q_ur=proj["q_ur"], q_radius=(r * sf[level]).astype(np.float32),
               q_min_level=(level - 1).astype(np.int32), q_max_level=level.copy(), q_flags=proj["q_flags"], use_ratio=1, nn_ratio=nn,
               check_orientation=0, view_cos=view_cos, level=level, th=th)
    return out


def local_problem(case, ow=None):
    """Local map points before the visibility test.  `ow`: the camera centre as the reference's UpdatePoseMatrices derives it from Tcw
    (float gemm); the generator asks the reference for it, the tests take it from the fixture."""
    from airdos_b200 import synth
    seed, n_kp, n_q, th, nn = case
    pr = synth.make_tracking_problem(seed, n_kp=n_kp, n_q=n_q)
    pm = synth.tracking_problem_as_local_map(pr, seed=seed, nn_ratio=nn, th=th)
    if ow is not None:
        pm["ow"] = np.asarray(ow, np.float32)
    return pm


def ref_local(L, pm):
    from airdos_b200.capi import KP_DTYPE
    kps = np.ascontiguousarray(pm["kps"], KP_DTYPE); nk = len(kps); nq = len(pm["q_flags"])
    fx, fy, cx, cy, mbf, mb = [float(v) for v in pm["cam"]]
    mnx, mny, mxx, mxy = [float(v) for v in pm["bounds"]]
    sf = np.ascontiguousarray(pm["scale_factors"], np.float32)
    a = [np.ascontiguousarray(pm[k], t) for k, t in (("u_right", np.float32), ("desc", np.uint8), ("taken", np.uint8), ("tcw_cur", np.float32),
                                                     ("mp_xw", np.float32), ("mp_normal", np.float32), ("mp_min_distance", np.float32),
                                                     ("mp_max_distance", np.float32), ("q_desc", np.uint8), ("q_flags", np.uint8))]
    km = np.zeros(nk, np.int32); inview = np.zeros(nq, np.uint8); track = np.zeros((nq, 4), np.float32); level = np.zeros(nq, np.int32); ow = np.zeros(3, np.float32)
    L.ref_search_local_map.restype = C.c_int
    L.ref_search_local_map.argtypes = ([C.c_void_p] * 4 + [C.c_int] + [C.c_float] * 4 + [C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 6 +
                                       [C.c_float] * 9 + [C.c_void_p] * 5)
    n = L.ref_search_local_map(P(kps), P(a[0]), P(a[1]), P(a[2]), nk, mnx, mny, mxx, mxy, P(sf), len(sf), P(a[3]), nq, P(a[4]), P(a[5]), P(a[6]), P(a[7]),
                               P(a[8]), P(a[9]), fx, fy, cx, cy, mbf, float(pm["view_cos_limit"]), float(pm["log_scale_factor"]), float(pm["th"]),
                               float(pm["nn_ratio"]), P(km), P(inview), P(track), P(level), P(ow))
    return int(n), km, inview, track, level, ow


def fuse_problem(case, ow=None):
    from airdos_b200 import synth
    seed, n_kp, n_q, th = case
    pr = synth.make_tracking_problem(seed, n_kp=n_kp, n_q=n_q, dup_frac=0.3)
    pf = synth.tracking_problem_as_fuse(pr, seed=seed, th=th)
    if ow is not None:
        pf["ow"] = np.asarray(ow, np.float32)
    return pf


def ref_fuse(L, pf):
    from airdos_b200.capi import KP_DTYPE
    kps = np.ascontiguousarray(pf["kps"], KP_DTYPE); nk = len(kps); nq = len(pf["q_flags"])
    fx, fy, cx, cy, mbf, mb = [float(v) for v in pf["cam"]]
    mnx, mny, mxx, mxy = [float(v) for v in pf["bounds"]]
    sf = np.ascontiguousarray(pf["scale_factors"], np.float32); isg = np.ascontiguousarray(pf["inv_level_sigma2"], np.float32)
    a = [np.ascontiguousarray(pf[k], t) for k, t in (("u_right", np.float32), ("desc", np.uint8), ("tcw_cur", np.float32), ("mp_xw", np.float32),
                                                     ("mp_normal", np.float32), ("mp_min_distance", np.float32), ("mp_max_distance", np.float32),
                                                     ("q_desc", np.uint8), ("q_flags", np.uint8))]
    fused = np.zeros(nq, np.int32); ow = np.zeros(3, np.float32)
    L.ref_fuse.restype = C.c_int
    L.ref_fuse.argtypes = ([C.c_void_p] * 3 + [C.c_int] + [C.c_float] * 4 + [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 6 +
                           [C.c_float] * 7 + [C.c_void_p, C.c_void_p])
    n = L.ref_fuse(P(kps), P(a[0]), P(a[1]), nk, mnx, mny, mxx, mxy, P(sf), P(isg), len(sf), P(a[2]), nq, P(a[3]), P(a[4]), P(a[5]), P(a[6]), P(a[7]), P(a[8]),
                   fx, fy, cx, cy, mbf, float(pf["log_scale_factor"]), float(pf["th"]), P(fused), P(ow))
    return int(n), fused, ow


def ref_last(L, pr):
    from airdos_b200.capi import KP_DTYPE
    kps = np.ascontiguousarray(pr["kps"], KP_DTYPE); nk = len(kps); nq = len(pr["q_flags"])
    fx, fy, cx, cy, mbf, mb = [float(v) for v in pr["cam"]]
    mnx, mny, mxx, mxy = [float(v) for v in pr["bounds"]]
    sf = np.ascontiguousarray(pr["scale_factors"], np.float32)
    km = np.zeros(nk, np.int32)
    a = [np.ascontiguousarray(pr[k], t) for k, t in (("u_right", np.float32), ("desc", np.uint8), ("taken", np.uint8), ("tcw_cur", np.float32),
                                                     ("tcw_last", np.float32), ("last_xw", np.float32), ("last_octave", np.int32), ("q_angle", np.float32),
                                                     ("q_desc", np.uint8), ("q_flags", np.uint8))]
    L.ref_search_last_frame.restype = C.c_int
    L.ref_search_last_frame.argtypes = ([C.c_void_p] * 4 + [C.c_int] + [C.c_float] * 4 + [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int] +
                                        [C.c_void_p] * 5 + [C.c_float] * 7 + [C.c_int, C.c_int, C.c_void_p])
    n = L.ref_search_last_frame(P(kps), P(a[0]), P(a[1]), P(a[2]), nk, mnx, mny, mxx, mxy, P(sf), len(sf), P(a[3]), P(a[4]), nq, P(a[5]), P(a[6]), P(a[7]),
                                P(a[8]), P(a[9]), fx, fy, cx, cy, mbf, mb, float(pr["th"]), int(pr["mono"]), int(pr["check_orientation"]), P(km))
    return int(n), km


def ref_map(L, pr):
    from airdos_b200.capi import KP_DTYPE
    kps = np.ascontiguousarray(pr["kps"], KP_DTYPE); nk = len(kps); nq = len(pr["q_flags"])
    mnx, mny, mxx, mxy = [float(v) for v in pr["bounds"]]
    sf = np.ascontiguousarray(pr["scale_factors"], np.float32)
    km = np.zeros(nk, np.int32)
    a = [np.ascontiguousarray(pr[k], t) for k, t in (("u_right", np.float32), ("desc", np.uint8), ("taken", np.uint8), ("q_u", np.float32), ("q_v", np.float32),
                                                     ("q_ur", np.float32), ("level", np.int32), ("view_cos", np.float32), ("q_desc", np.uint8), ("q_flags", np.uint8))]
    L.ref_search_map_points.restype = C.c_int
    L.ref_search_map_points.argtypes = ([C.c_void_p] * 4 + [C.c_int] + [C.c_float] * 4 + [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 7 +
                                        [C.c_float, C.c_float, C.c_void_p])
    n = L.ref_search_map_points(P(kps), P(a[0]), P(a[1]), P(a[2]), nk, mnx, mny, mxx, mxy, P(sf), len(sf), nq, P(a[3]), P(a[4]), P(a[5]), P(a[6]), P(a[7]),
                                P(a[8]), P(a[9]), float(pr["th"]), float(pr["nn_ratio"]), P(km))
    return int(n), km


def problem_crc(pr):
    keys = [k for k in ("kps", "u_right", "desc", "taken", "q_desc", "q_angle", "q_flags", "last_xw", "last_octave", "tcw_cur", "tcw_last", "q_u", "q_v",
                        "q_ur", "q_radius", "view_cos", "level", "mp_xw", "mp_normal", "mp_min_distance", "mp_max_distance") if k in pr]
    return crc(*[pr[k] for k in keys])


def main():
    import oracle
    oracle.build()
    L = C.CDLL(LIB)
    out = {"last_cases": np.array(LAST_CASES, np.float64), "map_cases": np.array(MAP_CASES, np.float64), "local_cases": np.array(LOCAL_CASES, np.float64),
           "fuse_cases": np.array(FUSE_CASES, np.float64)}
    for i, case in enumerate(LAST_CASES):
        pr = last_problem(case)
        n, km = ref_last(L, pr)
        out[f"last{i}_n"] = np.int32(n); out[f"last{i}_kp_match"] = km; out[f"last{i}_crc"] = np.int64(problem_crc(pr))
        print(f"last-frame case {i}: {n} matches of {len(pr['q_flags'])} map points")
    for i, case in enumerate(MAP_CASES):
        pr = map_problem(case)
        n, km = ref_map(L, pr)
        out[f"map{i}_n"] = np.int32(n); out[f"map{i}_kp_match"] = km; out[f"map{i}_crc"] = np.int64(problem_crc(pr))
        print(f"map-point case {i}: {n} matches of {len(pr['q_flags'])} map points")
    for i, case in enumerate(FUSE_CASES):
        pf = fuse_problem(case)
        n, fused, ow = ref_fuse(L, pf)
        out[f"fuse{i}_n"] = np.int32(n); out[f"fuse{i}_fused_with"] = fused; out[f"fuse{i}_crc"] = np.int64(problem_crc(pf)); out[f"fuse{i}_ow"] = ow
        print(f"fuse case {i}: {n} of {len(fused)} map points fused")
    for i, case in enumerate(LOCAL_CASES):
        pm = local_problem(case)
        n, km, inview, track, level, ow = ref_local(L, pm)
        out[f"local{i}_n"] = np.int32(n); out[f"local{i}_kp_match"] = km; out[f"local{i}_crc"] = np.int64(problem_crc(pm))
        out[f"local{i}_in_view"] = inview; out[f"local{i}_track"] = track; out[f"local{i}_level"] = level; out[f"local{i}_ow"] = ow
        print(f"local-map case {i}: {int(inview.sum())} of {len(inview)} points in view, {n} matches; |ow - float64 ow| = {np.abs(ow - pm['ow']).max():.2e}")
    path = os.path.join(ROOT, "tests", "golden", "search_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
